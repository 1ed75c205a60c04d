import os

import numpy as np
import torch

import paramgen

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

CASES = {
    "blurry_t3_32": (1, 3, 32, 32, 26, 2, False),
    "sharp_t2_b2_48x32": (2, 2, 48, 32, 6, 2, True),
    "deblur_t1_ev5_64": (1, 1, 64, 64, 3, 5, False),
    "blurry_t4_64": (1, 4, 64, 64, 26, 2, False),
}


def load(case):
    z = np.load(os.path.join(GOLDEN_DIR, case + ".npz"), allow_pickle=False)
    names = [str(n) for n in z["names"]]
    sizes = z["grad_sample_sizes"]
    offs = np.concatenate([[0], np.cumsum(sizes)])
    samples = {n: torch.from_numpy(z["grad_samples"][offs[i]:offs[i + 1]]) for i, n in enumerate(names)}
    norms = {n: float(z["grad_norm"][i]) for i, n in enumerate(names)}
    return {"out": torch.from_numpy(z["out"]), "loss": float(z["loss"]), "names": names,
            "grad_norm": norms, "grad_samples": samples, "dead": [str(d) for d in z["dead"]]}


def check_grads(grads, gold, rtol_norm, atol_rel_samples):
    """grads: name -> tensor. Norms within rtol_norm (relative); sampled elements within
    atol_rel_samples * max(|golden norm| / sqrt(numel) * 10, tiny)."""
    bad = []
    for n in gold["names"]:
        g = grads[n].detach().float().cpu()
        gn = gold["grad_norm"][n]
        if n in gold["dead"]:
            if g.abs().max().item() != 0.0:
                bad.append((n, "dead param has non-zero grad"))
            continue
        mine = g.double().norm().item()
        if abs(mine - gn) > rtol_norm * max(gn, 1e-12):
            bad.append((n, f"norm {mine:.6g} vs {gn:.6g}"))
        idx = paramgen.grad_sample_index(n, g.numel())
        scale = gn / max(g.numel(), 1) ** 0.5
        err = (g.flatten()[idx] - gold["grad_samples"][n]).abs().max().item()
        if err > atol_rel_samples * max(scale, 1e-12):
            bad.append((n, f"sample err {err:.3g} vs rms {scale:.3g}"))
    return bad


def sample_error_ratios(grads, gold):
    """name -> max |sampled gradient element - golden| / (golden tensor RMS), live parameters only."""
    out = {}
    for n in gold["names"]:
        if n in gold["dead"] or gold["grad_norm"][n] <= 0:
            continue
        g = grads[n].detach().float().cpu()
        idx = paramgen.grad_sample_index(n, g.numel())
        scale = gold["grad_norm"][n] / max(g.numel(), 1) ** 0.5
        out[n] = (g.flatten()[idx] - gold["grad_samples"][n]).abs().max().item() / max(scale, 1e-30)
    return out
