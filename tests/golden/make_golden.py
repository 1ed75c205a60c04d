"""Generate golden vectors from the UNMODIFIED reference (run in the build container only):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

For each case: deterministic parameters (tests/paramgen.py) are loaded into the reference module with
strict=True, the reference runs forward + Charbonnier + backward in fp32 on CPU, and the output plus, per
parameter, the gradient L2 norm and 32 sampled gradient elements are stored in tests/golden/<case>.npz.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import ref_loader  # noqa: E402
import paramgen  # noqa: E402

CASES = {
    # name: (B, T, H, W, img_chn, ev_chn, x5d)
    "blurry_t3_32": (1, 3, 32, 32, 26, 2, False),
    "sharp_t2_b2_48x32": (2, 2, 48, 32, 6, 2, True),
    "deblur_t1_ev5_64": (1, 1, 64, 64, 3, 5, False),
    "blurry_t4_64": (1, 4, 64, 64, 26, 2, False),
}


def main():
    torch.set_num_threads(8)
    for name, (B, T, H, W, ic, ec, x5d) in CASES.items():
        net = ref_loader.build(ic, ec)
        shapes = {k: tuple(v.shape) for k, v in net.state_dict().items()}
        P = paramgen.make_params(shapes, seed=0)
        net.load_state_dict(P, strict=True)
        x, ev, gt = paramgen.make_inputs(B, T, H, W, ic, ec, x5d=x5d)
        out = net(x=x, event=ev)
        loss = torch.sqrt((out - gt) ** 2 + 1e-12).mean()
        loss.backward()
        rec = {"out": out.detach().numpy(), "loss": np.float64(loss.item()),
               "names": np.array(sorted(shapes))}
        gn, gs, dead = [], [], []
        for k in sorted(shapes):
            p = dict(net.named_parameters())[k]
            if p.grad is None:
                dead.append(k)
                gn.append(0.0)
                gs.append(np.zeros(min(32, p.numel()), np.float32))
                continue
            gn.append(p.grad.double().norm().item())
            idx = paramgen.grad_sample_index(k, p.numel())
            gs.append(p.grad.flatten()[idx].numpy())
        rec["grad_norm"] = np.array(gn)
        rec["grad_samples"] = np.concatenate(gs)
        rec["grad_sample_sizes"] = np.array([len(g) for g in gs])
        rec["dead"] = np.array(dead)
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + ".npz"), **rec)
        print(name, "out", tuple(out.shape), "loss", loss.item(), "dead", len(dead))


if __name__ == "__main__":
    main()
