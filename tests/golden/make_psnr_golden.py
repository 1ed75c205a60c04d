"""Golden vectors for the validation-metric restatement, from the UNMODIFIED reference functions (build container only):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_psnr_golden.py

Imports basicsr/utils/img_util.py (`tensor2img`, :59-121) and basicsr/metrics/psnr_ssim.py (`calculate_psnr`, :9-61) of
/root/reference through stub packages (their package __init__ chains need lmdb / skimage, absent here; `skimage.metrics`
is stubbed -- calculate_psnr does not use it) and stores, per case, the float input frames, the uint8 images tensor2img
returns and the PSNR calculate_psnr returns -> tests/golden/psnr_cases.npz.
"""
import importlib
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = os.environ.get("REFID_REFERENCE", "/root/reference")


def load_reference_functions():
    sys.dont_write_bytecode = True
    for name, path in (("basicsr", f"{REF}/basicsr"), ("basicsr.utils", f"{REF}/basicsr/utils"),
                       ("basicsr.metrics", f"{REF}/basicsr/metrics")):
        m = types.ModuleType(name)
        m.__path__ = [path]
        sys.modules[name] = m
    sk = types.ModuleType("skimage")
    sk.metrics = types.ModuleType("skimage.metrics")
    sys.modules["skimage"], sys.modules["skimage.metrics"] = sk, sk.metrics
    img_util = importlib.import_module("basicsr.utils.img_util")
    psnr = importlib.import_module("basicsr.metrics.psnr_ssim")
    return img_util.tensor2img, psnr.calculate_psnr


def cases():
    g = torch.Generator().manual_seed(5)
    out = {}
    a = torch.rand(3, 24, 40, generator=g)
    out["uniform"] = (a, (a + 0.02 * torch.randn(a.shape, generator=g)), 0)
    out["out_of_range"] = (1.4 * torch.rand(3, 16, 16, generator=g) - 0.2, torch.rand(3, 16, 16, generator=g), 0)
    b = torch.rand(3, 32, 32, generator=g)
    out["crop4"] = (b, (b + 0.05 * torch.randn(b.shape, generator=g)), 4)
    ties = (torch.arange(3 * 8 * 8, dtype=torch.float32).view(3, 8, 8) % 256 + 0.5) / 255.0  # exact .5 ties: round half to even
    out["ties"] = (ties, ties.flip(0), 0)
    dark = torch.rand(3, 8, 8, generator=g) * (0.9 / 255.0)       # quantises to 0/1 only: calculate_psnr's max_value = 1 rule
    out["dark_max_value_1"] = (dark, torch.rand(3, 8, 8, generator=g) * (1.4 / 255.0), 0)
    out["identical"] = (a, a.clone(), 0)
    return out


if __name__ == "__main__":
    tensor2img, calculate_psnr = load_reference_functions()
    rec = {}
    for name, (p, q, crop) in cases().items():
        ip, iq = tensor2img([p]), tensor2img([q])
        rec[name + ".pred"], rec[name + ".gt"] = p.numpy(), q.numpy()
        rec[name + ".img_pred"], rec[name + ".img_gt"] = ip, iq
        rec[name + ".crop"] = np.int64(crop)
        rec[name + ".psnr"] = np.float64(calculate_psnr(ip, iq, crop))
        print(name, ip.shape, ip.dtype, rec[name + ".psnr"])
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "psnr_cases.npz"), **rec)
