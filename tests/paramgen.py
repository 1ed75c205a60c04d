"""Deterministic, construction-order-independent parameters and inputs shared by the golden generator
(run on the reference module) and the parity tests (run on the oracle and on the CUDA path)."""
import zlib

import torch


def make_params(shapes, seed=0, dtype=torch.float32):
    """name -> tensor. Scales keep activations O(1) over T recurrent steps; beta/gamma are NON-zero
    (the reference zero-initialises them, which would hide EGACA from every test)."""
    P = {}
    for name in sorted(shapes):
        shp = shapes[name]
        g = torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(name.encode())) % (2 ** 31))
        r = torch.randn(shp, generator=g, dtype=torch.float32)
        if name.endswith(".beta") or name.endswith(".gamma"):
            t = 0.3 * r
        elif ".norm" in name and name.endswith(".weight"):
            t = 1.0 + 0.1 * r
        elif name.endswith(".bias"):
            t = 0.05 * r
        else:
            fan_in = shp[1] * shp[2] * shp[3]
            if "transposed_conv2d" in name:
                fan_in = shp[0]  # each output pixel sees Cin inputs (k=2,s=2)
            gain = 1.0
            if ".main.2.0." in name or "resblocks" in name:
                gain = 0.5
            t = gain * r / fan_in ** 0.5
        P[name] = t.to(dtype)
    return P


def make_inputs(B, T, H, W, img_chn, ev_chn, seed=1234, x5d=False):
    """Synthetic inputs with the distributions of SURVEY.md 8(d): images U[0,1), voxels 80% zeros else N(0,1)."""
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(B, img_chn, H, W, generator=g)
    if img_chn == 26:  # channels 3-12 and 16-25 are voxel bins (image_npy_dataset.py:212-221)
        vox = torch.randn(B, img_chn, H, W, generator=g) * (torch.rand(B, img_chn, H, W, generator=g) > 0.8)
        idx = list(range(3, 13)) + list(range(16, 26))
        x[:, idx] = vox[:, idx]
    ev = torch.randn(B, T, ev_chn, H, W, generator=g) * (torch.rand(B, T, ev_chn, H, W, generator=g) > 0.8)
    gt = torch.rand(B, T, 3, H, W, generator=g)
    if x5d:
        x = x.view(B, 2, img_chn // 2, H, W)
    return x, ev, gt


def grad_sample_index(name, numel, k=32):
    g = torch.Generator().manual_seed(zlib.crc32(("idx" + name).encode()) % (2 ** 31))
    return torch.randint(0, numel, (min(k, numel),), generator=g)
