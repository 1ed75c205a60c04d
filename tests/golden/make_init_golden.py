"""Fingerprint of the UNMODIFIED reference's default initialisation (run in the build container only):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_init_golden.py

`torch.manual_seed(0)` then `FinalBidirectionAttenfusion(img_chn=26, ev_chn=2, num_encoders=3, base_num_channels=32,
num_block=1, num_residual_blocks=2)` of the reference; stores, per parameter in state_dict order, the float64 sum, the
float64 sum of squares and the first four elements -> tests/golden/init_seed0_img26.npz.  The drop-in module must consume
the RNG in the same order (same constructors, same order) to reproduce them.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402


def fingerprint(sd):
    names = list(sd)
    return {"names": np.array(names),
            "sum": np.array([sd[k].double().sum().item() for k in names]),
            "sumsq": np.array([sd[k].double().pow(2).sum().item() for k in names]),
            "first4": np.stack([np.pad(sd[k].flatten()[:4].numpy(), (0, max(0, 4 - sd[k].numel()))) for k in names])}


if __name__ == "__main__":
    torch.manual_seed(0)
    net = ref_loader.build(26, 2)
    fp = fingerprint(net.state_dict())
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "init_seed0_img26.npz"), **fp)
    print(len(fp["names"]), "tensors; source:", ref_loader.source())
