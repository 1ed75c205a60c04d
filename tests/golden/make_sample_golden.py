"""Golden vectors for the training-sample assembly, from the UNMODIFIED reference functions (build container only):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_sample_golden.py

Calls `triple_random_crop` and `augment` of basicsr/data/transforms.py and `img2tensor` of basicsr/utils/img_util.py
(imported through stub packages) on synthetic frames with Python's `random` seeded, then applies the inline steps of
`__getitem__` that follow (image_npy_dataset.py:196-232, quoted verbatim below: they are statements inside a method that
needs image / event files on disk).  The random draws are recovered by replaying the same seed.
-> tests/golden/sample_pack_cases.npz
"""
import importlib
import os
import random
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = os.environ.get("REFID_REFERENCE", "/root/reference")


def load():
    sys.dont_write_bytecode = True
    for name, path in (("basicsr", f"{REF}/basicsr"), ("basicsr.utils", f"{REF}/basicsr/utils"), ("basicsr.data", f"{REF}/basicsr/data")):
        mod = types.ModuleType(name)
        mod.__path__ = [path]
        sys.modules[name] = mod
    return importlib.import_module("basicsr.data.transforms"), importlib.import_module("basicsr.utils.img_util")


CASES = {  # name: (H, W, m, n, gt_size, seed)
    "gopro_11p1_crop": (40, 56, 11, 1, 32, 4),        # hflip + vflip + transpose
    "gopro_11p3_crop": (48, 48, 11, 3, 24, 7),        # hflip + vflip
    "small_m_full_frame": (24, 32, 3, 2, None, 3),    # no crop, non-square: hflip + transpose -> (W,H) planes
    "vflip_transpose": (40, 40, 4, 1, 16, 18),
    "plain": (40, 56, 11, 1, 32, 13),
}

if __name__ == "__main__":
    T, U = load()
    rec = {}
    for name, (H, W, m, n, gt_size, seed) in CASES.items():
        rng = np.random.RandomState(seed)
        nb = 2 * m + n + 1
        img_lqs = [rng.rand(H, W, 3).astype(np.float32) for _ in range(2)]
        img_gts = [rng.rand(H, W, 3).astype(np.float32) for _ in range(2 * m + n)]
        voxels = [rng.randn(H, W, nb).astype(np.float32)]
        rec[name + ".lqs"], rec[name + ".gts"], rec[name + ".voxel"] = np.stack(img_lqs), np.stack(img_gts), voxels[0]
        random.seed(seed)
        if gt_size is not None:
            img_gts, img_lqs, voxels = T.triple_random_crop(img_gts, img_lqs, voxels, gt_size, 1, "gt")
            if not isinstance(voxels, list):
                voxels = [voxels]
        # ---- image_npy_dataset.py:196-205 ----
        num_lq, num_gt = len(img_lqs), len(img_gts)
        img_lqs.extend(img_gts)
        img_lqs.extend(voxels)
        img_results = T.augment(img_lqs, True, True)
        img_results = U.img2tensor(img_results)
        lqs = torch.stack(img_results[:num_lq], dim=0)
        gts = torch.stack(img_results[num_lq:num_lq + num_gt], dim=0)
        voxels_list = img_results[num_lq + num_gt:]
        # ---- :209-221 (return_deblur_voxel) ----
        left_deblur_voxel = voxels_list[0][1:m, :, :]
        right_deblur_voxel = voxels_list[0][m + 2 + n:, :, :]
        lq = torch.cat((lqs[0], left_deblur_voxel, lqs[1], right_deblur_voxel), dim=0)
        # ---- :223-232 (one_voxel_flg) ----
        vox = torch.stack(voxels_list, dim=0).squeeze(0)
        vox = torch.stack([vox[i:i + 2, :, :] for i in range(vox.shape[0] - 1)], dim=0)
        # the draws, replayed
        random.seed(seed)
        top = left = 0
        if gt_size is not None:
            top, left = random.randint(0, H - gt_size), random.randint(0, W - gt_size)
        flags = [random.random() < 0.5 for _ in range(3)]
        rec[name + ".cfg"] = np.array([m, n, -1 if gt_size is None else gt_size, seed, top, left] + [int(f) for f in flags], dtype=np.int64)
        rec[name + ".lq"], rec[name + ".vox"], rec[name + ".gt"] = lq.numpy(), vox.numpy(), gts.numpy()
        print(name, tuple(lq.shape), tuple(vox.shape), tuple(gts.shape), "top/left", top, left, "flags", flags)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "sample_pack_cases.npz"), **rec)
