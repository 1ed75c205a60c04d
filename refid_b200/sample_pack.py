"""Training-sample assembly on the GPU: the step of the reference's dataset between the voxel grid
and the network's inputs -- basicsr/data/image_npy_dataset.py:189-232: `triple_random_crop` (transforms.py:163-238),
`augment` (flip / flip / transpose, transforms.py:88-129), the "deblur voxel" channel packing of `lq` (:209-221) and the
sliding two-bin windows of `voxel` (:226-232) -- as three gather launches (csrc/samplepack.cu) on tensors that already
live on the device (e.g. the output of refid_b200.event_util.events_to_voxel_grid).  No CPU path.
"""
import ctypes
import random

import torch

from . import _lib


def draw_crop_and_flips(h, w, gt_size, use_hflip=True, use_rot=True):
    """The reference's random draws, in its order: top, left (transforms.py:212-213), then hflip, vflip, rot90
    (transforms.py:110-112), from Python's `random` -- seeding `random` reproduces the reference's choices."""
    top = left = 0
    if gt_size is not None:
        top = random.randint(0, h - gt_size)
        left = random.randint(0, w - gt_size)
    hflip = bool(use_hflip and random.random() < 0.5)
    vflip = bool(use_rot and random.random() < 0.5)
    rot90 = bool(use_rot and random.random() < 0.5)
    return top, left, hflip, vflip, rot90


def _gather(src_a, src_b, table, top, left, ph, pw, hflip, vflip, rot90):
    L = _lib.lib()
    L.refid_crop_flip_gather.argtypes = [ctypes.c_void_p, ctypes.c_void_p] + [ctypes.c_int] * 2 + [ctypes.c_void_p] + \
        [ctypes.c_int] * 8 + [ctypes.c_void_p, ctypes.c_void_p]
    H, W = src_a.shape[-2:]
    dev = src_a.device
    tab = torch.tensor(table, dtype=torch.int32, device=dev)
    oh, ow = (pw, ph) if rot90 else (ph, pw)
    out = torch.empty(len(table), oh, ow, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(L.refid_crop_flip_gather(_lib.ptr(src_a), _lib.ptr(src_b), H, W, _lib.ptr(tab), len(table), top, left, ph, pw,
                                            int(hflip), int(vflip), int(rot90), _lib.ptr(out),
                                            ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), "refid_crop_flip_gather")
    return out


def pack_blurry_sample(img_lqs, voxel, img_gts, m, n, gt_size=None, top=0, left=0, hflip=False, vflip=False, rot90=False,
                       return_deblur_voxel=True):
    """img_lqs (2,3,H,W) RGB, voxel (num_bins,H,W) with num_bins = 2m+n+1, img_gts (2m+n,3,H,W); all CUDA fp32.
    Returns {'lq': (6+2(m-1),h,w) [or (2,3,h,w) without the deblur voxel], 'voxel': (num_bins-1,2,h,w), 'gt': (2m+n,3,h,w)}
    as `__getitem__` builds them (image_npy_dataset.py:189-244)."""
    for t in (img_lqs, voxel, img_gts):
        if not (torch.is_tensor(t) and t.is_cuda):
            raise RuntimeError("refid_b200.sample_pack needs CUDA tensors (no CPU path)")
    img_lqs, voxel, img_gts = img_lqs.float().contiguous(), voxel.float().contiguous(), img_gts.float().contiguous()
    nb, H, W = voxel.shape
    assert img_lqs.shape == (2, 3, H, W) and img_gts.shape[1:] == (3, H, W)
    assert nb == 2 * m + n + 1, f"voxel has {nb} bins, expected 2m+n+1 = {2 * m + n + 1}"
    ph, pw = (gt_size, gt_size) if gt_size is not None else (H, W)
    g = (top, left, ph, pw, hflip, vflip, rot90)
    frames = img_lqs.view(6, H, W)
    if return_deblur_voxel:  # lq = cat(left_lq, voxel[1:m], right_lq, voxel[m+2+n:])   (:209-221); voxel planes are -1-v
        table = [0, 1, 2] + [-1 - v for v in range(1, m)] + [3, 4, 5] + [-1 - v for v in range(m + 2 + n, nb)]
        lq = _gather(frames, voxel, table, *g)
    else:
        lq = _gather(frames, voxel, list(range(6)), *g)
        lq = lq.view(2, 3, lq.shape[-2], lq.shape[-1])
    win = [v for t in range(nb - 1) for v in (t, t + 1)]  # voxels[i:i+2] for i in range(num_bins-1)   (:226-232)
    vox = _gather(voxel, None, win, *g)
    vox = vox.view(nb - 1, 2, vox.shape[-2], vox.shape[-1])
    gts = img_gts.view(-1, H, W)
    gt = _gather(gts, None, list(range(gts.shape[0])), *g)
    gt = gt.view(-1, 3, gt.shape[-2], gt.shape[-1])
    return {"lq": lq, "voxel": vox, "gt": gt}
