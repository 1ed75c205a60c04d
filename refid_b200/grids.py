"""Full-frame validation tiling on the GPU: the reference's `grids` mode (basicsr/models/twoImage_event_recurrent_model.py
:128-270) -- overlapping `val.crop_size` crops on an adaptive stride, optionally in `val.trans_num` of the 8 dihedral
orientations, concatenated on dim 0, run through the network (chunked by `val.max_minibatch`), un-oriented and
overlap-averaged back into the frame.  Crop and merge are single gather kernels (csrc/grids.cu); the placement list is
host arithmetic restated from :201-243.  The reference's methods only accept 4-D tensors (its `grids_voxel` cannot take
this network's (1,T,C,H,W) voxel); here every (1, ..., H, W) tensor is tiled over its trailing two dimensions.
"""
import ctypes
import math
import random

import torch

from . import _lib


def crop_positions(h, w, crop_size, trans_num=1, random_crop_num=0):
    """[(i, j, trans_idx)] in the reference's order; `random_crop_num` extra crops drawn like the reference (:234-240)."""
    num_row = (h - 1) // crop_size + 1
    num_col = (w - 1) // crop_size + 1
    step_j = crop_size if num_col == 1 else math.ceil((w - crop_size) / (num_col - 1) - 1e-8)
    step_i = crop_size if num_row == 1 else math.ceil((h - crop_size) / (num_row - 1) - 1e-8)
    idx = []
    i, last_i = 0, False
    while i < h and not last_i:
        j = 0
        if i + crop_size >= h:
            i, last_i = h - crop_size, True
        last_j = False
        while j < w and not last_j:
            if j + crop_size >= w:
                j, last_j = w - crop_size, True
            for t in range(trans_num):
                idx.append((i, j, t))
            j += step_j
        i += step_i
    for _ in range(random_crop_num):
        idx.append((random.randint(0, h - crop_size), random.randint(0, w - crop_size), random.randint(0, trans_num - 1)))
    return idx


def _bind():
    L = _lib.lib()
    if not getattr(L, "_grids_bound", False):
        a = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
             ctypes.c_void_p, ctypes.c_void_p]
        L.refid_grids_crop.argtypes = a
        L.refid_grids_merge.argtypes = a
        L._grids_bound = True
    return L


def _check(t):
    if not t.is_cuda:
        raise RuntimeError("refid_b200.grids runs on CUDA only; there is no CPU path")
    if t.shape[0] != 1:
        raise AssertionError("grids mode validates one frame at a time (b == 1)")  # reference: `assert b == 1`


def index_tensor(idx, device):
    return torch.tensor(idx, dtype=torch.int32, device=device).reshape(-1, 3).contiguous()


def crop(frame, idx, crop_size):
    """frame (1, ..., H, W) -> (len(idx), ..., crop_size, crop_size); idx from `crop_positions`."""
    _check(frame)
    f = frame.float().contiguous()
    H, W = f.shape[-2:]
    planes = f.numel() // (H * W)
    it = index_tensor(idx, f.device)
    out = torch.empty((len(idx),) + tuple(f.shape[1:-2]) + (crop_size, crop_size), dtype=torch.float32, device=f.device)
    with torch.cuda.device(f.device):
        _lib.check(_bind().refid_grids_crop(_lib.ptr(f), planes, H, W, _lib.ptr(it), len(idx), crop_size, _lib.ptr(out),
                                            ctypes.c_void_p(torch.cuda.current_stream(f.device).cuda_stream)), "refid_grids_crop")
    return out


def merge(parts, idx, h, w):
    """parts (len(idx), ..., cs, cs) -> (1, ..., h, w): overlap average of the un-oriented crops."""
    if not parts.is_cuda:
        raise RuntimeError("refid_b200.grids runs on CUDA only; there is no CPU path")
    p = parts.float().contiguous()
    cs = p.shape[-1]
    if p.shape[0] != len(idx) or p.shape[-2] != cs:
        raise ValueError("parts do not match the crop list")
    cover = torch.zeros(h, w, dtype=torch.bool)
    for i, j, _ in idx:
        cover[i:i + cs, j:j + cs] = True
    if not bool(cover.all()):
        raise ValueError("the crop list does not cover the frame")
    planes = p.numel() // (len(idx) * cs * cs)
    it = index_tensor(idx, p.device)
    out = torch.empty((1,) + tuple(p.shape[1:-2]) + (h, w), dtype=torch.float32, device=p.device)
    with torch.cuda.device(p.device):
        _lib.check(_bind().refid_grids_merge(_lib.ptr(p), planes, h, w, _lib.ptr(it), len(idx), cs, _lib.ptr(out),
                                             ctypes.c_void_p(torch.cuda.current_stream(p.device).cuda_stream)), "refid_grids_merge")
    return out
