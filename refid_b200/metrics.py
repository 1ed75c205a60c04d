"""Validation metrics of the reference on the GPU.

The reference's validation loop converts every output and ground-truth frame on the host -- `tensor2img` (clamp, x255,
round, uint8, RGB->BGR, basicsr/utils/img_util.py:59-121) -- and then calls `calculate_psnr(sr_img, gt_img, crop_border)`
(basicsr/metrics/psnr_ssim.py:9-61).  Here one CUDA pass quantises both tensors and accumulates the squared uint8
differences as exact integers (csrc/metrics.cu, C entry `refid_quant_psnr`); the host only forms
`20 log10(255 / sqrt(mse))` in double, so the values are bit-identical to the reference's.

  * `psnr_frames(pred, gt, crop_border)`  == [calculate_psnr(tensor2img(p), tensor2img(g), crop_border) for p, g in frames]
  * `tensor2img(tensor)`                  == the reference's function for 3-D / (1,3,H,W) CUDA tensors, uint8, min_max (0,1)
"""
import ctypes
import math

import numpy as np
import torch

from . import _lib


def _frames(t):
    if not torch.is_tensor(t):
        raise TypeError(f"tensor expected, got {type(t)}")
    if not t.is_cuda:
        raise RuntimeError("refid_b200.metrics needs CUDA tensors (no CPU path)")
    t = t.detach().float()
    if t.dim() == 3:
        t = t.unsqueeze(0)
    elif t.dim() > 4:
        t = t.reshape(-1, *t.shape[-3:])
    if t.dim() != 4:
        raise TypeError(f"Only support (..., C, H, W) tensors. But received with dimension: {t.dim()}")
    return t.contiguous()


def _run(pred, gt, crop_border, reverse, want_images):
    L = _lib.lib()
    L.refid_quant_psnr.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                   ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                   ctypes.c_void_p, ctypes.c_void_p]
    F, C, H, W = pred.shape
    dev = pred.device
    ssd = torch.empty(F, dtype=torch.int64, device=dev) if gt is not None else None
    mx = torch.empty(F, dtype=torch.int32, device=dev) if gt is not None else None
    ip = torch.empty((F, H, W, C), dtype=torch.uint8, device=dev) if want_images else None
    ig = torch.empty((F, H, W, C), dtype=torch.uint8, device=dev) if (want_images and gt is not None) else None
    st = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    with torch.cuda.device(dev):
        _lib.check(L.refid_quant_psnr(_lib.ptr(pred), _lib.ptr(gt), F, C, H, W, int(crop_border), 1 if reverse else 0,
                                      _lib.ptr(ssd), _lib.ptr(mx), _lib.ptr(ip), _lib.ptr(ig), st), "refid_quant_psnr")
    return ssd, mx, ip, ig


def psnr_frames(pred, gt, crop_border=0):
    """PSNR of every (C,H,W) frame of `pred` against `gt` (any leading dimensions), as the reference computes it from the
    uint8 images: a list of Python floats, `inf` for identical frames."""
    p, g = _frames(pred), _frames(gt)
    if p.shape != g.shape:
        raise AssertionError(f"Image shapes are differnet: {tuple(p.shape)}, {tuple(g.shape)}.")
    ssd, mx, _, _ = _run(p, g, crop_border, False, False)
    F, C, H, W = p.shape
    count = (H - 2 * crop_border) * (W - 2 * crop_border) * C
    out = []
    for s, m in zip(ssd.cpu().tolist(), mx.cpu().tolist()):
        if s == 0:
            out.append(float("inf"))
            continue
        mse = np.float64(s) / np.float64(count)
        max_value = 1.0 if m <= 1 else 255.0  # psnr_ssim.py:60
        out.append(float(20.0 * np.log10(max_value / np.sqrt(mse))))
    return out


def tensor2img(tensor, rgb2bgr=True, out_type=np.uint8, min_max=(0, 1)):
    """The reference's `tensor2img` for a (3|1,H,W) or (1,3|1,H,W) CUDA tensor: (H,W,C) BGR uint8 ndarray ((H,W) for one
    channel), quantised on the GPU."""
    if isinstance(tensor, list):
        return [tensor2img(t, rgb2bgr, out_type, min_max) for t in tensor]
    if out_type != np.uint8 or tuple(min_max) != (0, 1):
        raise NotImplementedError("refid_b200 tensor2img: uint8 output with min_max=(0, 1) (what the validation loop uses)")
    t = tensor.squeeze(0) if torch.is_tensor(tensor) and tensor.dim() == 4 and tensor.size(0) == 1 else tensor
    if torch.is_tensor(t) and t.dim() != 3:
        raise NotImplementedError("refid_b200 tensor2img: one (C,H,W) frame per call (no make_grid)")
    p = _frames(t)
    _, _, ip, _ = _run(p, None, 0, bool(rgb2bgr), True)
    img = ip[0].cpu().numpy()
    return np.squeeze(img, axis=2) if img.shape[2] == 1 else img
