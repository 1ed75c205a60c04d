"""`events_to_voxel_grid` with the reference's signature (basicsr/data/event_util.py:6-66) for event arrays that already live
on the GPU, and the sliding two-bin windows the dataset builds from it (basicsr/data/image_npy_dataset.py:175-188).
SURVEY.md 8f rank 4; kernels in csrc/voxel.cu, C entry `refid_events_to_voxel`.  No CPU path.
"""
import ctypes

import torch

from . import _lib


def events_to_voxel_grid(events, num_bins, width, height, return_format="CHW"):
    """events: CUDA float32 tensor [N,4], rows [timestamp, x, y, polarity] -> voxel grid (num_bins,H,W) or (H,W,num_bins)."""
    if not (torch.is_tensor(events) and events.is_cuda):
        raise RuntimeError("refid_b200.event_util needs a CUDA tensor of events (no CPU path)")
    assert events.dim() == 2 and events.shape[1] == 4
    assert num_bins > 0 and width > 0 and height > 0
    if return_format not in ("CHW", "HWC"):
        raise ValueError(f"return_format {return_format}")
    ev = events.detach().float().contiguous()
    L = _lib.lib()
    L.refid_events_to_voxel.argtypes = [ctypes.c_void_p, ctypes.c_long, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                        ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    hwc = return_format == "HWC"
    out = torch.empty((height, width, num_bins) if hwc else (num_bins, height, width), dtype=torch.float32, device=ev.device)
    scratch = torch.empty(num_bins * height * width, dtype=torch.int64, device=ev.device)
    st = ctypes.c_void_p(torch.cuda.current_stream(ev.device).cuda_stream)
    with torch.cuda.device(ev.device):
        _lib.check(L.refid_events_to_voxel(_lib.ptr(ev), ev.shape[0], num_bins, width, height, 1 if hwc else 0, _lib.ptr(scratch),
                                           _lib.ptr(out), st), "refid_events_to_voxel")
    return out


def sliding_two_bin_voxels(event_chunks, width, height):
    """One 2-bin (H,W,2) voxel per consecutive pair of event chunks (image_npy_dataset.py:175-188)."""
    return [events_to_voxel_grid(torch.cat((a, b), dim=0), 2, width, height, "HWC")
            for a, b in zip(event_chunks[:-1], event_chunks[1:])]
