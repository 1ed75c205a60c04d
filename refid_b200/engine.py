"""ctypes binding of the network engine in librefid_b200.so (C ABI: include/refid_b200.h).

PyTorch only supplies device memory and the CUDA stream; every kernel is launched by the library.
"""
import ctypes

import torch

from . import _lib

c_void_p, c_int, c_long, c_size_t = ctypes.c_void_p, ctypes.c_int, ctypes.c_long, ctypes.c_size_t


class _Cfg(ctypes.Structure):
    _fields_ = [("img_chn", c_int), ("ev_chn", c_int), ("out_chn", c_int), ("base_num_channels", c_int)]


class _Entry(ctypes.Structure):
    _fields_ = [("key", ctypes.c_char * 96), ("kind", c_int), ("taps", c_int), ("R", c_int), ("Cc", c_int),
                ("nbias", c_int), ("w_off", c_long), ("b_off", c_long)]


KIND_CONV3, KIND_CONV1, KIND_DOWN4, KIND_UP2, KIND_ROWS5, KIND_RAW = range(6)


class FlatEntry(ctypes.Structure):
    _fields_ = [("ptr", c_void_p), ("flat_off", c_long), ("mode", c_int), ("taps", c_int), ("R", c_int), ("Cc", c_int)]


def flat_gather(table, n, flat):
    L = _bind(_lib.lib())
    st = c_void_p(torch.cuda.current_stream(flat.device).cuda_stream)
    _lib.check(L.refid_flat_gather(table, n, _lib.ptr(flat), flat.numel(), st), "refid_flat_gather")


def flat_scatter(table, n, gflat):
    L = _bind(_lib.lib())
    st = c_void_p(torch.cuda.current_stream(gflat.device).cuda_stream)
    _lib.check(L.refid_flat_scatter(table, n, _lib.ptr(gflat), st), "refid_flat_scatter")


def _bind(L):
    if getattr(L, "_refid_bound", False):
        return L
    L.refid_create.argtypes = [ctypes.POINTER(_Cfg), ctypes.POINTER(c_void_p)]
    L.refid_destroy.argtypes = [c_void_p]
    L.refid_num_param_entries.argtypes = [c_void_p]
    L.refid_param_entry_at.argtypes = [c_void_p, c_int, ctypes.POINTER(_Entry)]
    L.refid_flat_floats.argtypes = [c_void_p]
    L.refid_flat_floats.restype = c_long
    L.refid_wpack_bytes.argtypes = [c_void_p]
    L.refid_wpack_bytes.restype = c_size_t
    L.refid_workspace_bytes.argtypes = [c_void_p, c_int, c_int, c_int, c_int, c_int, ctypes.POINTER(c_size_t)]
    L.refid_plan.argtypes = [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]
    L.refid_pack_weights.argtypes = [c_void_p, c_void_p, c_void_p]
    L.refid_forward.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
    L.refid_backward.argtypes = [c_void_p, c_void_p, c_void_p]
    L.refid_profile.argtypes = [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p]
    L.refid_profile_csv.argtypes = [c_void_p, c_int, ctypes.c_char_p, c_void_p]
    L.refid_num_launches.argtypes = [c_void_p, ctypes.POINTER(c_int), ctypes.POINTER(c_int)]
    L.refid_flat_gather.argtypes = [ctypes.POINTER(FlatEntry), c_int, c_void_p, c_long, c_void_p]
    L.refid_flat_scatter.argtypes = [ctypes.POINTER(FlatEntry), c_int, c_void_p, c_void_p]
    L.refid_set_option.argtypes = [c_void_p, ctypes.c_char_p, c_long]
    L.refid_graph_stats.argtypes = [c_void_p, ctypes.POINTER(c_long * 4)]
    L.refid_plan_storage.argtypes = [c_void_p]
    L.refid_graph_error.argtypes = [c_void_p]
    L.refid_graph_error.restype = ctypes.c_char_p
    L.refid_debug_tensor.argtypes = [c_void_p, ctypes.c_char_p, ctypes.POINTER(c_void_p)] + [ctypes.POINTER(c_int)] * 5
    L._refid_bound = True
    return L


class Engine:
    """One engine = one network configuration; `plan` binds it to a problem size and to fixed device buffers."""

    def __init__(self, img_chn, ev_chn, out_chn=3, base_num_channels=32):
        self.L = _bind(_lib.lib())
        self.h = c_void_p()
        cfg = _Cfg(img_chn, ev_chn, out_chn, base_num_channels)
        _lib.check(self.L.refid_create(ctypes.byref(cfg), ctypes.byref(self.h)), "refid_create")
        self.entries = []
        for i in range(self.L.refid_num_param_entries(self.h)):
            e = _Entry()
            _lib.check(self.L.refid_param_entry_at(self.h, i, ctypes.byref(e)), "refid_param_entry_at")
            self.entries.append({"key": e.key.decode(), "kind": e.kind, "taps": e.taps, "R": e.R, "Cc": e.Cc,
                                 "nbias": e.nbias, "w_off": e.w_off, "b_off": e.b_off})
        self.flat_floats = self.L.refid_flat_floats(self.h)
        self.wpack_bytes = self.L.refid_wpack_bytes(self.h)
        self.shape = None

    def __del__(self):
        try:
            if self.h:
                self.L.refid_destroy(self.h)
                self.h = c_void_p()
        except Exception:  # interpreter shutdown
            pass

    def set_option(self, name, value):
        """Engine option (include/refid_b200.h): "infer_fp16" (before plan), "graphs"."""
        _lib.check(self.L.refid_set_option(self.h, name.encode(), int(value)), "refid_set_option")

    def graph_stats(self):
        a = (c_long * 4)()
        self.L.refid_graph_stats(self.h, ctypes.byref(a))
        d = {"captures": a[0], "replays": a[1], "eager": a[2], "failures": a[3]}
        if a[3]:
            d["error"] = self.L.refid_graph_error(self.h).decode()
        return d

    def storage(self):
        """16-bit storage type of the current plan's activations / packed weights."""
        return torch.float16 if self.L.refid_plan_storage(self.h) else torch.bfloat16

    def workspace_bytes(self, B, T, H, W, train):
        n = c_size_t(0)
        _lib.check(self.L.refid_workspace_bytes(self.h, B, T, H, W, int(train), ctypes.byref(n)), "refid_workspace_bytes")
        return n.value

    def plan(self, B, T, H, W, train, workspace, wpack, grad_flat):
        with torch.cuda.device(workspace.device):
            self._plan(B, T, H, W, train, workspace, wpack, grad_flat)

    def _plan(self, B, T, H, W, train, workspace, wpack, grad_flat):
        _lib.check(self.L.refid_plan(self.h, B, T, H, W, int(train), _lib.ptr(workspace), _lib.ptr(wpack),
                                     _lib.ptr(grad_flat)), "refid_plan")
        self.shape = (B, T, H, W, bool(train))
        self._keep = (workspace, wpack, grad_flat)  # the plan holds raw pointers into these

    def _stream(self):
        # the stream of the device the plan's buffers live on (not of whatever device happens to be current)
        return c_void_p(torch.cuda.current_stream(self._keep[0].device).cuda_stream)

    def pack_weights(self, flat):
        assert flat.dtype == torch.float32 and flat.is_contiguous() and flat.numel() == self.flat_floats
        with torch.cuda.device(flat.device):
            self._pack_weights(flat)

    def _pack_weights(self, flat):
        _lib.check(self.L.refid_pack_weights(self.h, _lib.ptr(flat), self._stream()), "refid_pack_weights")

    def forward(self, x, event, out):
        for t in (x, event, out):
            assert t.dtype == torch.float32 and t.is_contiguous() and t.is_cuda
        with torch.cuda.device(x.device):
            _lib.check(self.L.refid_forward(self.h, _lib.ptr(x), _lib.ptr(event), _lib.ptr(out), self._stream()), "refid_forward")

    def backward(self, grad_out):
        assert grad_out.dtype == torch.float32 and grad_out.is_contiguous() and grad_out.is_cuda
        with torch.cuda.device(grad_out.device):
            _lib.check(self.L.refid_backward(self.h, _lib.ptr(grad_out), self._stream()), "refid_backward")

    def profile(self, with_backward=True):
        """Per-class device time / algorithmic FLOPs / launch count of one forward(+backward) replay (see header)."""
        ms, fl, n = (ctypes.c_double * 6)(), (ctypes.c_double * 6)(), (ctypes.c_long * 6)()
        _lib.check(self.L.refid_profile(self.h, int(with_backward), ms, fl, n, self._stream()), "refid_profile")
        names = ("conv_other_fwd", "conv_other_dgrad", "wgrad", "other", "conv3x3_fwd", "conv3x3_dgrad")
        return {k: {"ms": ms[i], "flops": fl[i], "launches": n[i]} for i, k in enumerate(names)}

    def profile_csv(self, path, with_backward=True):
        _lib.check(self.L.refid_profile_csv(self.h, int(with_backward), str(path).encode(), self._stream()), "refid_profile_csv")

    def num_launches(self):
        a, b = c_int(0), c_int(0)
        self.L.refid_num_launches(self.h, ctypes.byref(a), ctypes.byref(b))
        return a.value, b.value

    def debug_tensor(self, name):
        """Named intermediate activation as an (N,C,H,W) fp32 tensor (copy)."""
        p = c_void_p()
        d = [c_int(0) for _ in range(5)]
        _lib.check(self.L.refid_debug_tensor(self.h, name.encode(), ctypes.byref(p), *[ctypes.byref(v) for v in d]),
                   "refid_debug_tensor")
        N, H, W, C, pitch = [v.value for v in d]
        ws = self._keep[0]
        off = p.value - ws.data_ptr()
        n = ((N * H * W - 1) * pitch + C) * 2
        raw = ws[off:off + n].view(self.storage())
        t = torch.as_strided(raw, (N, H, W, C), (H * W * pitch, W * pitch, pitch, 1))
        return t.float().permute(0, 3, 1, 2).contiguous()
