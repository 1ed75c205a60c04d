"""CharbonnierLoss with the reference's interface (basicsr/models/losses/losses.py:143-173; the function it wraps is
`sqrt((pred - target)**2 + eps)`, :28-30, reduced by `weighted_loss`), computed by ONE CUDA pass that produces the loss
value and d loss / d pred together (csrc/loss.cu, C entry `refid_charbonnier`).  SURVEY.md 8f rank 1.

The reference's wrapper builds it as `getattr(loss_module, train_opt['pixel_opt'].pop('type'))(**pixel_opt)`
(twoImage_event_recurrent_model.py:52-58) and calls `self.cri_pix(pred, self.gt)` (:284); this class accepts the same
keyword arguments and call.  Element-wise `weight` and `reduction='none'` are not used by any shipped option file and raise
NotImplementedError here (there is no torch / CPU fallback in this package).
"""
import ctypes

import torch
from torch import nn

from . import _lib

_reduction_modes = ["none", "mean", "sum"]


class _CharbonnierFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, target, eps, loss_weight, mean):
        if not pred.is_cuda:
            raise RuntimeError("refid_b200.losses.CharbonnierLoss needs CUDA tensors (no CPU path)")
        p = pred.detach().float().contiguous()
        t = target.detach().float().contiguous()
        if p.shape != t.shape:
            raise ValueError(f"pred {tuple(p.shape)} and target {tuple(t.shape)} differ in shape")
        L = _lib.lib()
        L.refid_charbonnier.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                        ctypes.c_long, ctypes.c_float, ctypes.c_float, ctypes.c_int, ctypes.c_void_p]
        need_grad = ctx.needs_input_grad[0]
        grad = torch.empty_like(p) if need_grad else None
        scratch = torch.empty(L.refid_charbonnier_scratch_bytes(), dtype=torch.uint8, device=p.device)
        loss = torch.empty((), dtype=torch.float32, device=p.device)
        st = ctypes.c_void_p(torch.cuda.current_stream(p.device).cuda_stream)
        with torch.cuda.device(p.device):
            _lib.check(L.refid_charbonnier(_lib.ptr(p), _lib.ptr(t), _lib.ptr(grad), _lib.ptr(scratch), _lib.ptr(loss),
                                           p.numel(), float(eps), float(loss_weight), 1 if mean else 0, st), "refid_charbonnier")
        ctx.grad = grad
        ctx.in_dtype = pred.dtype
        return loss

    @staticmethod
    def backward(ctx, go):
        g = ctx.grad
        if g is None:
            return None, None, None, None, None
        # d loss / d pred was produced by the forward pass; the upstream gradient (the scalar 1 for a training loss) is
        # applied on the device, and costs no memory traffic when it is exactly 1
        L = _lib.lib()
        L.refid_scale_by_device_scalar.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_long, ctypes.c_void_p]
        s = go.detach().float().reshape(1).contiguous()
        st = ctypes.c_void_p(torch.cuda.current_stream(g.device).cuda_stream)
        with torch.cuda.device(g.device):
            _lib.check(L.refid_scale_by_device_scalar(_lib.ptr(g), _lib.ptr(s), g.numel(), st), "refid_scale_by_device_scalar")
        ctx.grad = None
        return g.to(ctx.in_dtype), None, None, None, None


class CharbonnierLoss(nn.Module):
    def __init__(self, loss_weight=1.0, reduction="mean", eps=1e-12):
        super().__init__()
        if reduction not in ["none", "mean", "sum"]:
            raise ValueError(f"Unsupported reduction mode: {reduction}. Supported ones are: {_reduction_modes}")
        self.loss_weight = loss_weight
        self.reduction = reduction
        self.eps = eps

    def forward(self, pred, target, weight=None, **kwargs):
        if weight is not None or self.reduction == "none":
            raise NotImplementedError("refid_b200 CharbonnierLoss: element-wise weights / reduction='none' are not on the hot path")
        return _CharbonnierFn.apply(pred, target, self.eps, self.loss_weight, self.reduction == "mean")
