"""Fused global-norm clip + AdamW / Adam step: the two calls that close the reference's
`optimize_parameters` -- `torch.nn.utils.clip_grad_norm_(self.net_g.parameters(), 0.01)` and `self.optimizer_g.step()`
(basicsr/models/twoImage_event_recurrent_model.py:304-307; optimizer built at :67-95 from `train.optim_g`) -- as two
multi-tensor CUDA passes over all parameters (csrc/optim.cu, C entries `refid_optim_*`).

`ClipAdamW` / `ClipAdam` take torch.optim.AdamW's / Adam's constructor arguments and keep torch's per-parameter state
layout (`step`, `exp_avg`, `exp_avg_sq`), so `state_dict()` / `load_state_dict()` interchange with the reference's
checkpoints.  `clip_grad_norm_(max_norm)` records the clip for the next `step()` (the norm needs every gradient, so it is
computed by the step's first pass); the clipped gradients are applied, not written back to `.grad`.
"""
import ctypes

import torch

from . import _lib


class _ClipAdamBase(torch.optim.Optimizer):
    _decoupled = True

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, amsgrad=False):
        if amsgrad:
            raise NotImplementedError("amsgrad is not used by the reference's option files")
        if lr < 0 or eps < 0 or not 0 <= betas[0] < 1 or not 0 <= betas[1] < 1 or weight_decay < 0:
            raise ValueError("invalid optimizer hyper-parameter")
        super().__init__(params, dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay))
        self._max_norm = 0.0
        self._handle = None
        self._key = None
        self._norm = None
        self._tables = None
        self._steps = None
        self._step_params = None

    def clip_grad_norm_(self, max_norm):
        """Global 2-norm clip of all gradients to `max_norm`, fused into the next step()."""
        self._max_norm = float(max_norm)

    def last_grad_norm(self):
        """(total gradient norm, clip coefficient) of the last step, as a device tensor of two floats."""
        return self._norm

    def _sync_steps(self):
        """Host step counts -> the `step` tensors of torch's state layout (before the state is exported or re-read)."""
        if getattr(self, "_tables", None) is not None and self._step_params is not None:
            for p, k in zip(self._step_params, self._steps):
                if p in self.state and "step" in self.state[p]:
                    self.state[p]["step"].fill_(float(k))

    def state_dict(self):
        self._sync_steps()
        return super().state_dict()

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._tables = None  # the moment tensors and step counts were replaced

    def __del__(self):
        try:
            if self._handle is not None:
                _lib.lib().refid_optim_destroy(self._handle)
        except Exception:
            pass

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        # The reference builds two groups (`optim_params`, and `optim_params_lowlr` for names no parameter of this network
        # has -- empty, with its own lr; twoImage_event_recurrent_model.py:67-91).  Empty groups are carried through
        # state_dict()/load_state_dict() untouched so optimizer checkpoints interchange; the step runs on the non-empty one.
        groups = [g for g in self.param_groups if len(g["params"]) > 0]
        if len(groups) != 1:
            raise NotImplementedError("one non-empty parameter group (the reference's option files produce exactly one)")
        g = groups[0]
        params = [p for p in g["params"]]
        L = _lib.lib()
        n = len(params)
        key = tuple((p.data_ptr(), p.numel()) for p in params)
        if key != self._key:
            for p in params:
                if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous()):
                    raise RuntimeError("ClipAdamW needs contiguous fp32 CUDA parameters (no CPU path)")
            self._sync_steps()  # the tables are about to be rebuilt from the `step` tensors: make them current first
            if self._handle is not None:
                _lib.check(L.refid_optim_destroy(self._handle), "refid_optim_destroy")
            numel = (ctypes.c_long * n)(*[p.numel() for p in params])
            h = ctypes.c_void_p()
            with torch.cuda.device(params[0].device):
                _lib.check(L.refid_optim_create(n, numel, ctypes.byref(h)), "refid_optim_create")
            self._handle, self._key = h, key
            self._norm = torch.zeros(2, dtype=torch.float32, device=params[0].device)
            self._tables = None
        if self._tables is None or any(len(self.state[p]) == 0 for p in params):
            # torch's state layout; the host keeps the step counts as ints and writes them into the `step` tensors only
            # when the state is exported (183 scalar tensor updates per step would cost more host time than the kernels)
            for p in params:
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = torch.tensor(0.0)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            VP = ctypes.c_void_p * n
            self._steps = [int(self.state[p]["step"].item()) for p in params]
            self._step_params = list(params)
            self._tables = (VP, VP(*[p.data_ptr() for p in params]), VP(*[self.state[p]["exp_avg"].data_ptr() for p in params]),
                            VP(*[self.state[p]["exp_avg_sq"].data_ptr() for p in params]))
            # pointer-typed (not array-size-typed) prototype: optimizers over different tensor counts coexist
            L.refid_optim_step.argtypes = [ctypes.c_void_p] + [ctypes.c_void_p] * 4 + [ctypes.c_float] * 6 + \
                                          [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        VP, a_p, a_m, a_v = self._tables
        keep, ptrs, steps = [], [], self._steps
        for i, p in enumerate(params):
            gr = p.grad
            if gr is None:
                ptrs.append(None)
                continue
            if gr.dtype != torch.float32 or not gr.is_contiguous():
                gr = gr.detach().float().contiguous()
                keep.append(gr)
            ptrs.append(gr.data_ptr())
            steps[i] += 1
        if all(q is None for q in ptrs):
            return loss
        a_g = VP(*ptrs)
        a_steps = (ctypes.c_long * n)(*steps)
        stream = ctypes.c_void_p(torch.cuda.current_stream(params[0].device).cuda_stream)
        with torch.cuda.device(params[0].device):
            cast = lambda a: ctypes.cast(a, ctypes.c_void_p)
            _lib.check(L.refid_optim_step(self._handle, cast(a_p), cast(a_g), cast(a_m), cast(a_v), self._max_norm, g["lr"],
                                          g["betas"][0], g["betas"][1], g["eps"], g["weight_decay"], cast(a_steps),
                                          1 if self._decoupled else 0, _lib.ptr(self._norm), stream), "refid_optim_step")
        # the kernel wrote the parameters through raw pointers: tell autograd / the engine's packed-weight cache
        torch.autograd.graph.increment_version([p for p, q in zip(params, ptrs) if q is not None])
        self._max_norm = 0.0  # like the reference, the clip is requested before every step
        return loss


class ClipAdamW(_ClipAdamBase):
    """torch.optim.AdamW (decoupled weight decay) with the gradient clip fused in."""
    _decoupled = True

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, amsgrad=False):
        super().__init__(params, lr, betas, eps, weight_decay, amsgrad)


class ClipAdam(_ClipAdamBase):
    """torch.optim.Adam (L2 weight decay added to the gradient) with the gradient clip fused in."""
    _decoupled = False

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False):
        super().__init__(params, lr, betas, eps, weight_decay, amsgrad)
