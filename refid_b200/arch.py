"""`FinalBidirectionAttenfusion` on the B200-native engine.

Drop-in for the reference class of the same name (basicsr/models/archs/XXNet_final_attenfusion_arch.py:81-218):
same constructor signature, same `forward(x, event) -> (B,T,out_chn,H,W)` contract, and a parameter tree whose
`state_dict()` is key-for-key and shape-for-shape identical (183 tensors), so reference checkpoints load
with strict=True.  The modules below only HOLD parameters (constructed in the reference's order so that the same seed
gives the same initial weights); no torch.nn layer is ever called.  All arithmetic of the forward and backward pass runs
in librefid_b200.so (hand-written sm_100a kernels) through refid_b200.engine; there is no PyTorch or CPU fallback.

The only torch ops on the path are parameter plumbing: the parameters are gathered (differentiably) into the engine's
flat "gradient layout" vector, with three algebraic folds that remove elementwise passes from the hot loop:
  * LayerNorm2d affine (fusion_modules.py:125-134) folded into the 1x1 conv that follows it,
  * EGACA's `beta` folded into conv3 (fusion_modules.py:317-321) and `gamma` into conv5, which is then K-concatenated
    with conv_y_side (fusion_modules.py:325-333),
  * the level-0 in-convs of both directions stacked into one conv over all T event slices.
autograd maps the engine's flat gradient back through these folds to the named parameters, so DDP hooks fire as usual.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import engine as _engine


class _Holder(nn.Module):
    """Parameter container mirroring one reference sub-module (never called)."""

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("parameter holder; the network runs in librefid_b200.so")


def _conv_layer(cin, cout, k, pad):  # ConvLayer, norm=None (recurrent_sub_modules.py:52-84)
    h = _Holder()
    h.conv2d = nn.Conv2d(cin, cout, k, 1, pad)
    return h


def _trunk(cin, c):  # ConvResidualBlocks(num_block=1) (recurrent_sub_modules.py:710-758)
    first = nn.Conv2d(cin, c, 3, 1, 1)
    blk = _Holder()
    blk.conv1 = nn.Conv2d(c, c, 3, 1, 1)
    blk.conv2 = nn.Conv2d(c, c, 3, 1, 1)
    with torch.no_grad():  # default_init_weights([conv1, conv2], 0.1) (:752-753, :776-804)
        for m in (blk.conv1, blk.conv2):
            nn.init.kaiming_normal_(m.weight)
            m.weight.mul_(0.1)
            m.bias.zero_()
    t = _Holder()
    t.main = nn.Sequential(first, nn.Identity(), nn.Sequential(blk))
    return t


def _egaca(c, c_out):  # CrossmodalAtten_imgeventalladd (fusion_modules.py:237-288)
    a = _Holder()
    a.conv1 = nn.Conv2d(c, c, 1)
    a.conv2 = nn.Conv2d(c, c, 3, padding=1, groups=c)
    a.conv1_e = nn.Conv2d(c, c, 1)
    a.conv2_e = nn.Conv2d(c, c, 3, padding=1, groups=c)
    a.conv3 = nn.Conv2d(2 * c, c, 1)
    a.se_1 = nn.Sequential(nn.Identity(), nn.Conv2d(c, c // 2, 1), nn.Identity(), nn.Conv2d(c // 2, c, 1), nn.Identity())
    a.se_2 = nn.Sequential(nn.Identity(), nn.Conv2d(c, c // 2, 1), nn.Identity(), nn.Conv2d(c // 2, c, 1), nn.Identity())
    a.conv4 = nn.Conv2d(c, 2 * c, 1)
    a.conv5 = nn.Conv2d(2 * c, c_out, 1)
    a.conv_y_side = nn.Conv2d(c, c_out, 1)
    for n in ("norm1", "norm1_e", "norm2"):
        ln = _Holder()
        ln.weight = nn.Parameter(torch.ones(c))
        ln.bias = nn.Parameter(torch.zeros(c))
        setattr(a, n, ln)
    a.beta = nn.Parameter(torch.zeros(1, c, 1, 1))
    a.gamma = nn.Parameter(torch.zeros(1, c_out, 1, 1))
    return a


def _evr_layer(cin, c, fuse, atten):  # SimpleRecurrentThenDownAttenfusionmodifiedConvLayer (:245-268)
    l = _Holder()
    l.conv = _conv_layer(cin, c, 3, 1)
    if atten:
        l.atten_fuse = _egaca(cin, c)
    rb = _Holder()
    rb.forward_trunk = _trunk(2 * c, c)
    l.recurrent_block = rb
    if fuse:
        l.fuse_two_dir = _conv_layer(2 * c, c, 1, 0)
    l.down = nn.Conv2d(c, c, 4, 2, 1, bias=False)
    return l


def sync_flat_grad(g, group):
    """In-place mean of the flat gradient over the data-parallel group (one all-reduce; NCCL on GPUs, gloo in tests)."""
    import torch.distributed as dist
    if g.is_cuda:
        dist.all_reduce(g, op=dist.ReduceOp.AVG, group=group)  # NCCL averages in the reduction itself
    else:  # gloo has no AVG
        dist.all_reduce(g, op=dist.ReduceOp.SUM, group=group)
        g.div_(dist.get_world_size(group))
    return g


def broadcast_parameters(module, group, src=0):
    """Rank `src`'s parameters and buffers to every rank of `group` -- what DistributedDataParallel does at construction
    (reference base_model.py:66-72).  The reference seeds rank r with `manual_seed + r` (train.py:56), so without this
    the replicas would start from different weights and, applying the same averaged gradient, never meet."""
    import torch.distributed as dist
    ts = [p.data for p in module.parameters()] + [b.data for b in module.buffers()]
    if not ts:
        return
    flat = torch.cat([t.reshape(-1).float() for t in ts])
    dist.broadcast(flat, src=dist.get_global_rank(group, src) if group is not None else src, group=group)
    off = 0
    with torch.no_grad():
        for t in ts:
            t.copy_(flat[off:off + t.numel()].view_as(t))
            off += t.numel()


class _FlatParams(torch.autograd.Function):
    """Parameter tensors -> flat vector (refid_flat_gather) with the inverse scatter as its backward."""

    @staticmethod
    def _table(ent, ptrs):
        tab = (_engine.FlatEntry * len(ent))()
        for i, ((off, mode, taps, R, Cc), p) in enumerate(zip(ent, ptrs)):
            tab[i].ptr, tab[i].flat_off, tab[i].mode, tab[i].taps, tab[i].R, tab[i].Cc = p, off, mode, taps, R, Cc
        return tab

    @staticmethod
    def forward(ctx, flat_floats, ent, *ins):
        dev = ins[0].device
        if dev.type != "cuda":
            raise RuntimeError("refid_b200 runs on CUDA (sm_100a) only; there is no CPU path")
        srcs = [t.detach() if (t.is_contiguous() and t.dtype == torch.float32) else t.detach().float().contiguous() for t in ins]
        flat = torch.empty(flat_floats, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _engine.flat_gather(_FlatParams._table(ent, [t.data_ptr() for t in srcs]), len(ent), flat)
        ctx.ent, ctx.shapes, ctx.dev = ent, [tuple(t.shape) for t in ins], dev
        return flat

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous()
        sizes = [int(math.prod(s)) for s in ctx.shapes]
        offs = [0]
        for n in sizes:
            offs.append(offs[-1] + (n + 3) // 4 * 4)  # 16-byte aligned pieces of one buffer
        buf = torch.empty(offs[-1], dtype=torch.float32, device=ctx.dev)
        base = buf.data_ptr()
        with torch.cuda.device(ctx.dev):
            _engine.flat_scatter(_FlatParams._table(ctx.ent, [base + 4 * o for o in offs[:-1]]), len(ctx.ent), g)
        grads = [buf[o:o + n].view(s) if need else None
                 for o, n, s, need in zip(offs[:-1], sizes, ctx.shapes, ctx.needs_input_grad[2:])]
        return (None, None, *grads)


class _RefidFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, event, flat, mod):
        B, T = event.shape[:2]
        H, W = event.shape[-2:]
        train = bool(ctx.needs_input_grad[2])
        st = mod._state_for(B, T, H, W, train, x.device)
        eng = st["engine"]
        if flat is not None:  # None: the packed weights of this plan are current (forward-only call, parameters unchanged)
            eng.pack_weights(flat.detach().contiguous())
            st["packed_key"] = mod._pending_key
        out = torch.empty(B, T, mod.out_chn, H, W, device=x.device, dtype=torch.float32)
        eng.forward(x, event, out)
        st["generation"] += 1
        ctx.st, ctx.generation, ctx.mod = st, st["generation"], mod
        return out

    @staticmethod
    def backward(ctx, grad_out):
        st = ctx.st
        if st["generation"] != ctx.generation:
            raise RuntimeError("refid_b200: the activations saved for this backward were overwritten by a later forward "
                               "of the same shape (one forward/backward in flight per module and shape)")
        st["engine"].backward(grad_out.contiguous().float())
        g = st["grad_flat"]
        group = getattr(ctx.mod, "grad_sync_group", None)
        if group is not None:
            # Data parallelism: the only collective on the path is this all-reduce (mean) of the flat gradient --
            # one contiguous NCCL call over NVLink/NVSwitch instead of DDP's per-bucket reduction.
            sync_flat_grad(g, group)
        # `g` is the engine's own flat gradient buffer: the only consumer is _FlatParams.backward, which scatters it into the
        # per-parameter gradient tensors during this same backward pass (no copy; the buffer is rewritten by the next backward)
        return None, None, g, None


class FinalBidirectionAttenfusion(nn.Module):
    """Bi-directional event-recurrent U-Net with EGACA fusion (reference :81-218) on the sm_100a engine."""

    def __init__(self, img_chn, ev_chn, out_chn=3, skip_type='sum', recurrent_block_type='convlstm', activation='sigmoid',
                 num_encoders=4, base_num_channels=32, num_residual_blocks=2, norm=None, use_recurrent_upsample_conv=True,
                 num_block=3, use_first_dcn=False, use_reversed_voxel=False, infer_dtype='fp16'):
        super().__init__()
        # `infer_dtype` is the one key this backend adds to `network_g` (default keeps every shipped option file valid):
        # 16-bit storage type of forward-only (no_grad) calls -- 'fp16' (11 significant bits; output within ~1e-3 of the
        # fp32 reference, PSNR within 0.01 dB) or 'bf16' (bit-identical to the training forward).  Training is bf16.
        if infer_dtype not in ('fp16', 'bf16'):
            raise ValueError("infer_dtype must be 'fp16' or 'bf16'")
        self.infer_dtype = infer_dtype
        # Configurations no shipped option file uses are refused rather than silently diverging.
        for name, got, want in (("skip_type", skip_type, 'sum'), ("norm", norm, None), ("num_encoders", num_encoders, 3),
                                ("num_block", num_block, 1), ("num_residual_blocks", num_residual_blocks, 2),
                                ("base_num_channels", base_num_channels, 32),
                                ("use_recurrent_upsample_conv", use_recurrent_upsample_conv, True)):
            if got != want:
                raise NotImplementedError(f"refid_b200 supports {name}={want!r} only (got {got!r})")
        if not 1 <= out_chn <= 8:
            raise NotImplementedError("refid_b200 supports 1 <= out_chn <= 8")
        # recurrent_block_type, activation, use_first_dcn, use_reversed_voxel are accepted and have no effect,
        # exactly as in the reference (:59 stores torch.sigmoid but pred never applies it).
        self.img_chn, self.ev_chn, self.out_chn = img_chn, ev_chn, out_chn
        self.use_reversed_voxel = use_reversed_voxel
        b = base_num_channels
        self.head = _conv_layer(ev_chn, b, 5, 2)
        self.encoders_backward = nn.ModuleList()
        self.encoders_forward = nn.ModuleList()
        for l in range(3):
            cin, c = b << l, b << (l + 1)
            self.encoders_backward.append(_evr_layer(cin, c, False, l == 1))
            self.encoders_forward.append(_evr_layer(cin, c, True, l == 1))
        self.head_img = _conv_layer(img_chn, b, 5, 2)
        self.img_encoders = nn.ModuleList()
        for l in range(3):  # ImageEncoderConvBlock (recurrent_sub_modules.py:22-39)
            cin, c = b << l, b << (l + 1)
            e = _Holder()
            e.identity = nn.Conv2d(cin, c, 1, 1, 0)
            e.conv_1 = nn.Conv2d(cin, c, 3, padding=1)
            e.conv_2 = nn.Conv2d(c, c, 3, padding=1)
            e.down = nn.Conv2d(c, c, 4, 2, 1, bias=False)
            self.img_encoders.append(e)
        self.resblocks = nn.ModuleList()
        for _ in range(2):  # ResidualBlock (:468-486)
            r = _Holder()
            r.conv1 = nn.Conv2d(8 * b, 8 * b, 3, 1, 1)
            r.conv2 = nn.Conv2d(8 * b, 8 * b, 3, 1, 1)
            self.resblocks.append(r)
        self.decoders = nn.ModuleList()
        for i in range(3):  # TransposeRecurrentConvLayer (:370-384)
            cin = (8 * b) >> i
            d = _Holder()
            d.transposed_conv2d = nn.ConvTranspose2d(cin, cin // 2, 2, stride=2, padding=0)
            d.forward_trunk = _trunk(cin, cin // 2)
            self.decoders.append(d)
        self.pred = _conv_layer(b, out_chn, 3, 1)
        self._grad_sync_group = None
        # time steps per chunk of the level-major training schedule (engine option "tchunk"; None: the engine's default, as many as fit;
        # 0: step-major).  Not a constructor key: it changes launch order, never results beyond fp32 summation order.
        self.train_tchunk = None
        self._engines = {}   # device -> Engine (parameter table)
        self._states = {}    # (B,T,H,W,train,device) -> planned engine + buffers

    # ------------------------------------------------------------------------------------------
    # data parallelism without the DDP wrapper
    # ------------------------------------------------------------------------------------------
    @property
    def grad_sync_group(self):
        """torch.distributed group for the flat-gradient all-reduce inside backward (None: no collective).  Assigning a
        group also broadcasts rank 0's parameters to the group, as DDP's constructor does."""
        return self._grad_sync_group

    @grad_sync_group.setter
    def grad_sync_group(self, group):
        self._grad_sync_group = group
        if group is not None:
            broadcast_parameters(self, group)

    # ------------------------------------------------------------------------------------------
    # parameters -> the engine's flat vector
    # ------------------------------------------------------------------------------------------
    def _effective(self, key, P):
        """(weight as a conv-shaped tensor, bias) for one engine site, as differentiable functions of the parameters."""
        if key == "head_img" or key == "head":
            return P[key + ".conv2d.weight"], P[key + ".conv2d.bias"]
        if key == "enc0_in":
            return (torch.cat((P["encoders_backward.0.conv.conv2d.weight"], P["encoders_forward.0.conv.conv2d.weight"]), 0),
                    torch.cat((P["encoders_backward.0.conv.conv2d.bias"], P["encoders_forward.0.conv.conv2d.bias"]), 0))
        if key == "pred":
            w, b = P["pred.conv2d.weight"], P["pred.conv2d.bias"]
            return F.pad(w, (0, 0, 0, 0, 0, 0, 0, 32 - w.shape[0])), F.pad(b, (0, 32 - b.shape[0]))
        if ".atten_fuse." in key:
            a, leaf = key.split(".atten_fuse.")
            a += ".atten_fuse"
            if leaf in ("conv1", "conv1_e", "conv4"):  # LayerNorm affine folded into the following 1x1 conv
                n = {"conv1": "norm1", "conv1_e": "norm1_e", "conv4": "norm2"}[leaf]
                w, b = P[f"{a}.{leaf}.weight"], P[f"{a}.{leaf}.bias"]
                nw, nb = P[f"{a}.{n}.weight"], P[f"{a}.{n}.bias"]
                return w * nw.view(1, -1, 1, 1), b + w[:, :, 0, 0] @ nb
            if leaf == "conv3":  # x*beta folded (fusion_modules.py:317-321)
                beta = P[a + ".beta"].view(-1)
                return P[a + ".conv3.weight"] * beta.view(-1, 1, 1, 1), P[a + ".conv3.bias"] * beta
            if leaf == "conv5s":  # y = conv_y_side(y) + conv5(x)*gamma as one GEMM over K = [y ; x] (:325-333)
                gamma = P[a + ".gamma"].view(-1)
                w = torch.cat((P[a + ".conv_y_side.weight"], P[a + ".conv5.weight"] * gamma.view(-1, 1, 1, 1)), 1)
                return w, P[a + ".conv_y_side.bias"] + P[a + ".conv5.bias"] * gamma
            return P[f"{a}.{leaf}.weight"], P[f"{a}.{leaf}.bias"]  # conv2 / conv2_e (depthwise), se_1.1 / se_1.3
        for suffix in (".conv", ".fuse_two_dir"):
            if key.endswith(suffix):
                return P[key + ".conv2d.weight"], P[key + ".conv2d.bias"]
        return P[key + ".weight"], P.get(key + ".bias")

    _FOLDED = ("head", "head_img", "enc0_in", "pred")

    def _is_plain(self, key):
        """The site's flat entry is a pure permutation of ONE parameter tensor (no fold, no stacking, no padding)."""
        if key in self._FOLDED:
            return False
        if ".atten_fuse." in key:
            return key.rsplit(".", 1)[1] in ("conv2", "conv2_e") or key.endswith((".se_1.1", ".se_1.3"))
        return True

    def _flat_inputs(self, eng):
        """Host side of the flat parameter vector: the tensors the gather kernel reads and, per tensor, its table entry
        (flat offset, mode, taps, R, Cc).  Plain sites hand their parameter tensors over as they are (the kernel applies
        the (Cout,Cin,kh,kw) -> [tap][Cin][Cout] permutation, mode 1); the dozen folded / stacked / padded sites are first
        brought into the flat layout by a few differentiable torch ops (mode 0)."""
        P = dict(self.named_parameters())
        ins, ent = [], []  # input tensors; (flat offset, mode, taps, R, Cc) per input
        for e in eng.entries:
            key, kind = e["key"], e["kind"]
            w, b = self._effective(key, P)
            if self._is_plain(key) and kind != _engine.KIND_ROWS5:
                if kind == _engine.KIND_RAW:
                    ent.append((e["w_off"], 0, 1, e["R"], e["Cc"]))
                else:  # Conv2d (Cout,Cin,kh,kw) / ConvTranspose2d (Cin,Cout,2,2): the kernel permutes
                    assert tuple(w.shape[:2]) == (e["Cc"], e["R"]) and w.shape[2] * w.shape[3] == e["taps"], (e, tuple(w.shape))
                    ent.append((e["w_off"], 1, e["taps"], e["R"], e["Cc"]))
                assert w.numel() == e["taps"] * e["R"] * e["Cc"], (e, tuple(w.shape))
                ins.append(w)
            else:
                if kind == _engine.KIND_ROWS5:  # (32,Cin,5,5) -> [ky][kx*Cin + c][32], K zero-padded to R
                    g = w.permute(2, 3, 1, 0).reshape(5, 5 * w.shape[1], w.shape[0])
                    g = F.pad(g, (0, 0, 0, e["R"] - g.shape[1]))
                elif kind == _engine.KIND_RAW:
                    g = w.reshape(e["R"], e["Cc"])
                else:  # Conv2d (Cout,Cin,kh,kw) -> [ky*kw+kx][Cin][Cout]
                    g = w.permute(2, 3, 1, 0).reshape(e["taps"], w.shape[1], w.shape[0])
                assert g.numel() == e["taps"] * e["R"] * e["Cc"], (e, tuple(w.shape))
                ins.append(g.contiguous())
                ent.append((e["w_off"], 0, e["taps"], e["R"], e["Cc"]))
            if e["nbias"]:
                assert b is not None and b.numel() == e["nbias"], e
                ins.append(b)
                ent.append((e["b_off"], 0, 1, 1, e["nbias"]))
        return ins, tuple(ent)

    def _flat(self, eng):
        """The engine's flat parameter vector as a differentiable function of the parameters: ONE gather kernel over the
        table of `_flat_inputs` (refid_flat_gather); its backward is ONE scatter (refid_flat_scatter)."""
        ins, ent = self._flat_inputs(eng)
        return _FlatParams.apply(eng.flat_floats, ent, *ins)

    # ------------------------------------------------------------------------------------------
    # engine / plan cache
    # ------------------------------------------------------------------------------------------
    def _state_for(self, B, T, H, W, train, device):
        key = (B, T, H, W, train, str(device))
        st = self._states.get(key)
        if st is None:
            eng = _engine.Engine(self.img_chn, self.ev_chn, self.out_chn, 32)
            eng.set_option("infer_fp16", self.infer_dtype == 'fp16')
            if self.train_tchunk is not None:
                eng.set_option("tchunk", int(self.train_tchunk))
            ws = torch.empty(eng.workspace_bytes(B, T, H, W, train), dtype=torch.uint8, device=device)
            wpack = torch.empty(eng.wpack_bytes, dtype=torch.uint8, device=device)
            grad_flat = torch.zeros(eng.flat_floats, dtype=torch.float32, device=device) if train else None
            eng.plan(B, T, H, W, train, ws, wpack, grad_flat)
            st = {"engine": eng, "grad_flat": grad_flat, "generation": 0, "packed_key": None}
            self._states[key] = st
        return st

    def _table_engine(self):
        if "table" not in self._engines:
            self._engines["table"] = _engine.Engine(self.img_chn, self.ev_chn, self.out_chn, 32)
        return self._engines["table"]

    def release_buffers(self):
        """Drop all cached plans and their workspaces."""
        self._states.clear()

    def forward(self, x, event):
        if not x.is_cuda:
            raise RuntimeError("refid_b200 runs on CUDA (sm_100a) only; there is no CPU path")
        if x.dim() == 5:  # (B,t,c,H,W) -> (B,t*c,H,W)   (reference :140-141)
            x = x.flatten(1, 2)
        if event.dim() != 5 or x.dim() != 4:
            raise ValueError("expected x (B,C,H,W) or (B,t,c,H,W) and event (B,T,C,H,W)")
        B, T, ec, H, W = event.shape
        if x.shape[1] != self.img_chn or ec != self.ev_chn or x.shape[0] != B or tuple(x.shape[-2:]) != (H, W):
            raise ValueError(f"shape mismatch: x {tuple(x.shape)} event {tuple(event.shape)} for img_chn={self.img_chn} "
                             f"ev_chn={self.ev_chn}")
        if H % 8 or W % 8:
            raise ValueError("H and W must be multiples of 8 (three stride-2 levels)")
        params = list(self.parameters())
        self._pending_key = tuple((p.data_ptr(), p._version) for p in params)
        if torch.is_grad_enabled() and any(p.requires_grad for p in params):
            flat = self._flat(self._table_engine())  # differentiable: autograd maps the flat gradient back through the folds
        else:
            # forward-only: the flat vector and the packed weights are rebuilt only when a parameter changed since this
            # plan last packed them (every in-place update bumps the parameter's version counter; refid_b200.optim does too)
            st = self._state_for(B, T, H, W, False, x.device)
            flat = None
            if st["packed_key"] != self._pending_key:
                with torch.no_grad():
                    flat = self._flat(self._table_engine())
        return _RefidFunction.apply(x.float().contiguous(), event.float().contiguous(), flat, self)

    def extra_repr(self):
        return (f"img_chn={self.img_chn}, ev_chn={self.ev_chn}, out_chn={self.out_chn}, infer_dtype={self.infer_dtype}, "
                f"backend=librefid_b200.so (sm_100a)")


def flat_param_count(img_chn, ev_chn):
    return sum(int(math.prod(p.shape)) for p in FinalBidirectionAttenfusion(img_chn, ev_chn, num_encoders=3, num_block=1).parameters())
