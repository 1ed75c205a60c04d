"""Single-kernel parity cases (CUDA tap-GEMM / wgrad vs torch fp32 on bf16-rounded operands)."""
import ctypes

import torch
import torch.nn.functional as F

from refid_b200 import _lib, packing

CK_3X3, CK_1X1, CK_DOWN4, CK_UP2, CK_DOWN4_DGRAD, CK_UP2_DGRAD = range(6)
CK_DOWN4_DGRAD_HALO = 7
CK_DOWN4_HALO = 8
ACT_NONE, ACT_LRELU, ACT_GELU = 0, 1, 2


def nhwc(t):  # (N,C,H,W) fp32 -> (N,H,W,C) bf16 contiguous
    return t.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)


def nchw(t):  # (N,H,W,C) bf16 -> (N,C,H,W) fp32
    return t.float().permute(0, 3, 1, 2).contiguous()


def rb(t):
    return t.to(torch.bfloat16).float()


def run_conv(kind, ins, wp, wrows_per_tap, cout, out_shape, parity=0, w_row0=0, bias=None, pre=None, sv=None, act=0,
             slope=0.0, split=False, post=None, f32=None, out_init=None):
    """ins: list of NHWC bf16 tensors. Returns dict of outputs (NHWC bf16)."""
    L = _lib.lib()
    N, H, W, C0 = ins[0].shape
    cg = cout // 2 if split else cout
    mk = (lambda: torch.full(out_shape[:3] + (cg,), float("nan"), device="cuda", dtype=torch.bfloat16)) \
        if out_init is None else (lambda: out_init.clone())
    out = mk()
    out_b = mk() if split else None
    out2 = mk() if post is not None else None
    in1 = ins[1] if len(ins) > 1 else None
    rc = L.refid_test_conv(kind, parity, _lib.ptr(ins[0]), C0, _lib.ptr(in1), in1.shape[3] if in1 is not None else 0,
                           N, H, W, _lib.ptr(wp), ctypes.c_long(wp.shape[0]), wp.shape[1], wrows_per_tap, w_row0, cout,
                           _lib.ptr(bias), _lib.ptr(pre), _lib.ptr(sv), act, ctypes.c_float(slope), _lib.ptr(out),
                           _lib.ptr(out_b), _lib.ptr(out2), _lib.ptr(post), _lib.ptr(f32), None)
    _lib.check(rc, "refid_test_conv")
    torch.cuda.synchronize()
    flag = _lib.abort_flag()
    if flag:
        raise RuntimeError(f"kernel aborted: mbarrier timeout code 0x{flag:x}")
    return {"out": out, "out_b": out_b, "out2": out2}


def run_wgrad(kind, ps, q, rows):
    L = _lib.lib()
    N, H, W, C0 = ps[0].shape
    p1 = ps[1] if len(ps) > 1 else None
    out = torch.zeros(rows, q.shape[3], device="cuda", dtype=torch.float32)
    rc = L.refid_test_wgrad(kind, _lib.ptr(ps[0]), C0, _lib.ptr(p1), p1.shape[3] if p1 is not None else 0, N, H, W,
                            _lib.ptr(q), q.shape[3], _lib.ptr(out), None)
    _lib.check(rc, "refid_test_wgrad")
    torch.cuda.synchronize()
    flag = _lib.abort_flag()
    if flag:
        raise RuntimeError(f"kernel aborted: mbarrier timeout code 0x{flag:x}")
    return out


def _err(a, b):
    a, b = a.float(), b.float()
    return (a - b).abs().max().item(), b.abs().max().item()


def g(*shape, seed=0):
    gen = torch.Generator(device="cuda").manual_seed(seed)
    return torch.randn(*shape, device="cuda", generator=gen)


# Each case returns (max_abs_err, ref_max_abs). Tolerance: bf16 output rounding => err <= ~1e-2 * ref_max.
def case_conv3x3(N=2, H=16, W=16, cin=64, cout=64, cin2=0, bias=True, act=ACT_LRELU, slope=0.1, pre=False):
    x = rb(g(N, cin + cin2, H, W, seed=1))
    w = rb(g(cout, cin + cin2, 3, 3, seed=2) / (3 * (cin + cin2) ** 0.5))
    b = g(cout, seed=3) * 0.1 if bias else None
    r = rb(g(N, cout, H, W, seed=4)) if pre else None
    ref = F.conv2d(x, w, b, padding=1)
    if pre:
        ref = ref + r
    if act == ACT_LRELU:
        ref = F.leaky_relu(ref, slope)
    elif act == ACT_GELU:
        ref = F.gelu(ref)
    ins = [nhwc(x[:, :cin])] + ([nhwc(x[:, cin:])] if cin2 else [])
    o = run_conv(CK_3X3, ins, packing.pack_fwd(w).to(torch.bfloat16), cout, cout, (N, H, W), bias=b,
                 pre=nhwc(r) if pre else None, act=act, slope=slope)
    return _err(nchw(o["out"]), ref)


def case_dgrad3x3_split(N=2, H=16, W=16, c=64):
    """dgrad of conv3x3 (2c -> c): two output groups, second masked by a saved activation."""
    gy = rb(g(N, c, H, W, seed=1))
    w = rb(g(c, 2 * c, 3, 3, seed=2) / (3 * (2 * c) ** 0.5))
    sv = rb(g(N, c, H, W, seed=5))
    ref = F.conv_transpose2d(gy, w, padding=1)
    wp = packing.pack_dgrad_s1(w).to(torch.bfloat16)
    o = run_conv(CK_3X3, [nhwc(gy)], wp, 2 * c, 2 * c, (N, H, W), split=True, sv=nhwc(sv), act=ACT_LRELU, slope=0.25)
    mask = torch.where(sv > 0, 1.0, 0.25)
    e1 = _err(nchw(o["out"]), ref[:, :c] * mask)
    e2 = _err(nchw(o["out_b"]), ref[:, c:] * mask)
    return max(e1[0], e2[0]), max(e1[1], e2[1])


def case_conv1x1(N=2, H=16, W=16, cin=128, cout=64):
    x = rb(g(N, cin, H, W, seed=1))
    w = rb(g(cout, cin, 1, 1, seed=2) / cin ** 0.5)
    b = g(cout, seed=3) * 0.1
    post = rb(g(N, cout, H, W, seed=6))
    ref = F.gelu(F.conv2d(x, w, b))
    o = run_conv(CK_1X1, [nhwc(x)], packing.pack_fwd(w).to(torch.bfloat16), cout, cout, (N, H, W), bias=b, act=ACT_GELU,
                 post=nhwc(post))
    e1 = _err(nchw(o["out"]), ref)
    e2 = _err(nchw(o["out2"]), ref + post)
    return max(e1[0], e2[0]), max(e1[1], e2[1])


def case_down4(N=2, H=16, W=16, c=64):
    x = rb(g(N, c, H, W, seed=1))
    w = rb(g(c, c, 4, 4, seed=2) / (4 * c ** 0.5))
    ref = F.conv2d(x, w, None, stride=2, padding=1)
    o = run_conv(CK_DOWN4, [nhwc(x)], packing.pack_fwd(w).to(torch.bfloat16), c, c, (N, H // 2, W // 2))
    return _err(nchw(o["out"]), ref)


def case_down4_halo(N=2, H=32, W=32, c=64, post=False):
    """Stride-2 conv as a masked 3x3 over the four parity views (halo-conv engine), optional skip-sum second output."""
    x = rb(g(N, c, H, W, seed=1))
    w = rb(g(c, c, 4, 4, seed=2) / (4 * c ** 0.5))
    ref = F.conv2d(x, w, None, stride=2, padding=1)
    po = rb(g(N, c, H // 2, W // 2, seed=6)) if post else None
    o = run_conv(CK_DOWN4_HALO, [nhwc(x)], packing.pack_down_fwd_halo(w).to(torch.bfloat16), c, c, (N, H // 2, W // 2),
                 post=nhwc(po) if post else None)
    e1 = _err(nchw(o["out"]), ref)
    if post:
        e2 = _err(nchw(o["out2"]), ref + po)
        return max(e1[0], e2[0]), max(e1[1], e2[1])
    return e1


def case_up2(N=2, H=8, W=8, cin=128, cout=64):
    x = rb(g(N, cin, H, W, seed=1))
    w = rb(g(cin, cout, 2, 2, seed=2) / cin ** 0.5)
    b = g(cout, seed=3) * 0.1
    ref = F.conv_transpose2d(x, w, b, stride=2)
    o = run_conv(CK_UP2, [nhwc(x)], packing.pack_up2_fwd(w).to(torch.bfloat16), 4 * cout, cout, (N, 2 * H, 2 * W), bias=b)
    return _err(nchw(o["out"]), ref)


def case_down4_dgrad(N=2, H=8, W=8, c=64):
    gy = rb(g(N, c, H, W, seed=1))
    w = rb(g(c, c, 4, 4, seed=2) / (4 * c ** 0.5))
    ref = F.conv_transpose2d(gy, w, stride=2, padding=1)
    wp = packing.pack_down_dgrad(w).to(torch.bfloat16)
    out = torch.full((N, 2 * H, 2 * W, c), float("nan"), device="cuda", dtype=torch.bfloat16)
    for par in range(4):
        o = run_conv(CK_DOWN4_DGRAD, [nhwc(gy)], wp, c, c, (N, 2 * H, 2 * W), parity=par, w_row0=par * 4 * c, out_init=out)
        out = o["out"]
    return _err(nchw(out), ref)


def case_down4_dgrad_halo(N=2, H=8, W=8, c=64, masked=True):
    """All four output parities in one halo-conv launch (tap masks + stride-2 scatter), accumulate + LeakyReLU mask."""
    gy = rb(g(N, c, H, W, seed=1))
    w = rb(g(c, c, 4, 4, seed=2) / (4 * c ** 0.5))
    sv = rb(g(N, c, 2 * H, 2 * W, seed=5))
    ref = F.conv_transpose2d(gy, w, stride=2, padding=1)
    if masked:
        ref = ref * torch.where(sv > 0, 1.0, 0.2)
    wp = packing.pack_down_dgrad_halo(w).to(torch.bfloat16)
    o = run_conv(CK_DOWN4_DGRAD_HALO, [nhwc(gy)], wp, 4 * c, c, (N, 2 * H, 2 * W), sv=nhwc(sv) if masked else None,
                 act=ACT_LRELU if masked else 0, slope=0.2)
    return _err(nchw(o["out"]), ref)


def case_up2_dgrad(N=2, H=16, W=16, cin=128, cout=64):
    """dgrad of ConvTranspose(cin->cout): dIn = conv2x2s2(dOut)."""
    go = rb(g(N, cout, H, W, seed=1))
    w = rb(g(cin, cout, 2, 2, seed=2) / cout ** 0.5)
    ref = F.conv2d(go, w, None, stride=2)
    o = run_conv(CK_UP2_DGRAD, [nhwc(go)], packing.pack_up2_dgrad(w).to(torch.bfloat16), cin, cin, (N, H // 2, W // 2))
    return _err(nchw(o["out"]), ref)


def case_wgrad3x3(N=2, H=16, W=16, cin=64, cout=64, cin2=0):
    x = rb(g(N, cin + cin2, H, W, seed=1))
    gy = rb(g(N, cout, H, W, seed=2))
    ref = torch.nn.grad.conv2d_weight(x, (cout, cin + cin2, 3, 3), gy, padding=1)
    ps = [nhwc(x[:, :cin])] + ([nhwc(x[:, cin:])] if cin2 else [])
    d = run_wgrad(CK_3X3, ps, nhwc(gy), 9 * (cin + cin2))
    return _err(packing.unpack_wgrad_conv(d, cout, cin + cin2, 3, 3), ref)


def case_wgrad1x1(N=2, H=16, W=16, cin=128, cout=64):
    x = rb(g(N, cin, H, W, seed=1))
    gy = rb(g(N, cout, H, W, seed=2))
    ref = torch.nn.grad.conv2d_weight(x, (cout, cin, 1, 1), gy)
    d = run_wgrad(CK_1X1, [nhwc(x)], nhwc(gy), cin)
    return _err(packing.unpack_wgrad_conv(d, cout, cin, 1, 1), ref)


def case_wgrad_down4(N=2, H=16, W=16, c=64):
    x = rb(g(N, c, H, W, seed=1))
    gy = rb(g(N, c, H // 2, W // 2, seed=2))
    ref = torch.nn.grad.conv2d_weight(x, (c, c, 4, 4), gy, stride=2, padding=1)
    d = run_wgrad(CK_DOWN4, [nhwc(x)], nhwc(gy), 16 * c)
    return _err(packing.unpack_wgrad_conv(d, c, c, 4, 4), ref)


def case_wgrad_up2(N=2, H=8, W=8, cin=128, cout=64):
    x = rb(g(N, cin, H, W, seed=1))
    go = rb(g(N, cout, 2 * H, 2 * W, seed=2))
    # dW[ci,co,a,b] = sum x[ci,y,x] * go[co,2y+a,2x+b]
    ref = torch.einsum("nihw,nohawb->ioab", x, go.view(N, cout, H, 2, W, 2))
    d = run_wgrad(CK_UP2_DGRAD, [nhwc(go)], nhwc(x), 4 * cout)
    return _err(packing.unpack_wgrad_up2(d, cin, cout), ref)


CASES = {
    "conv3x3_64_64": lambda: case_conv3x3(),
    "conv3x3_64_64_nobias_noact": lambda: case_conv3x3(bias=False, act=ACT_NONE),
    "conv3x3_dual_64+64_64": lambda: case_conv3x3(cin2=64, pre=True),
    "conv3x3_32_64_bk32": lambda: case_conv3x3(cin=32),
    "conv3x3_dual_32+32_32": lambda: case_conv3x3(cin=32, cin2=32, cout=32),
    "conv3x3_32_32_ragged_pre": lambda: case_conv3x3(cin=32, cout=32, H=40, W=28, N=3, pre=True),
    "conv3x3_32_128_big": lambda: case_conv3x3(cin=32, cout=128, H=64, W=48, N=2),
    "conv3x3_96_64": lambda: case_conv3x3(cin=96, cout=64, H=24, W=24),
    "dgrad3x3_split_32": lambda: case_dgrad3x3_split(c=32),
    "conv3x3_128_128": lambda: case_conv3x3(cin=128, cout=128, H=8, W=8),
    "conv3x3_128_128_pair_odd_tiles": lambda: case_conv3x3(cin=128, cout=128, H=32, W=24, N=3, pre=True),
    "conv3x3_dual_128+128_128_pair": lambda: case_conv3x3(cin=128, cin2=128, cout=128, H=64, W=64, N=2),
    "conv3x3_256_256_pair": lambda: case_conv3x3(cin=256, cout=256, H=32, W=40, N=2),
    "conv3x3_256_256_h4": lambda: case_conv3x3(cin=256, cout=256, H=4, W=4, N=3),
    "conv3x3_ragged_24x20": lambda: case_conv3x3(H=24, W=20, N=1),
    # 168 two-block items on 148 SMs: the 20 items of the last round run as 40 half items (tail splitting, haloconv.cuh)
    "conv3x3_64_64_tail_split": lambda: case_conv3x3(H=128, W=112, N=3),
    "conv3x3_64_64_tail_split_ragged_pre": lambda: case_conv3x3(H=120, W=108, N=3, pre=True),
    "conv3x3_128_128_tail_split_streamed": lambda: case_conv3x3(cin=128, cout=128, H=128, W=112, N=3, pre=True),
    "conv3x3_dual_64+64_64_tail_split": lambda: case_conv3x3(cin2=64, H=120, W=112, N=3),
    "conv3x3_32_32_tail_split": lambda: case_conv3x3(cin=32, cout=32, H=128, W=112, N=3),
    "dgrad3x3_split_tail_split": lambda: case_dgrad3x3_split(N=3, H=128, W=112),
    "conv3x3_64_64_gelu": lambda: case_conv3x3(act=ACT_GELU),
    "dgrad3x3_split": lambda: case_dgrad3x3_split(),
    "conv1x1_128_64_gelu_post": lambda: case_conv1x1(),
    "down4_64": lambda: case_down4(),
    "down4_128_ragged": lambda: case_down4(c=128, H=12, W=20, N=1),
    "up2_128_64": lambda: case_up2(),
    "up2_64_32": lambda: case_up2(cin=64, cout=32),
    "down4_halo_64": lambda: case_down4_halo(),
    "down4_halo_64_ragged_post": lambda: case_down4_halo(H=40, W=24, N=3, post=True),
    "down4_halo_128": lambda: case_down4_halo(c=128, H=48, W=32, N=2),
    "down4_halo_256": lambda: case_down4_halo(c=256, H=16, W=16, N=2, post=True),
    "down4_dgrad_64": lambda: case_down4_dgrad(),
    "down4_dgrad_halo_64": lambda: case_down4_dgrad_halo(),
    "down4_dgrad_halo_128_ragged": lambda: case_down4_dgrad_halo(c=128, H=20, W=12, N=3),
    "down4_dgrad_halo_256": lambda: case_down4_dgrad_halo(c=256, H=8, W=8, N=1, masked=False),
    "up2_dgrad_128_64": lambda: case_up2_dgrad(),
    "wgrad3x3_64_64": lambda: case_wgrad3x3(),
    "wgrad3x3_dual_64+64_64": lambda: case_wgrad3x3(cin2=64),
    "wgrad3x3_32_32": lambda: case_wgrad3x3(cin=32, cout=32),
    "wgrad3x3_dual_32+32_32_ragged": lambda: case_wgrad3x3(cin=32, cin2=32, cout=32, H=40, W=20, N=3),
    "wgrad3x3_32_32_many_tiles": lambda: case_wgrad3x3(cin=32, cout=32, H=64, W=96, N=4),
    "wgrad3x3_128_128": lambda: case_wgrad3x3(cin=128, cout=128, H=8, W=8),
    "wgrad3x3_256_256": lambda: case_wgrad3x3(cin=256, cout=256, H=8, W=8),
    "wgrad3x3_ragged": lambda: case_wgrad3x3(H=24, W=20, N=3),
    "wgrad3x3_dual_128+128_128_ragged": lambda: case_wgrad3x3(cin=128, cout=128, cin2=128, H=24, W=12, N=3),
    "wgrad3x3_128_256": lambda: case_wgrad3x3(cin=128, cout=256, H=16, W=24, N=2),
    "wgrad3x3_64_64_many_tiles": lambda: case_wgrad3x3(H=64, W=64, N=5),
    "wgrad1x1_128_64": lambda: case_wgrad1x1(),
    "wgrad_down4_64": lambda: case_wgrad_down4(),
    "wgrad_down4_64_ragged": lambda: case_wgrad_down4(N=3, H=40, W=24),
    "wgrad_down4_128": lambda: case_wgrad_down4(N=2, H=48, W=32, c=128),
    "wgrad_down4_256": lambda: case_wgrad_down4(N=2, H=16, W=32, c=256),
    "wgrad_up2_128_64": lambda: case_wgrad_up2(),
}
