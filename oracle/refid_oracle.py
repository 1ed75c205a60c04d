"""CPU oracle for the REFID hot path (TEST INFRASTRUCTURE -- never imported by the product).

A functional restatement (torch CPU ops, fp32 or fp64, autograd for the backward) of
`FinalBidirectionAttenfusion.forward` and the Charbonnier training objective.  It exists only so that
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs can check and
time-compare the CUDA path; `refid_b200/` must never import it.

Parity pin: the reference ships no tests or golden vectors for this path, so the oracle is
pinned against outputs of the reference module itself, generated in the build container by
`tests/golden/make_golden.py` (which imports /root/reference unmodified) and committed under `tests/golden/`.
`tests/test_oracle_golden.py` replays those vectors.

Every function cites the reference lines it restates (paths relative to /root/reference/basicsr/models/archs/).
The parameter dictionary `P` uses the reference's state_dict names verbatim.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# ----------------------------------------------------------------------------------------------
# parameter inventory (names/shapes as the reference module tree produces them)
# ----------------------------------------------------------------------------------------------
def param_shapes(img_chn: int, ev_chn: int, base: int = 32, out_chn: int = 3) -> "Dict[str, tuple]":
    """Ordered name -> shape map of the 183 parameters (XXNet_final_attenfusion_arch.py:90-128)."""
    S: Dict[str, tuple] = {}

    def conv(name, cout, cin, k, bias=True):
        S[name + ".weight"] = (cout, cin, k, k)
        if bias:
            S[name + ".bias"] = (cout,)

    def trunk(prefix, cin, c):  # ConvResidualBlocks(num_block=1), recurrent_sub_modules.py:710-758
        conv(prefix + ".main.0", c, cin, 3)
        conv(prefix + ".main.2.0.conv1", c, c, 3)
        conv(prefix + ".main.2.0.conv2", c, c, 3)

    def egaca(prefix, c, c_out):  # fusion_modules.py:237-288
        conv(prefix + ".conv1", c, c, 1)
        S[prefix + ".conv2.weight"] = (c, 1, 3, 3)
        S[prefix + ".conv2.bias"] = (c,)
        conv(prefix + ".conv1_e", c, c, 1)
        S[prefix + ".conv2_e.weight"] = (c, 1, 3, 3)
        S[prefix + ".conv2_e.bias"] = (c,)
        conv(prefix + ".conv3", c, 2 * c, 1)
        conv(prefix + ".se_1.1", c // 2, c, 1)
        conv(prefix + ".se_1.3", c, c // 2, 1)
        conv(prefix + ".se_2.1", c // 2, c, 1)
        conv(prefix + ".se_2.3", c, c // 2, 1)
        conv(prefix + ".conv4", 2 * c, c, 1)
        conv(prefix + ".conv5", c_out, 2 * c, 1)
        conv(prefix + ".conv_y_side", c_out, c, 1)
        for n in ("norm1", "norm1_e", "norm2"):
            S[f"{prefix}.{n}.weight"] = (c,)
            S[f"{prefix}.{n}.bias"] = (c,)
        S[prefix + ".beta"] = (1, c, 1, 1)
        S[prefix + ".gamma"] = (1, c_out, 1, 1)

    conv("head.conv2d", base, ev_chn, 5)
    for l in range(3):
        cin, c = base * 2 ** l, base * 2 ** (l + 1)
        for d, fuse in (("encoders_backward", False), ("encoders_forward", True)):
            p = f"{d}.{l}"
            conv(p + ".conv.conv2d", c, cin, 3)
            if l == 1:
                egaca(p + ".atten_fuse", cin, c)
            trunk(p + ".recurrent_block.forward_trunk", 2 * c, c)
            if fuse:
                conv(p + ".fuse_two_dir.conv2d", c, 2 * c, 1)
            conv(p + ".down", c, c, 4, bias=False)
    conv("head_img.conv2d", base, img_chn, 5)
    for l in range(3):
        cin, c = base * 2 ** l, base * 2 ** (l + 1)
        p = f"img_encoders.{l}"
        conv(p + ".identity", c, cin, 1)
        conv(p + ".conv_1", c, cin, 3)
        conv(p + ".conv_2", c, c, 3)
        conv(p + ".down", c, c, 4, bias=False)
    for i in range(2):
        conv(f"resblocks.{i}.conv1", 8 * base, 8 * base, 3)
        conv(f"resblocks.{i}.conv2", 8 * base, 8 * base, 3)
    for i in range(3):
        cin = base * 2 ** (3 - i)
        c = cin // 2
        S[f"decoders.{i}.transposed_conv2d.weight"] = (cin, c, 2, 2)
        S[f"decoders.{i}.transposed_conv2d.bias"] = (c,)
        trunk(f"decoders.{i}.forward_trunk", 2 * c, c)
    conv("pred.conv2d", out_chn, base, 3)
    return S


# Parameters that never influence the output: grad must be exactly zero / None.
def dead_params(P_names) -> List[str]:
    dead = []
    for n in P_names:
        if ".1.conv.conv2d." in n or ".se_2." in n or n == "encoders_backward.2.down.weight":
            dead.append(n)
    return dead


# ----------------------------------------------------------------------------------------------
# building blocks
# ----------------------------------------------------------------------------------------------
def _lrelu(x, s):
    return F.leaky_relu(x, s)


def layernorm2d(x: Tensor, w: Tensor, b: Tensor, eps: float = 1e-6) -> Tensor:
    """Per-pixel LayerNorm over channels, biased variance (fusion_modules.py:97-134)."""
    mu = x.mean(1, keepdim=True)
    var = (x - mu).pow(2).mean(1, keepdim=True)
    y = (x - mu) / (var + eps).sqrt()
    return y * w.view(1, -1, 1, 1) + b.view(1, -1, 1, 1)


def egaca(P, p: str, event_feat: Tensor, image_feat: Tensor) -> Tensor:
    """CrossmodalAtten_imgeventalladd.forward (fusion_modules.py:290-333).

    se_1 gates BOTH modalities from the event branch's pooled statistics (:312-313); se_2 is never used.
    GELU is the exact erf form (:271)."""
    c = event_feat.shape[1]
    x = layernorm2d(image_feat, P[p + ".norm1.weight"], P[p + ".norm1.bias"])
    xe = layernorm2d(event_feat, P[p + ".norm1_e.weight"], P[p + ".norm1_e.bias"])
    x = F.conv2d(x, P[p + ".conv1.weight"], P[p + ".conv1.bias"])
    x = F.gelu(F.conv2d(x, P[p + ".conv2.weight"], P[p + ".conv2.bias"], padding=1, groups=c))
    xe = F.conv2d(xe, P[p + ".conv1_e.weight"], P[p + ".conv1_e.bias"])
    xe = F.gelu(F.conv2d(xe, P[p + ".conv2_e.weight"], P[p + ".conv2_e.bias"], padding=1, groups=c))
    pooled = xe.mean((2, 3), keepdim=True)
    s = F.relu(F.conv2d(pooled, P[p + ".se_1.1.weight"], P[p + ".se_1.1.bias"]))
    s = torch.sigmoid(F.conv2d(s, P[p + ".se_1.3.weight"], P[p + ".se_1.3.bias"]))
    x = torch.cat((x * s, xe * s), 1)
    x = F.conv2d(x, P[p + ".conv3.weight"], P[p + ".conv3.bias"])
    y = event_feat + image_feat + x * P[p + ".beta"]
    x = layernorm2d(y, P[p + ".norm2.weight"], P[p + ".norm2.bias"])
    x = F.gelu(F.conv2d(x, P[p + ".conv4.weight"], P[p + ".conv4.bias"]))
    x = F.conv2d(x, P[p + ".conv5.weight"], P[p + ".conv5.bias"])
    y = F.conv2d(y, P[p + ".conv_y_side.weight"], P[p + ".conv_y_side.bias"])
    return y + x * P[p + ".gamma"]


def trunk(P, p: str, x: Tensor, h_prev: Optional[Tensor]) -> Tensor:
    """conv3x3(cat(x,h)) + LReLU(0.1) then one ResidualBlockNoBN (recurrent_sub_modules.py:659-678,710-758).
    `h_prev is None` means the all-zero initial state (:666-668, :395-397)."""
    if h_prev is None:
        h_prev = torch.zeros_like(x)
    v = _lrelu(F.conv2d(torch.cat((x, h_prev), 1), P[p + ".main.0.weight"], P[p + ".main.0.bias"], padding=1), 0.1)
    r = F.relu(F.conv2d(v, P[p + ".main.2.0.conv1.weight"], P[p + ".main.2.0.conv1.bias"], padding=1))
    return v + F.conv2d(r, P[p + ".main.2.0.conv2.weight"], P[p + ".main.2.0.conv2.bias"], padding=1)


def evr_layer(P, p: str, level: int, x: Tensor, y: Optional[Tensor], h_prev: Optional[Tensor],
              h_other: Optional[Tensor], rec=None, tag: str = ""):
    """SimpleRecurrentThenDownAttenfusionmodifiedConvLayer.forward (recurrent_sub_modules.py:270-296).

    The in-conv is a ConvLayer (conv + LReLU 0.2) followed by a second LReLU 0.2 (:279-285) => slope 0.04.
    Level 1 replaces it with EGACA (:274-276).  Returns (downsampled output, recurrent state h)."""
    if y is not None and level == 1:
        u = egaca(P, p + ".atten_fuse", x, y)
    else:
        if y is not None:
            x = x + y
        u = F.conv2d(x, P[p + ".conv.conv2d.weight"], P[p + ".conv.conv2d.bias"], padding=1)
        u = _lrelu(_lrelu(u, 0.2), 0.2)
    h = trunk(P, p + ".recurrent_block.forward_trunk", u, h_prev)
    o = h
    if h_other is not None:
        o = _lrelu(F.conv2d(torch.cat((h, h_other), 1), P[p + ".fuse_two_dir.conv2d.weight"],
                            P[p + ".fuse_two_dir.conv2d.bias"]), 0.2)
    o = F.conv2d(o, P[p + ".down.weight"], None, stride=2, padding=1)
    if rec is not None:
        rec[tag + ".h"] = h.detach()
        rec[tag + ".d"] = o.detach()
        if level > 0:
            rec[tag + ".u"] = u.detach()
    return o, h


def image_encoder_block(P, p: str, x: Tensor) -> Tensor:
    """ImageEncoderConvBlock.forward (recurrent_sub_modules.py:41-49)."""
    a = _lrelu(F.conv2d(x, P[p + ".conv_1.weight"], P[p + ".conv_1.bias"], padding=1), 0.2)
    a = _lrelu(F.conv2d(a, P[p + ".conv_2.weight"], P[p + ".conv_2.bias"], padding=1), 0.2)
    a = a + F.conv2d(x, P[p + ".identity.weight"], P[p + ".identity.bias"])
    return F.conv2d(a, P[p + ".down.weight"], None, stride=2, padding=1)


def res_block(P, p: str, x: Tensor) -> Tensor:
    """ResidualBlock.forward, norm=None (recurrent_sub_modules.py:487-503)."""
    o = F.relu(F.conv2d(x, P[p + ".conv1.weight"], P[p + ".conv1.bias"], padding=1))
    o = F.conv2d(o, P[p + ".conv2.weight"], P[p + ".conv2.bias"], padding=1)
    return F.relu(o + x)


def decoder_layer(P, p: str, x: Tensor, s_prev: Optional[Tensor]) -> Tensor:
    """TransposeRecurrentConvLayer.forward (recurrent_sub_modules.py:385-408): convT 2x2 s2, then the trunk."""
    up = F.conv_transpose2d(x, P[p + ".transposed_conv2d.weight"], P[p + ".transposed_conv2d.bias"], stride=2)
    return trunk(P, p + ".forward_trunk", up, s_prev)


# ----------------------------------------------------------------------------------------------
# the network
# ----------------------------------------------------------------------------------------------
def forward(P: Dict[str, Tensor], x: Tensor, event: Tensor, rec=None) -> Tensor:
    """FinalBidirectionAttenfusion.forward (XXNet_final_attenfusion_arch.py:130-218).

    x: (B,img_chn,H,W) or (B,t,c,H,W); event: (B,T,ev_chn,H,W); returns (B,T,out_chn,H,W).
    Reproduces the state aliasing at :181 -- every forward-sweep frame is fused with the backward state
    produced by the LAST backward step (frame 0), not with its own frame's state."""
    if x.dim() == 5:
        x = x.flatten(1, 2)
    B, T = event.shape[:2]
    ev = event.flatten(0, 1)
    head = _lrelu(F.conv2d(x, P["head_img.conv2d.weight"], P["head_img.conv2d.bias"], padding=2), 0.2)
    e = _lrelu(F.conv2d(ev, P["head.conv2d.weight"], P["head.conv2d.bias"], padding=2), 0.2)
    e = e.view(B, T, *e.shape[1:])
    xb = []
    f = head
    for l in range(3):
        f = image_encoder_block(P, f"img_encoders.{l}", f)
        xb.append(f)
    if rec is not None:
        rec["head_img"] = head.detach()
        rec["e_all"] = e.detach().transpose(0, 1).flatten(0, 1)  # t-major frames, as the engine stores them
        for l in range(3):
            rec[f"xb{l}"] = xb[l].detach()

    hb: List[Optional[Tensor]] = [None, None, None]
    for t in range(T - 1, -1, -1):
        cur = e[:, t]
        for l in range(3):
            cur, hb[l] = evr_layer(P, f"encoders_backward.{l}", l, cur, xb[l - 1] if l > 0 else None, hb[l], None,
                                   rec, f"b.t{t}.l{l}")
    # hb[l] now holds the final (frame 0) backward state: the only one the forward sweep ever sees.

    hf: List[Optional[Tensor]] = [None, None, None]
    sd: List[Optional[Tensor]] = [None, None, None]
    outs = []
    for t in range(T):
        cur = e[:, t]
        skips = []
        for l in range(3):
            cur, hf[l] = evr_layer(P, f"encoders_forward.{l}", l, cur, xb[l - 1] if l > 0 else None, hf[l], hb[l],
                                   rec, f"f.t{t}.l{l}")
            skips.append(cur)
        cur = res_block(P, "resblocks.0", cur + xb[2])
        if rec is not None:
            rec[f"f.t{t}.res0"] = cur.detach()
        cur = res_block(P, "resblocks.1", cur)
        if rec is not None:
            rec[f"f.t{t}.res1"] = cur.detach()
        for i in range(3):
            cur = decoder_layer(P, f"decoders.{i}", cur + skips[2 - i], sd[i])
            sd[i] = cur
            if rec is not None:
                rec[f"f.t{t}.dec{i}.h"] = cur.detach()
        outs.append(F.conv2d(cur + head, P["pred.conv2d.weight"], P["pred.conv2d.bias"], padding=1))
    return torch.stack(outs, 1)


def charbonnier(pred: Tensor, target: Tensor, eps: float = 1e-12) -> Tensor:
    """CharbonnierLoss, reduction='mean', loss_weight 1 (basicsr/models/losses/losses.py:28-30,143-173)."""
    return torch.sqrt((pred - target) ** 2 + eps).mean()


def loss_and_grads(P: Dict[str, Tensor], x: Tensor, event: Tensor, gt: Tensor):
    """One `optimize_parameters` gradient computation without the optimiser
    (basicsr/models/twoImage_event_recurrent_model.py:273-303): forward, Charbonnier, backward.
    Parameters whose grad is None in the reference get zeros (the wrapper's `0*sum(p.sum())`, :301)."""
    Q = {k: v.detach().clone().requires_grad_(True) for k, v in P.items()}
    out = forward(Q, x, event)
    loss = charbonnier(out, gt)
    loss.backward()
    grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in Q.items()}
    return out.detach(), loss.detach(), grads


# ----------------------------------------------------------------------------------------------
# metrics restated for the PSNR parity check
# ----------------------------------------------------------------------------------------------
def tensor2img_uint8(t: Tensor):
    """clamp[0,1] -> x255 -> round (half to even, numpy) -> uint8 for a (3,H,W) tensor, channel order kept
    (basicsr/utils/img_util.py:90-117; the RGB->BGR swap there permutes channels only).  Pinned by
    tests/golden/psnr_cases.npz (generated from the unmodified reference function)."""
    return (t.detach().float().clamp(0, 1) * 255.0).round().to(torch.uint8)


def psnr_uint8(a, b, crop_border: int = 0) -> float:
    """calculate_psnr on uint8 (C,H,W) images: float64 MSE inside the crop border, `inf` for identical images, and the
    reference's peak rule `max_value = 1 if img1.max() <= 1 else 255` (basicsr/metrics/psnr_ssim.py:47-61).  Pinned by
    tests/golden/psnr_cases.npz."""
    a, b = a.double(), b.double()
    if crop_border:
        a = a[..., crop_border:-crop_border, crop_border:-crop_border]
        b = b[..., crop_border:-crop_border, crop_border:-crop_border]
    mse = ((a - b) ** 2).mean().item()
    if mse == 0:
        return float("inf")
    max_value = 1.0 if a.max().item() <= 1 else 255.0
    return 20.0 * math.log10(max_value / math.sqrt(mse))
