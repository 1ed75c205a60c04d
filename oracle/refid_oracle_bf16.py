"""CPU oracle with the engine's bf16 storage points emulated (TEST INFRASTRUCTURE -- never imported by the product).

`refid_oracle.py` is the fp32 restatement of the reference and is pinned to golden vectors of the reference itself.
The CUDA engine stores every activation (and every activation gradient) in bf16 and multiplies bf16 weights, so against
the fp32 oracle it can only be checked to bf16 noise (a few 1e-3 on outputs, several % on individual gradient tensors
after ~100 layers of recurrence).  This file restates the SAME network (same reference lines; see refid_oracle.py for
the citations) with a round-to-bf16 at exactly the places where the engine writes bf16 -- forward values and, through a
custom autograd function, the gradients flowing back through those tensors -- and with the engine's algebraic folds
(LayerNorm affine into the next 1x1 conv, beta into conv3, gamma*conv5 concatenated with conv_y_side, both level-0
in-convs stacked).  tests/ check (a) this file against refid_oracle.py on CPU (bf16-noise tolerance) and (b) the CUDA
engine against this file (tight tolerance), which is what proves the backward pass term by term.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


class _Q(torch.autograd.Function):
    """bf16 round trip on the value and on the gradient (a tensor the engine stores in bf16, forward and backward)."""

    @staticmethod
    def forward(ctx, t):
        return t.to(torch.bfloat16).to(torch.float32)

    @staticmethod
    def backward(ctx, g):
        return g.to(torch.bfloat16).to(torch.float32)


def q(t: Tensor) -> Tensor:
    return _Q.apply(t)


def qw(w: Tensor) -> Tensor:
    """bf16-rounded weight as the GEMMs see it; gradient passes straight to the fp32 master."""
    return w + (w.to(torch.bfloat16).to(torch.float32) - w).detach()


def _lrelu(x, s):
    return F.leaky_relu(x, s)


def lnhat(x: Tensor, eps: float = 1e-6) -> Tensor:
    mu = x.mean(1, keepdim=True)
    var = (x - mu).pow(2).mean(1, keepdim=True)
    return (x - mu) / (var + eps).sqrt()


def _fold_ln(P, conv: str, norm: str):
    w, b = P[conv + ".weight"], P[conv + ".bias"]
    return w * P[norm + ".weight"].view(1, -1, 1, 1), b + w[:, :, 0, 0] @ P[norm + ".bias"]


def egaca_image(P, p: str, xi: Tensor) -> Tensor:
    w, b = _fold_ln(P, p + ".conv1", p + ".norm1")
    a = q(F.conv2d(q(lnhat(xi)), qw(w), b))
    return q(F.gelu(F.conv2d(a, P[p + ".conv2.weight"], P[p + ".conv2.bias"], padding=1, groups=a.shape[1])))


def egaca_step(P, p: str, xe: Tensor, xi: Tensor, g_i: Tensor) -> Tensor:
    w, b = _fold_ln(P, p + ".conv1_e", p + ".norm1_e")
    a = q(F.conv2d(q(lnhat(xe)), qw(w), b))
    ge_full = F.gelu(F.conv2d(a, P[p + ".conv2_e.weight"], P[p + ".conv2_e.bias"], padding=1, groups=a.shape[1]))
    g_e = q(ge_full)
    pooled = ge_full.mean((2, 3), keepdim=True)
    s = F.relu(F.conv2d(pooled, P[p + ".se_1.1.weight"], P[p + ".se_1.1.bias"]))
    s = torch.sigmoid(F.conv2d(s, P[p + ".se_1.3.weight"], P[p + ".se_1.3.bias"]))
    # the gate is folded into conv3: each sample's weights are scaled by s on their K side and rounded to bf16; the gated
    # tensor itself is never stored (engine.cu: egaca_step)
    beta = P[p + ".beta"].view(-1)
    w3 = P[p + ".conv3.weight"] * beta.view(-1, 1, 1, 1)                      # (64,128,1,1)
    ws = qw(w3.unsqueeze(0) * torch.cat((s, s), 1).view(s.shape[0], 1, -1, 1, 1))  # (B,64,128,1,1)
    gcat = torch.cat((g_i.expand(g_e.shape[0], -1, -1, -1), g_e), 1)
    y3 = torch.einsum("bok,bkhw->bohw", ws[:, :, :, 0, 0], gcat) + (P[p + ".conv3.bias"] * beta).view(1, -1, 1, 1)
    y = q(y3 + xe + xi)
    w4, b4 = _fold_ln(P, p + ".conv4", p + ".norm2")
    g4 = q(F.gelu(F.conv2d(q(lnhat(y)), qw(w4), b4)))
    gamma = P[p + ".gamma"].view(-1)
    w5 = torch.cat((P[p + ".conv_y_side.weight"], P[p + ".conv5.weight"] * gamma.view(-1, 1, 1, 1)), 1)
    return q(F.conv2d(torch.cat((y, g4), 1), qw(w5), P[p + ".conv_y_side.bias"] + P[p + ".conv5.bias"] * gamma))


def trunk(P, p: str, u: Tensor, h_prev: Optional[Tensor], post: Optional[Tensor] = None):
    if h_prev is None:
        h_prev = torch.zeros_like(u)
    v = q(_lrelu(F.conv2d(torch.cat((u, h_prev), 1), qw(P[p + ".main.0.weight"]), P[p + ".main.0.bias"], padding=1), 0.1))
    r = q(F.relu(F.conv2d(v, qw(P[p + ".main.2.0.conv1.weight"]), P[p + ".main.2.0.conv1.bias"], padding=1)))
    hp = F.conv2d(r, qw(P[p + ".main.2.0.conv2.weight"]), P[p + ".main.2.0.conv2.bias"], padding=1) + v
    return q(hp), (q(hp + post) if post is not None else None)


def forward(P: Dict[str, Tensor], x: Tensor, event: Tensor) -> Tensor:
    if x.dim() == 5:
        x = x.flatten(1, 2)
    B, T = event.shape[:2]
    x = x.to(torch.bfloat16).float()
    ev = event.flatten(0, 1).to(torch.bfloat16).float()
    head = q(_lrelu(F.conv2d(x, qw(P["head_img.conv2d.weight"]), P["head_img.conv2d.bias"], padding=2), 0.2))
    xb = []
    f = head
    for l in range(3):
        p = f"img_encoders.{l}"
        t1 = q(_lrelu(F.conv2d(f, qw(P[p + ".conv_1.weight"]), P[p + ".conv_1.bias"], padding=1), 0.2))
        idn = q(F.conv2d(f, qw(P[p + ".identity.weight"]), P[p + ".identity.bias"]))
        c2 = _lrelu(F.conv2d(t1, qw(P[p + ".conv_2.weight"]), P[p + ".conv_2.bias"], padding=1), 0.2)
        f = q(F.conv2d(q(c2 + idn), qw(P[p + ".down.weight"]), None, stride=2, padding=1))
        xb.append(f)
    e = q(_lrelu(F.conv2d(ev, qw(P["head.conv2d.weight"]), P["head.conv2d.bias"], padding=2), 0.2))
    w0 = torch.cat((P["encoders_backward.0.conv.conv2d.weight"], P["encoders_forward.0.conv.conv2d.weight"]), 0)
    b0 = torch.cat((P["encoders_backward.0.conv.conv2d.bias"], P["encoders_forward.0.conv.conv2d.bias"]), 0)
    u0 = q(_lrelu(F.conv2d(e, qw(w0), b0, padding=1), 0.04)).view(B, T, 128, *e.shape[2:])
    g_i = [egaca_image(P, f"encoders_{d}.1.atten_fuse", xb[0]) for d in ("backward", "forward")]

    def level_in(d, di, l, t, cur):
        if l == 0:
            return u0[:, t, 64 * di:64 * di + 64]
        if l == 1:
            return egaca_step(P, f"encoders_{d}.1.atten_fuse", cur, xb[0], g_i[di])
        p = f"encoders_{d}.2.conv.conv2d"
        return q(_lrelu(F.conv2d(cur, qw(P[p + ".weight"]), P[p + ".bias"], padding=1), 0.04))

    hb: List[Optional[Tensor]] = [None, None, None]
    for t in range(T - 1, -1, -1):
        cur = None
        for l in range(3):
            u = level_in("backward", 0, l, t, cur)
            hb[l], _ = trunk(P, f"encoders_backward.{l}.recurrent_block.forward_trunk", u, hb[l])
            if l < 2:
                dp = F.conv2d(hb[l], qw(P[f"encoders_backward.{l}.down.weight"]), None, stride=2, padding=1)
                cur = q(dp + xb[1]) if l == 1 else q(dp)
    hf: List[Optional[Tensor]] = [None, None, None]
    sd: List[Optional[Tensor]] = [None, None, None]
    outs = []
    for t in range(T):
        cur = None
        dn = []
        for l in range(3):
            p = f"encoders_forward.{l}"
            u = level_in("forward", 1, l, t, cur)
            hf[l], _ = trunk(P, p + ".recurrent_block.forward_trunk", u, hf[l])
            hfu = q(_lrelu(F.conv2d(torch.cat((hf[l], hb[l]), 1), qw(P[p + ".fuse_two_dir.conv2d.weight"]),
                                    P[p + ".fuse_two_dir.conv2d.bias"]), 0.2))
            dp = F.conv2d(hfu, qw(P[p + ".down.weight"]), None, stride=2, padding=1)
            dn.append(q(dp))
            cur = q(dp + xb[l]) if l >= 1 else dn[-1]
        xin = cur
        for i in range(2):
            p = f"resblocks.{i}"
            r = q(F.relu(F.conv2d(xin, qw(P[p + ".conv1.weight"]), P[p + ".conv1.bias"], padding=1)))
            op = F.relu(F.conv2d(r, qw(P[p + ".conv2.weight"]), P[p + ".conv2.bias"], padding=1) + xin)
            xin = q(op + dn[2]) if i == 1 else q(op)
        for i in range(3):
            p = f"decoders.{i}"
            up = q(F.conv_transpose2d(xin, qw(P[p + ".transposed_conv2d.weight"]), P[p + ".transposed_conv2d.bias"], stride=2))
            sd[i], xin = trunk(P, p + ".forward_trunk", up, sd[i], dn[1 - i] if i < 2 else head)
        outs.append(F.conv2d(xin, qw(P["pred.conv2d.weight"]), P["pred.conv2d.bias"], padding=1))
    return torch.stack(outs, 1)


def vjp(P: Dict[str, Tensor], x: Tensor, event: Tensor, cot: Tensor):
    """out, {name: dL/dparam} for L = <out, cot> (parameters without a path to the output get zeros)."""
    Q = {k: v.detach().clone().requires_grad_(True) for k, v in P.items()}
    out = forward(Q, x, event)
    (out * cot).sum().backward()
    return out.detach(), {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in Q.items()}
