"""Packed weight layouts consumed by the tap-GEMM kernels (bf16 row-major matrices [rows][K]).

These torch restatements document the layouts and cross-check the CUDA packing kernels in tests.
tap order is row-major over (ky,kx); see csrc/convop.cu for the matching tap tables.
"""
import torch


def pack_fwd(w):
    """Conv2d weight (Cout,Cin,kh,kw) -> [kh*kw*Cout][Cin]; row = tap*Cout + co."""
    co, ci, kh, kw = w.shape
    return w.permute(2, 3, 0, 1).reshape(kh * kw * co, ci).contiguous()


def pack_dgrad_s1(w):
    """Data-gradient of a stride-1 'same' conv: rows = tap'*Cin + ci, cols = co, taps flipped."""
    co, ci, kh, kw = w.shape
    return w.flip(2, 3).permute(2, 3, 1, 0).reshape(kh * kw * ci, co).contiguous()


def pack_up2_fwd(w):
    """ConvTranspose2d weight (Cin,Cout,2,2) -> rows = (a*2+b)*Cout + co, cols = ci."""
    ci, co, _, _ = w.shape
    return w.permute(2, 3, 1, 0).reshape(4 * co, ci).contiguous()


def pack_up2_dgrad(w):
    """rows = (a*2+b)*Cin + ci, cols = co."""
    ci, co, _, _ = w.shape
    return w.permute(2, 3, 0, 1).reshape(4 * ci, co).contiguous()


_KIDX = ((1, 3), (0, 2))  # [output parity][slot] -> kernel index (csrc/convop.cu: down_dgrad_off)


def pack_down_dgrad(w):
    """Data-gradient of the 4x4 stride-2 pad-1 conv, one 4-tap block per output parity:
    rows = ((py*2+px)*4 + a*2+b)*Cin + ci, cols = co, value = W[co,ci,ky(py,a),kx(px,b)]."""
    co, ci, _, _ = w.shape
    out = torch.empty(4, 4, ci, co, dtype=w.dtype, device=w.device)
    for py in range(2):
        for px in range(2):
            for a in range(2):
                for b in range(2):
                    out[py * 2 + px, a * 2 + b] = w[:, :, _KIDX[py][a], _KIDX[px][b]].t()
    return out.reshape(16 * ci, co).contiguous()


def unpack_wgrad_conv(d, cout, cin, kh, kw):
    """wgrad output D[tap*Cin + ci][co] -> Conv2d weight gradient (Cout,Cin,kh,kw)."""
    return d.view(kh, kw, cin, cout).permute(3, 2, 0, 1).contiguous()


def unpack_wgrad_up2(d, cin, cout):
    """wgrad output D[(a*2+b)*Cout + co][ci] -> ConvTranspose2d weight gradient (Cin,Cout,2,2)."""
    return d.view(2, 2, cout, cin).permute(3, 2, 0, 1).contiguous()
