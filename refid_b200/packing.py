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


def pack_down_dgrad_halo(w):
    """Data-gradient of the 4x4 stride-2 pad-1 conv on the halo-conv engine (csrc/convop.cuh CK_DOWN4_DGRAD_HALO):
    rows = (tap9*4 + q)*Cin + ci, cols = co, where tap9 = (dy+1)*3 + (dx+1) indexes the 3x3 neighbourhood of dY and
    q = qy*2+qx the parity of the dX pixel (2i+qy, 2j+qx); value = W[co,ci,ky,kx] with ky = qy + 1 - 2*dy, kx = qx + 1 - 2*dx
    when inside the kernel, else 0."""
    co, ci, _, _ = w.shape
    out = torch.zeros(9, 4, ci, co, dtype=w.dtype, device=w.device)
    for t9 in range(9):
        dy, dx = t9 // 3 - 1, t9 % 3 - 1
        for q in range(4):
            ky, kx = (q >> 1) + 1 - 2 * dy, (q & 1) + 1 - 2 * dx
            if 0 <= ky < 4 and 0 <= kx < 4:
                out[t9, q] = w[:, :, ky, kx].t()
    return out.reshape(36 * ci, co).contiguous()


def pack_down_fwd_halo(w):
    """Forward of the 4x4 stride-2 pad-1 conv on the halo-conv engine (csrc/convop.cuh CK_DOWN4_HALO): a 3x3 conv over the
    four stride-2 parity views q = (py,px) of the input, K = 4*Cin: rows = tap9*Cout + co, cols = q*Cin + ci,
    value = W[co,ci,ky,kx] with ky = 2*dy + py + 1, kx = 2*dx + px + 1 when inside the kernel, else 0."""
    co, ci, _, _ = w.shape
    out = torch.zeros(9, co, 4, ci, dtype=w.dtype, device=w.device)
    for t9 in range(9):
        dy, dx = t9 // 3 - 1, t9 % 3 - 1
        for q in range(4):
            ky, kx = 2 * dy + (q >> 1) + 1, 2 * dx + (q & 1) + 1
            if 0 <= ky < 4 and 0 <= kx < 4:
                out[t9, :, q, :] = w[:, :, ky, kx]
    return out.reshape(9 * co, 4 * ci).contiguous()
