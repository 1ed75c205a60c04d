"""Test-side restatement of refid_flat_gather's rule (include/refid_b200.h): used on CPU to check the host logic that
builds the table (`FinalBidirectionAttenfusion._flat_inputs`) and on the GPU as the checker of the kernel.  Not product
code: the product path is the CUDA gather / scatter."""
import torch


def assemble(flat_floats, ins, ent):
    """flat[off + (t*R + r)*Cc + c] = src[(c*R + r)*taps + t] (mode 1) or a plain copy (mode 0); gaps are zero.
    Differentiable torch ops, so the folds behind the mode-0 inputs can be followed by autograd."""
    pieces, pos = [], 0
    dev = ins[0].device
    for t, (off, mode, taps, R, Cc) in sorted(zip(ins, ent), key=lambda p: p[1][0]):
        assert off >= pos, "entries overlap"
        assert t.numel() == taps * R * Cc, (tuple(t.shape), taps, R, Cc)
        if off > pos:
            pieces.append(torch.zeros(off - pos, device=dev))
        g = t.reshape(Cc, R, taps).permute(2, 1, 0) if mode == 1 else t
        pieces.append(g.reshape(-1).float())
        pos = off + t.numel()
    assert pos <= flat_floats
    if pos < flat_floats:
        pieces.append(torch.zeros(flat_floats - pos, device=dev))
    return torch.cat(pieces)
