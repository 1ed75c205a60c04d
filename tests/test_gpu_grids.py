"""GPU: full-frame validation tiling -- crop / merge kernels against vectors of the unmodified
reference methods (bit-exact), against the numpy oracle on a 5-D 720p frame, and the wrapper's grids -> test ->
grids_inverse path against the oracle network run crop by crop."""
import os

import numpy as np
import pytest
import torch

import paramgen

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "grids_cases.npz")


def test_crop_and_merge_bit_exact_against_reference_vectors():
    from refid_b200 import grids
    z = np.load(GOLD)
    for n in sorted({k.split(".")[0] for k in z.files}):
        cs, tn = (int(v) for v in z[n + ".cfg"])
        fr = torch.from_numpy(z[n + ".frame"]).cuda()
        idx = grids.crop_positions(fr.shape[-2], fr.shape[-1], cs, tn)
        parts = grids.crop(fr, idx, cs)
        assert np.array_equal(parts.cpu().numpy(), z[n + ".parts"]), n
        # the golden run formed these weights on the CPU (a CUDA `tensor / 10.0` multiplies by the reciprocal: 1 ulp off)
        out = parts * (1.0 + torch.arange(parts.shape[0]).view(-1, 1, 1, 1) / 10.0).cuda()
        merged = grids.merge(out, idx, fr.shape[-2], fr.shape[-1])
        assert np.array_equal(merged.cpu().numpy(), z[n + ".merged"]), n


def test_five_dimensional_720p_voxel_against_oracle():
    """BASELINE.json configs[4] frame size: (1,T,2,720,1280) voxel, 256-pixel crops, 8 orientations on a second pass."""
    from oracle import grids_oracle as G
    from refid_b200 import grids
    v = torch.randn(1, 3, 2, 720, 1280, generator=torch.Generator().manual_seed(4))
    for tn in (1, 8):
        idx = grids.crop_positions(720, 1280, 256, tn)
        ref_parts, ref_idx = G.grids(v.numpy(), 256, tn)
        assert idx == ref_idx and len(idx) == 15 * tn
        parts = grids.crop(v.cuda(), idx, 256)
        assert parts.shape == (15 * tn, 3, 2, 256, 256) and np.array_equal(parts.cpu().numpy(), ref_parts)
        w = torch.rand(parts.shape[0], 1, 1, 1, 1, generator=torch.Generator().manual_seed(5)).cuda() + 0.5
        merged = grids.merge(parts * w, idx, 720, 1280)
        ref = G.grids_inverse((parts * w).cpu().numpy(), idx, 720, 1280)
        assert merged.shape == (1, 3, 2, 720, 1280) and np.array_equal(merged.cpu().numpy(), ref)
    # identity: un-weighted crops merge back to the frame exactly where one crop covers, to rounding where several do
    idx = grids.crop_positions(720, 1280, 256, 1)
    back = grids.merge(grids.crop(v.cuda(), idx, 256), idx, 720, 1280)
    assert (back.cpu() - v).abs().max().item() < 1e-6
    with pytest.raises(ValueError, match="does not cover"):
        grids.merge(grids.crop(v.cuda(), idx[:-1], 256), idx[:-1], 720, 1280)
    with pytest.raises(RuntimeError, match="no CPU path"):
        grids.crop(v, idx, 256)


def test_wrapper_grids_validation_matches_oracle_crop_by_crop():
    """`val.grids` validation of one 96x160 frame with 64-pixel crops and max_minibatch chunking: equals the fp32 oracle
    network applied to every crop, merged by the grids oracle (fp16 forward-only tolerance)."""
    from oracle import grids_oracle as G
    from oracle import refid_oracle as O
    from refid_b200 import recurrent_model
    T, H, W, ic, ec = 2, 96, 160, 6, 2
    opt = {"network_g": {"type": "FinalBidirectionAttenfusion", "img_chn": ic, "ev_chn": ec, "num_encoders": 3,
                         "base_num_channels": 32, "num_block": 1, "num_residual_blocks": 2},
           "val": {"grids": True, "crop_size": 64, "max_minibatch": 4}}
    P = paramgen.make_params(O.param_shapes(ic, ec), seed=0)
    x, ev, _ = paramgen.make_inputs(1, T, H, W, ic, ec, x5d=True)
    m = recurrent_model.TwoImageEventRecurrentRestorationModel(opt, device="cuda")
    m.net_g.load_state_dict(P, strict=True)
    out = m.validate_frame({"lq": x, "voxel": ev})
    assert out.shape == (1, T, 3, H, W)
    assert m.lq.shape == (1, 2, 3, H, W) and m.voxel.shape == (1, T, ec, H, W)  # restored, as grids_inverse does
    xp, idx = G.grids(x.numpy(), 64)
    vp, _ = G.grids(ev.numpy(), 64)
    assert len(idx) == 6
    with torch.no_grad():
        ref_parts = O.forward(P, torch.from_numpy(xp.copy()), torch.from_numpy(vp.copy()))
    ref = G.grids_inverse(ref_parts.numpy(), idx, H, W)
    assert np.abs(out.cpu().numpy() - ref).max() < 2e-3
