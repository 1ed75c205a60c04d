"""CPU: the oracle restatement against the golden vectors produced by the unmodified reference."""
import pytest
import torch

import golden_util
import paramgen
from oracle import refid_oracle as O


@pytest.mark.parametrize("case", list(golden_util.CASES))
def test_oracle_matches_reference_golden(case):
    B, T, H, W, ic, ec, x5d = golden_util.CASES[case]
    gold = golden_util.load(case)
    shapes = O.param_shapes(ic, ec)
    assert sorted(shapes) == gold["names"], "parameter inventory differs from the reference state_dict"
    P = paramgen.make_params(shapes, seed=0)
    x, ev, gt = paramgen.make_inputs(B, T, H, W, ic, ec, x5d=x5d)
    torch.set_num_threads(8)
    out, loss, grads = O.loss_and_grads(P, x, ev, gt)
    assert out.shape == gold["out"].shape
    assert (out - gold["out"]).abs().max().item() < 2e-5
    assert abs(loss.item() - gold["loss"]) < 1e-6
    assert sorted(O.dead_params(shapes)) == sorted(gold["dead"])
    bad = golden_util.check_grads(grads, gold, rtol_norm=1e-3, atol_rel_samples=1e-2)
    assert not bad, bad[:5]


def test_param_count():
    n = sum(int(torch.tensor(s).prod()) for s in O.param_shapes(26, 2).values())
    assert n == 15928355
    assert len(O.param_shapes(26, 2)) == 183


def test_aliasing_quirk_matters():
    """Fact 1: the forward sweep must see the frame-0 backward state for every frame. A 'fixed' variant
    (per-frame backward state) must give a different answer, otherwise the test inputs cannot see the quirk."""
    shapes = O.param_shapes(6, 2)
    P = paramgen.make_params(shapes, seed=0)
    x, ev, _ = paramgen.make_inputs(1, 3, 32, 32, 6, 2)
    a = O.forward(P, x, ev)
    b = O.forward(P, x, ev.flip(1))
    assert (a - b.flip(1)).abs().max() > 1e-3


def test_psnr_restatement():
    a = torch.rand(3, 16, 16)
    b = (a + 0.01).clamp(0, 1)
    p = O.psnr_uint8(O.tensor2img_uint8(a), O.tensor2img_uint8(b))
    assert 35 < p < 45


EVENT_CASES = ["events_2bin_64x48", "events_5bin_40x40", "events_same_stamp", "events_dense_pixel"]


@pytest.mark.parametrize("case", EVENT_CASES)
def test_event_oracle_matches_reference_golden(case):
    """oracle/event_oracle.py against the voxel grids the unmodified reference function produced
    (tests/golden/make_event_golden.py): bit-exact -- same float32 / float64 arithmetic in the same order."""
    import os
    import numpy as np
    from oracle import event_oracle as E
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", case + ".npz"))
    n, bins, w, h, seed, srt = [int(v) for v in z["meta"]]
    ev = E.synthetic_events(n, w, h, seed, bool(srt))
    if case == "events_same_stamp":
        ev[:, 0] = 7.0
    out = E.events_to_voxel_grid(ev, bins, w, h, "CHW")
    assert out.dtype == np.float32 and out.shape == z["voxel"].shape
    assert np.array_equal(out, z["voxel"])
    assert np.array_equal(E.events_to_voxel_grid(ev, bins, w, h, "HWC"), z["voxel"].transpose(1, 2, 0))


def test_metric_oracle_matches_reference_tensor2img_and_calculate_psnr():
    """The restated validation metric against vectors of the unmodified reference functions
    (tests/golden/make_psnr_golden.py): uint8 images bit-identical (up to the RGB->BGR channel swap), PSNR to 1e-12,
    including the crop border, round-half-to-even ties, `inf`, and the max_value = 1 rule."""
    import os
    import numpy as np
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "psnr_cases.npz"))
    names = sorted({k.split(".")[0] for k in z.files})
    assert len(names) == 6
    for n in names:
        p, q = torch.from_numpy(z[n + ".pred"]), torch.from_numpy(z[n + ".gt"])
        ip, iq = O.tensor2img_uint8(p), O.tensor2img_uint8(q)
        assert np.array_equal(ip.flip(0).permute(1, 2, 0).numpy(), z[n + ".img_pred"]), n  # (H,W,C) BGR in the reference
        assert np.array_equal(iq.flip(0).permute(1, 2, 0).numpy(), z[n + ".img_gt"]), n
        want, got = float(z[n + ".psnr"]), O.psnr_uint8(ip, iq, int(z[n + ".crop"]))
        assert (np.isinf(want) and np.isinf(got)) or abs(want - got) < 1e-12, (n, want, got)


def test_grids_oracle_and_host_placement_match_reference_methods():
    """Crop placement, the 8 orientations and the overlap average against vectors of the unmodified reference methods
    (tests/golden/make_grids_golden.py), for the numpy oracle and for the product's host-side placement list."""
    import os
    import numpy as np
    from oracle import grids_oracle as G
    from refid_b200 import grids
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "grids_cases.npz"))
    names = sorted({k.split(".")[0] for k in z.files})
    assert len(names) == 4
    for n in names:
        cs, tn = (int(v) for v in z[n + ".cfg"])
        fr = z[n + ".frame"]
        parts, idx = G.grids(fr, cs, tn)
        assert np.array_equal(np.array(idx), z[n + ".idx"]), n
        assert grids.crop_positions(fr.shape[-2], fr.shape[-1], cs, tn) == idx, n
        assert np.array_equal(parts, z[n + ".parts"]), n
        out = (parts * (1.0 + np.arange(parts.shape[0], dtype=np.float32).reshape(-1, 1, 1, 1) / 10.0)).astype(np.float32)
        assert np.array_equal(G.grids_inverse(out, idx, fr.shape[-2], fr.shape[-1]), z[n + ".merged"]), n


def test_sample_assembly_oracle_matches_reference_functions():
    """Crop / flip / transpose / deblur-voxel packing / sliding windows against vectors built with the unmodified
    reference `triple_random_crop`, `augment`, `img2tensor` (tests/golden/make_sample_golden.py); also the product's
    replay of the reference's random draws (same `random` seed => same crop window and flags)."""
    import os
    import random
    import numpy as np
    from oracle import sample_oracle as S
    from refid_b200 import sample_pack
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sample_pack_cases.npz"))
    names = sorted({k.split(".")[0] for k in z.files})
    assert len(names) == 5
    for n in names:
        m, nn, gs, seed, top, left, hf, vf, rt = (int(v) for v in z[n + ".cfg"])
        gs = None if gs < 0 else gs
        r = S.assemble(list(z[n + ".lqs"]), list(z[n + ".gts"]), z[n + ".voxel"], m, nn, gs, top, left, hf, vf, rt)
        assert np.array_equal(r["lq"], z[n + ".lq"]) and np.array_equal(r["voxel"], z[n + ".vox"]) and np.array_equal(r["gt"], z[n + ".gt"]), n
        random.seed(seed)
        H, W = z[n + ".voxel"].shape[:2]
        assert sample_pack.draw_crop_and_flips(H, W, gs) == (top, left, bool(hf), bool(vf), bool(rt)), n
