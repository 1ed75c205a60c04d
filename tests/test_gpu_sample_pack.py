"""GPU: training-sample assembly bit-exact against vectors built with the unmodified reference
`triple_random_crop` / `augment` / `img2tensor` and the dataset's packing lines, and the event -> voxel -> packed sample
chain at the GoPro 11+1 shape."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sample_pack_cases.npz")


def _rgb_chw(a):  # (N,H,W,3) BGR -> (N,3,H,W) RGB, what img2tensor does before the network sees an image
    return torch.from_numpy(np.ascontiguousarray(a[..., ::-1].transpose(0, 3, 1, 2)))


def test_pack_bit_exact_against_reference_vectors():
    from refid_b200 import sample_pack
    z = np.load(GOLD)
    for n in sorted({k.split(".")[0] for k in z.files}):
        m, nn, gs, seed, top, left, hf, vf, rt = (int(v) for v in z[n + ".cfg"])
        r = sample_pack.pack_blurry_sample(_rgb_chw(z[n + ".lqs"]).cuda(), torch.from_numpy(z[n + ".voxel"].transpose(2, 0, 1).copy()).cuda(),
                                           _rgb_chw(z[n + ".gts"]).cuda(), m, nn, None if gs < 0 else gs, top, left, bool(hf),
                                           bool(vf), bool(rt))
        assert np.array_equal(r["lq"].cpu().numpy(), z[n + ".lq"]), n
        assert np.array_equal(r["voxel"].cpu().numpy(), z[n + ".vox"]), n
        assert np.array_equal(r["gt"].cpu().numpy(), z[n + ".gt"]), n


def test_events_to_packed_sample_feeds_the_network_shapes():
    """events (GPU) -> 24-bin voxel grid -> crop 256 + flips -> lq (26,256,256), voxel (23,2,256,256): the shapes of
    BASELINE.json configs[1], checked against the two oracles chained on the CPU."""
    from oracle import event_oracle as E
    from oracle import sample_oracle as S
    from refid_b200 import event_util, sample_pack
    H, W, m, n = 288, 320, 11, 1
    ev = E.synthetic_events(200000, W, H, seed=3)
    vox = event_util.events_to_voxel_grid(torch.from_numpy(ev).cuda(), 2 * m + n + 1, W, H, "CHW")
    rng = np.random.RandomState(0)
    lqs = rng.rand(2, H, W, 3).astype(np.float32)
    gts = rng.rand(2 * m + n, H, W, 3).astype(np.float32)
    r = sample_pack.pack_blurry_sample(_rgb_chw(lqs).cuda(), vox, _rgb_chw(gts).cuda(), m, n, 256, 17, 40, True, False, True)
    assert r["lq"].shape == (26, 256, 256) and r["voxel"].shape == (23, 2, 256, 256) and r["gt"].shape == (23, 3, 256, 256)
    ref_vox = E.events_to_voxel_grid(ev, 2 * m + n + 1, W, H, "HWC")
    ref = S.assemble(list(lqs), list(gts), ref_vox, m, n, 256, 17, 40, True, False, True)
    assert np.array_equal(r["gt"].cpu().numpy(), ref["gt"])
    assert np.array_equal(r["lq"].cpu().numpy()[[0, 1, 2, 13, 14, 15]], ref["lq"][[0, 1, 2, 13, 14, 15]])
    assert np.abs(r["voxel"].cpu().numpy() - ref["voxel"]).max() < 1e-4  # the rasterisation sums in a different order
    with pytest.raises(RuntimeError, match="no CPU path"):
        sample_pack.pack_blurry_sample(_rgb_chw(lqs), vox.cpu(), _rgb_chw(gts), m, n)
