"""GPU event rasterisation against the reference's golden voxel grids and the numpy oracle."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("case", ["events_2bin_64x48", "events_5bin_40x40", "events_same_stamp", "events_dense_pixel"])
def test_voxel_grid_matches_reference_golden(case):
    from oracle import event_oracle as E
    from refid_b200 import event_util
    z = np.load(os.path.join(GOLD, case + ".npz"))
    n, bins, w, h, seed, srt = [int(v) for v in z["meta"]]
    ev = E.synthetic_events(n, w, h, seed, bool(srt))
    if case == "events_same_stamp":
        ev[:, 0] = 7.0
    out = event_util.events_to_voxel_grid(torch.from_numpy(ev).cuda(), bins, w, h)
    ref = z["voxel"]
    # the reference sums sequentially in float32, the kernel exactly in fixed point: they differ by float32 rounding only
    tol = 1e-5 * max(1.0, float(np.abs(ref).max()))  # ~sqrt(events per cell) float32 roundings of the running sum
    assert out.shape == ref.shape and np.abs(out.cpu().numpy() - ref).max() <= tol
    hwc = event_util.events_to_voxel_grid(torch.from_numpy(ev).cuda(), bins, w, h, "HWC")
    assert torch.equal(hwc, out.permute(1, 2, 0))
    again = event_util.events_to_voxel_grid(torch.from_numpy(ev).cuda(), bins, w, h)
    assert torch.equal(again, out)  # bit-reproducible despite the atomics


def test_sliding_windows_full_frame_and_unsorted_events():
    """1280x720, 300k events per chunk, three chunks (two windows) with unsorted stamps inside a chunk: time span from
    the first / last ROW as in the reference; the sum over the grid equals the sum of the valid weights."""
    from oracle import event_oracle as E
    from refid_b200 import event_util
    w, h = 1280, 720
    chunks = [E.synthetic_events(300000, w, h, 30 + i, sorted_time=(i != 1)) for i in range(3)]
    for i, c in enumerate(chunks):
        c[:, 0] += i
    ref = E.sliding_two_bin_voxels(chunks, w, h)
    mine = event_util.sliding_two_bin_voxels([torch.from_numpy(c).cuda() for c in chunks], w, h)
    assert len(mine) == 2
    for a, b in zip(mine, ref):
        assert a.shape == (h, w, 2) and np.abs(a.cpu().numpy() - b).max() <= 1e-5


def test_voxel_errors():
    from refid_b200 import event_util
    with pytest.raises(RuntimeError):
        event_util.events_to_voxel_grid(torch.zeros(4, 4), 2, 8, 8)
