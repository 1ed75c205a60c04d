"""TEST INFRASTRUCTURE ONLY (imported by tests/ and nothing else).  CPU restatement, in numpy, of the reference's event
rasterisation `events_to_voxel_grid` (basicsr/data/event_util.py:6-66) and of the sliding two-bin windows built from it
(basicsr/data/image_npy_dataset.py:175-188).  Pinned by tests/golden/events_*.npz, generated from the unmodified
reference function by tests/golden/make_event_golden.py (tests/test_oracle_golden.py replays them).

Arithmetic, in the reference's order and types (its input rows are float32 [timestamp, x, y, polarity]):
  ts   = (num_bins - 1) * (t - t[0]) / (t[-1] - t[0])        float32; a zero span is replaced by 1.0        (:27-36)
  pol  = -1 where the stored polarity is 0                                                                   (:40-41)
  ti   = trunc(ts) ; dt = ts - ti (float64: float32 minus int64 promotes)                                    (:43-44)
  grid[ti][y][x]   += pol * (1 - dt)   if ti     < num_bins                                                  (:48-55)
  grid[ti+1][y][x] += pol * dt         if ti + 1 < num_bins                                                  (:57-59)
accumulated event by event into a float32 grid.
"""
import numpy as np


def events_to_voxel_grid(events: np.ndarray, num_bins: int, width: int, height: int, return_format: str = "CHW") -> np.ndarray:
    ev = np.asarray(events, dtype=np.float32)
    assert ev.ndim == 2 and ev.shape[1] == 4 and num_bins > 0 and width > 0 and height > 0
    t = ev[:, 0]
    first, last = t[0], t[-1]
    span = last - first
    if span == 0:
        span = np.float32(1.0)
    ts = (np.float32(num_bins - 1) * (t - first) / span).astype(np.float32)
    xs = ev[:, 1].astype(np.int64)
    ys = ev[:, 2].astype(np.int64)
    pol = np.where(ev[:, 3] == 0, np.float32(-1.0), ev[:, 3]).astype(np.float32)
    ti = ts.astype(np.int64)
    dt = ts.astype(np.float64) - ti
    grid = np.zeros(num_bins * height * width, np.float32)
    left = pol * (1.0 - dt)
    right = pol * dt
    ok = ti < num_bins
    np.add.at(grid, (xs + ys * width + ti * width * height)[ok], left[ok])
    ok = (ti + 1) < num_bins
    np.add.at(grid, (xs + ys * width + (ti + 1) * width * height)[ok], right[ok])
    grid = grid.reshape(num_bins, height, width)
    return grid if return_format == "CHW" else grid.transpose(1, 2, 0)


def sliding_two_bin_voxels(event_chunks, width: int, height: int):
    """One 2-bin voxel per consecutive pair of event chunks (image_npy_dataset.py:175-188), each (H, W, 2)."""
    out = []
    for a, b in zip(event_chunks[:-1], event_chunks[1:]):
        out.append(events_to_voxel_grid(np.concatenate((a, b), axis=0), 2, width, height, "HWC"))
    return out


def synthetic_events(n: int, width: int, height: int, seed: int, sorted_time: bool = True) -> np.ndarray:
    """float32 [n,4] rows [timestamp, x, y, polarity in {0,1}] -- deterministic test input, not a reference function."""
    rng = np.random.RandomState(seed)
    t = rng.uniform(10.0, 11.0, n).astype(np.float32)
    if sorted_time:
        t = np.sort(t)
    x = rng.randint(0, width, n).astype(np.float32)
    y = rng.randint(0, height, n).astype(np.float32)
    p = rng.randint(0, 2, n).astype(np.float32)
    return np.stack((t, x, y, p), axis=1)
