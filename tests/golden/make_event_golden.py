"""Generate golden vectors for the event rasterisation from the UNMODIFIED reference function (build container only):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_event_golden.py

basicsr/data/event_util.py is executed from /root/reference with a stub `basicsr.utils` (its Timer imports need lmdb etc.)
and with `np.int = int` restored (the reference predates numpy 1.24); inputs come from oracle.event_oracle.synthetic_events
so the fixtures only store the seed, the shape and the reference's output.
"""
import importlib.util
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import event_oracle as E  # noqa: E402

CASES = {
    # name: (n_events, num_bins, width, height, seed, sorted)
    "events_2bin_64x48": (5000, 2, 64, 48, 1, True),
    "events_5bin_40x40": (3000, 5, 40, 40, 2, True),
    "events_same_stamp": (64, 2, 16, 16, 3, True),
    "events_dense_pixel": (4000, 2, 4, 4, 4, True),
}


def load_reference():
    if not hasattr(np, "int"):
        np.int = int
    for name in ("basicsr", "basicsr.utils"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["basicsr.utils"].Timer = object
    sys.modules["basicsr.utils"].CudaTimer = object
    spec = importlib.util.spec_from_file_location("ref_event_util", "/root/reference/basicsr/data/event_util.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    ref = load_reference()
    for name, (n, bins, w, h, seed, srt) in CASES.items():
        ev = E.synthetic_events(n, w, h, seed, srt)
        if name == "events_same_stamp":
            ev[:, 0] = 7.0  # zero time span: the reference substitutes deltaT = 1
        out = ref.events_to_voxel_grid(ev.copy(), bins, w, h, "CHW")
        np.savez_compressed(os.path.join(os.path.dirname(__file__), name + ".npz"), voxel=out.astype(np.float32),
                            meta=np.array([n, bins, w, h, seed, int(srt)], np.int64))
        print(name, out.shape, float(np.abs(out).sum()))


if __name__ == "__main__":
    main()
