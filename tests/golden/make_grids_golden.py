"""Golden vectors for the `grids` crop / merge of full-frame validation, from the UNMODIFIED reference methods
(build container only):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_grids_golden.py

`TwoImageEventRecurrentRestorationModel.grids` / `.grids_inverse` / `.transpose` / `.transpose_inverse`
(basicsr/models/twoImage_event_recurrent_model.py:115-126,201-270) are called as unbound functions on a bare object that
carries `opt`, `lq`, `device`; the model file is imported through stub packages (its package __init__ chains need lmdb /
timm / skimage).  The reference's methods only handle 4-D tensors (its `grids_voxel` cannot take this network's 5-D
voxel), so the vectors pin the crop placement, the 8 transposes and the overlap averaging on 4-D frames; refid_b200.grids
applies the same placement to the trailing two dimensions of 5-D tensors.  -> tests/golden/grids_cases.npz
"""
import importlib
import logging
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = os.environ.get("REFID_REFERENCE", "/root/reference")


def load_model_class():
    sys.dont_write_bytecode = True

    def pkg(name, path=None, **attrs):
        m = types.ModuleType(name)
        if path:
            m.__path__ = [path]
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m

    pkg("basicsr", f"{REF}/basicsr")
    pkg("basicsr.models", f"{REF}/basicsr/models")
    pkg("basicsr.models.archs", None, define_network=lambda opt: None)
    pkg("basicsr.models.losses")
    pkg("basicsr.metrics")
    pkg("basicsr.utils", f"{REF}/basicsr/utils", get_root_logger=lambda *a, **k: logging.getLogger("basicsr"),
        imwrite=None, tensor2img=None)
    mod = importlib.import_module("basicsr.models.twoImage_event_recurrent_model")
    return mod.TwoImageEventRecurrentRestorationModel


CASES = {
    # name: (C, H, W, crop_size, trans_num)
    "one_crop": (3, 32, 32, 32, 1),
    "overlap_2x3": (4, 40, 72, 32, 1),
    "ragged_8trans": (2, 48, 40, 32, 8),
    "wide_720p_like": (1, 45, 80, 16, 1),
}

if __name__ == "__main__":
    M = load_model_class()
    rec = {}
    for name, (C, H, W, cs, tn) in CASES.items():
        g = torch.Generator().manual_seed(11)
        o = types.SimpleNamespace()
        o.opt = {"val": {"crop_size": cs, "trans_num": tn}}
        o.device = torch.device("cpu")
        o.lq = torch.rand(1, C, H, W, generator=g)
        o.voxel = o.lq
        o.transpose = lambda t, k: M.transpose(o, t, k)
        o.transpose_inverse = lambda t, k: M.transpose_inverse(o, t, k)
        frame = o.lq.clone()
        M.grids(o)
        parts = o.lq.clone()
        # a per-crop "network": each crop is scaled by (1 + its index / 10), so the overlap average is non-trivial
        o.output = parts * (1.0 + torch.arange(parts.shape[0]).view(-1, 1, 1, 1) / 10.0)
        o.origin_voxel = frame
        import contextlib, io
        with contextlib.redirect_stdout(io.StringIO()):
            M.grids_inverse(o)
        rec[name + ".frame"] = frame.numpy()
        rec[name + ".parts"] = parts.numpy()
        rec[name + ".idx"] = np.array([[d["i"], d["j"], d["trans_idx"]] for d in o.idxes], dtype=np.int64)
        rec[name + ".merged"] = o.output.numpy()
        rec[name + ".cfg"] = np.array([cs, tn], dtype=np.int64)
        print(name, tuple(parts.shape), "crops", len(o.idxes))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "grids_cases.npz"), **rec)
