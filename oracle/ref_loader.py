"""Import the UNMODIFIED reference arch files: from /root/reference in the build container, else from the byte-for-byte
staged copies under oracle/_ref/ (oracle/make_ref.py; git-ignored, shipped to the GPU box like the built .so).

TEST / BENCH INFRASTRUCTURE.  Used by tests/golden/make_golden.py, the live-reference validation of the oracle, and the
reference arms of bench.py (`--impl reference` on the host cores, `--impl reference-cuda` on the GPU).  Never imported by
refid_b200/.

The reference package's own __init__ chain needs lmdb/timm/skimage and a missing h5_image_dataset.py
(SURVEY.md section 0), so stub packages are pre-seeded in sys.modules and only the four arch files
(XXNet_final_attenfusion_arch.py, recurrent_sub_modules.py, fusion_modules.py, dcn_util.py) execute.
"""
import importlib
import logging
import os
import sys
import types

_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def _pick():
    for cand in (os.environ.get("REFID_REFERENCE"), "/root/reference", _STAGED):
        if cand and os.path.isfile(f"{cand}/basicsr/models/archs/XXNet_final_attenfusion_arch.py"):
            return cand
    return "/root/reference"


REF = _pick()


def available() -> bool:
    return os.path.isfile(f"{REF}/basicsr/models/archs/XXNet_final_attenfusion_arch.py")


def source() -> str:
    """Where the reference files are read from ('live' tree or the 'staged' byte-for-byte copies)."""
    return "staged oracle/_ref" if os.path.abspath(REF) == os.path.abspath(_STAGED) else f"live {REF}"


def load():
    sys.dont_write_bytecode = True  # the reference mount is read-only

    def _pkg(name, path):
        m = types.ModuleType(name)
        m.__path__ = [path]
        sys.modules[name] = m

    _pkg("basicsr", f"{REF}/basicsr")
    _pkg("basicsr.models", f"{REF}/basicsr/models")
    _pkg("basicsr.models.archs", f"{REF}/basicsr/models/archs")
    u = types.ModuleType("basicsr.utils")
    u.get_root_logger = lambda *a, **k: logging.getLogger("basicsr")
    sys.modules["basicsr.utils"] = u
    return importlib.import_module("basicsr.models.archs.XXNet_final_attenfusion_arch")


def build(img_chn, ev_chn):
    import contextlib
    import io

    ref = load()
    with contextlib.redirect_stdout(io.StringIO()):
        net = ref.FinalBidirectionAttenfusion(img_chn=img_chn, ev_chn=ev_chn, num_encoders=3, base_num_channels=32,
                                              num_block=1, num_residual_blocks=2)
    return net
