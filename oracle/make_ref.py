"""Recipe for oracle/_ref: the UNMODIFIED reference files of the hot path, staged next to the oracle so they travel to
the GPU box like the built .so (oracle/_ref/ is git-ignored: nothing of the reference enters the repository's history).

TEST / BENCH INFRASTRUCTURE.  `python -m oracle.make_ref` (also run by __graft_entry__.build()) copies, byte for byte,
from /root/reference (present in the build container only):

    basicsr/models/archs/XXNet_final_attenfusion_arch.py   the network (FinalBidirectionAttenfusion)
    basicsr/models/archs/recurrent_sub_modules.py          EvR blocks, encoder / decoder / residual blocks
    basicsr/models/archs/fusion_modules.py                 EGACA, LayerNorm2d
    basicsr/models/archs/dcn_util.py                       imported by the network file (unused by this configuration)
    basicsr/models/archs/__init__.py                       the `*_arch.py` scan + define_network (boundary test)

The reference is a pure-Python project: there is nothing to compile.  oracle/ref_loader.py imports these files through
stub packages (the reference's own package __init__ chain needs lmdb / timm / skimage, absent here).  Consumers: the
golden-vector generators, `bench.py --impl reference` (CPU) / `--impl reference-cuda`, tests/test_cpu_reference_boundary.py.
"""
import filecmp
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("REFID_REFERENCE", "/root/reference")
DST = os.path.join(HERE, "_ref")
FILES = ["basicsr/models/archs/XXNet_final_attenfusion_arch.py", "basicsr/models/archs/recurrent_sub_modules.py",
         "basicsr/models/archs/fusion_modules.py", "basicsr/models/archs/dcn_util.py", "basicsr/models/archs/__init__.py"]


def stage(verbose=False):
    """Copy the files if the reference tree is present; returns the list of staged paths (possibly from an earlier run)."""
    if os.path.isfile(os.path.join(SRC, FILES[0])):
        for rel in FILES:
            dst = os.path.join(DST, rel)
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            if not (os.path.exists(dst) and filecmp.cmp(os.path.join(SRC, rel), dst, shallow=False)):
                shutil.copyfile(os.path.join(SRC, rel), dst)
                if verbose:
                    print("staged", rel)
    return [os.path.join(DST, rel) for rel in FILES if os.path.exists(os.path.join(DST, rel))]


if __name__ == "__main__":
    print("\n".join(stage(verbose=True)))
