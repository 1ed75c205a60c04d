"""Build librefid_b200.so in-tree with nvcc for sm_100a (the only target).

Every csrc/*.cu is its own translation unit, compiled in parallel into refid_b200/build/*.o (git-ignored) and linked
into one shared library; a unit is recompiled only when it or a header is newer than its object."""
import concurrent.futures
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
OUT = os.path.join(HERE, "librefid_b200.so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _units():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _headers_mtime():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
    hs += [os.path.join(INCLUDE, f) for f in os.listdir(INCLUDE)]
    return max(os.path.getmtime(h) for h in hs)


def _flags():
    return os.environ.get("REFID_NVCC_FLAGS", "").split()  # diagnostic builds, e.g. -DREFID_HALO_TIMING


def _compile(unit, verbose):
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    obj = os.path.join(OBJ, unit[:-3] + ".o")
    cmd = [nvcc, *_flags(), *ARCH, "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-c", "-o", obj,
           os.path.join(CSRC, unit)]
    if verbose:
        cmd[1:1] = ["-Xptxas", "-v"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed on {unit}:\n" + r.stdout + r.stderr)
    return r.stderr


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    stamp = os.path.join(OBJ, ".flags")
    flags = " ".join(_flags())
    if not os.path.exists(stamp) or open(stamp).read() != flags:
        force = True
    hdr = _headers_mtime()
    todo = []
    for u in _units():
        obj = os.path.join(OBJ, u[:-3] + ".o")
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(hdr, os.path.getmtime(os.path.join(CSRC, u))):
            todo.append(u)
    objs = [os.path.join(OBJ, u[:-3] + ".o") for u in _units()]
    if not todo and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(o) for o in objs):
        return OUT
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(len(todo), os.cpu_count() or 4) or 1) as ex:
        for u, log in zip(todo, ex.map(lambda u: _compile(u, verbose), todo)):
            if verbose:
                print(f"---- {u}\n{log}")
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    r = subprocess.run([nvcc, *ARCH, "-shared", "-o", OUT, *objs], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    open(stamp, "w").write(flags)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
