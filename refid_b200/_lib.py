"""ctypes binding of librefid_b200.so (the C ABI declared in include/refid_b200.h).

There is no fallback: if the shared library is missing or a call fails, a RuntimeError is raised."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.environ.get("REFID_LIB") or os.path.join(_HERE, "librefid_b200.so")  # REFID_LIB: A/B of two builds (diagnostic)
_lib = None

c_void_p, c_int, c_long, c_float = ctypes.c_void_p, ctypes.c_int, ctypes.c_long, ctypes.c_float


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_PATH):
            raise RuntimeError(
                f"{_PATH} not found: build it with `python -m refid_b200.build` (nvcc, sm_100a). "
                "refid_b200 has no CPU or PyTorch fallback.")
        _lib = ctypes.CDLL(_PATH)
        _lib.refid_last_error.restype = ctypes.c_char_p
    return _lib


def check(rc, what):
    if rc != 0:
        raise RuntimeError(f"{what} failed: {lib().refid_last_error().decode()}")


def ptr(t):
    return c_void_p(0 if t is None else t.data_ptr())


def set_pdl(enable):
    """Programmatic dependent launch on/off for subsequent launches; returns the previous setting."""
    return bool(lib().refid_set_pdl(1 if enable else 0))


def abort_flag():
    v = ctypes.c_uint(0)
    check(lib().refid_abort_flag(ctypes.byref(v)), "refid_abort_flag")
    return v.value


def raise_if_aborted():
    """Sticky, sync-free health check (pinned host word written by a kernel whose bounded mbarrier wait timed out)."""
    L = lib()
    L.refid_abort_pending.restype = ctypes.c_uint
    v = L.refid_abort_pending()
    if v:
        raise RuntimeError(f"refid_b200: a kernel hit its bounded mbarrier wait (code 0x{v:x}); results after it are "
                           "invalid. refid_b200._lib.abort_flag() reads and clears the flag.")
