"""ctypes binding of libsonarfe.so (the C ABI declared in include/sonarfe.h).

Fails loudly: a missing shared object or a machine without a B200 raises -- there
is no CPU path behind these functions.
"""
import ctypes
import os
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsonarfe.so")

c_int, c_double, c_void_p, c_float = ctypes.c_int, ctypes.c_double, ctypes.c_void_p, ctypes.c_float

_lib = None
_lock = threading.Lock()


class SonarFEError(RuntimeError):
    pass


def load():
    """Load (building first if sources are newer and nvcc is present) and type the library."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            from . import build as _build
            _build.build()
        if not os.path.exists(LIB_PATH):
            raise SonarFEError(f"{LIB_PATH} is missing: build it with `python -m sonar_slam_b200.build`")
        lib = ctypes.CDLL(LIB_PATH)
        lib.sfe_version.restype = c_int
        lib.sfe_last_error.restype = ctypes.c_char_p
        lib.sfe_ctx_create.argtypes = [c_int, c_void_p, c_int, ctypes.POINTER(c_void_p)]
        lib.sfe_ctx_destroy.argtypes = [c_void_p]
        lib.sfe_ctx_destroy.restype = None
        lib.sfe_sync.argtypes = [c_void_p]
        lib.sfe_launch_count.argtypes = [c_void_p]
        lib.sfe_launch_count.restype = ctypes.c_uint64
        cfar_args = [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_double,
                     c_int, c_double, c_void_p, c_void_p]
        lib.sfe_cfar_dev.argtypes = cfar_args + [c_void_p]
        lib.sfe_cfar_host.argtypes = cfar_args
        _type_more(lib)
        _lib = lib
        return lib


def _type_more(lib):
    """argtypes of the entry points added after the CFAR milestone (kept in one place)."""
    for name, args in _EXTRA_SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = c_int


_EXTRA_SIGNATURES = {}


def check(rc, what=""):
    if rc != 0:
        msg = load().sfe_last_error().decode(errors="replace")
        raise SonarFEError(f"{what or 'libsonarfe'} failed (code {rc}): {msg}")


class Context:
    """One GPU + one stream + scratch (sfe_ctx).  Not thread-safe; make one per thread."""

    def __init__(self, device=0, stream="own"):
        """stream: "own" -> library-owned stream; otherwise a cudaStream_t integer handle of the
        caller (0 / None = the legacy default stream, e.g. torch.cuda.current_stream().cuda_stream)."""
        self.lib = load()
        h = c_void_p()
        own = 1 if stream == "own" else 0
        sp = None if (own or not stream) else c_void_p(int(stream))
        check(self.lib.sfe_ctx_create(int(device), sp, own, ctypes.byref(h)), "sfe_ctx_create")
        self.handle = h
        self.device = int(device)

    def sync(self):
        check(self.lib.sfe_sync(self.handle), "sfe_sync")

    @property
    def launches(self):
        return int(self.lib.sfe_launch_count(self.handle))

    def close(self):
        if getattr(self, "handle", None):
            self.lib.sfe_ctx_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_default_ctx = {}


def default_context(device=0):
    """Process-wide context used by the drop-in single-call API (bruce_slam.cfar / .pcl)."""
    key = (os.getpid(), threading.get_ident(), int(device))
    ctx = _default_ctx.get(key)
    if ctx is None:
        ctx = _default_ctx[key] = Context(device)
    return ctx


def ptr(a):
    """Host pointer of a C-contiguous numpy array (or None)."""
    if a is None:
        return None
    assert isinstance(a, np.ndarray) and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(c_void_p)
