"""ctypes view of tests/emu/_build/libteal_emu.so: the C ABI of include/teal_b200.h executed by the CPU emulation of
the kernels (tests/emu/cuda_emu.h) on numpy arrays.  Test infrastructure only -- nothing here is timed or shipped."""
from __future__ import annotations

import ctypes
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
if HERE not in sys.path:
    sys.path.insert(0, HERE)

_lib = None


def lib():
    global _lib
    if _lib is None:
        import build as emu_build                      # tests/emu/build.py
        from teal_b200 import _lib as binding         # only the signature table: the CUDA library is not loaded
        _lib = ctypes.CDLL(emu_build.build())
        for name, (res, args) in binding._SIGS.items():
            fn = getattr(_lib, name)
            fn.restype, fn.argtypes = res, args
        _lib.teal_emu_set_sm_count.argtypes = [ctypes.c_int]
        _lib.teal_emu_set_concurrent.argtypes = [ctypes.c_int]
    return _lib


def check(rc, what=""):
    if rc != 0:
        raise RuntimeError(f"{what}: rc={rc}: {lib().teal_last_error().decode()}")


def ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def aligned(shape, dtype, fill=None, align=128):
    """numpy array whose data pointer is `align`-byte aligned (the library requires 16-byte aligned weights)."""
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    raw = np.zeros(n + align, dtype=np.uint8)
    off = (-raw.ctypes.data) % align
    out = raw[off:off + n].view(dtype).reshape(shape)
    if fill is not None:
        out[...] = fill
    return out


def column_major(W):
    """[N, Z] fp16 weight -> the reference's storage [Z][N] (kernels/sparse_gemv.py:106), aligned."""
    N, Z = W.shape
    return aligned((Z, N), np.float16, W.T)


class Workspace:
    def __init__(self, out_cols=1 << 16):
        self.nbytes = lib().teal_workspace_bytes(out_cols)
        self.buf = aligned((self.nbytes,), np.uint8, 0)

    def assert_clean(self):
        """every kernel must leave counters and accumulators zeroed (no memset launch anywhere)"""
        scratch = 64 * 4                       # the trailing sum-of-squares scratch may hold values
        assert not self.buf[: self.nbytes - scratch].any(), "workspace not left zeroed"


def set_config(cfg):
    check(lib().teal_set_config(int(cfg)), "teal_set_config")


KNOBS = {"generation": 0, "spec_rows": 1, "spec_mode": 2}


def set_tuning(generation=1, spec_rows=0, spec_mode=1):
    for k, v in (("generation", generation), ("spec_rows", spec_rows), ("spec_mode", spec_mode)):
        check(lib().teal_set_tuning(KNOBS[k], int(v)), f"teal_set_tuning({k})")
