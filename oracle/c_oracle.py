"""ctypes loader of the plain-C oracle restatement (oracle/teal_oracle_c.c).  TEST INFRASTRUCTURE ONLY."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_build", "libteal_oracle_c.so")


def build() -> str:
    src = os.path.join(HERE, "teal_oracle_c.c")
    if not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(SO), exist_ok=True)
        subprocess.run(["gcc", "-O2", "-std=c99", "-shared", "-fPIC", "-o", SO, src, "-lm"], check=True)
    return SO


def load():
    lib = ctypes.CDLL(build())
    lib.teal_ref_round_threshold.restype = ctypes.c_double
    lib.teal_ref_round_threshold.argtypes = [ctypes.c_double]
    lib.teal_ref_survivors.restype = ctypes.c_int
    lib.teal_ref_survivors.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_double, ctypes.c_void_p]
    lib.teal_ref_sparse_gemv.restype = None
    lib.teal_ref_sparse_gemv.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_longlong,
                                         ctypes.c_double, ctypes.c_void_p, ctypes.c_void_p]
    lib.teal_ref_h2f.restype = ctypes.c_float
    lib.teal_ref_h2f.argtypes = [ctypes.c_uint16]
    lib.teal_ref_f2h.restype = ctypes.c_uint16
    lib.teal_ref_f2h.argtypes = [ctypes.c_double]
    return lib


def survivors(x16: np.ndarray, thr: float) -> np.ndarray:
    lib = load()
    x = np.ascontiguousarray(x16.reshape(-1).view(np.uint16))
    idx = np.empty(x.size, dtype=np.int32)
    n = lib.teal_ref_survivors(x.ctypes.data, x.size, float(thr), idx.ctypes.data)
    return idx[:n]


def sparse_gemv(x16: np.ndarray, w16_nz: np.ndarray, thr: float):
    """x16 fp16 [Z], w16_nz fp16 logical [N, Z]; returns (y fp16 [N], y32 fp32 [N])."""
    lib = load()
    N, Z = w16_nz.shape
    wt = np.ascontiguousarray(w16_nz.T).view(np.uint16)       # [Z][N] storage, like generate.py:296-317
    x = np.ascontiguousarray(x16.reshape(-1)).view(np.uint16)
    y = np.empty(N, dtype=np.uint16)
    y32 = np.empty(N, dtype=np.float32)
    lib.teal_ref_sparse_gemv(x.ctypes.data, wt.ctypes.data, Z, N, N, float(thr), y.ctypes.data, y32.ctypes.data)
    return y.view(np.float16), y32
