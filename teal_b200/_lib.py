"""ctypes binding of libteal_b200.so -- the only way Python reaches the CUDA kernels.

There is no CPU fallback: if the shared library is missing or was not built for sm_100a this
module raises at import of the symbol table, and every op raises at call time.
"""
from __future__ import annotations

import ctypes
import os
import re
from typing import Dict, List

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libteal_b200.so")
HEADER = os.path.join(HERE, "..", "include", "teal_b200.h")

_c = ctypes
_SIGS = {
    "teal_abi_version": (_c.c_int, []),
    "teal_last_error": (_c.c_char_p, []),
    "teal_device_sm_count": (_c.c_int, []),
    "teal_compiled_arch": (_c.c_int, []),
    "teal_workspace_bytes": (_c.c_size_t, [_c.c_int]),
    "teal_workspace_init": (_c.c_int, [_c.c_void_p, _c.c_size_t, _c.c_void_p]),
    "teal_sparse_gemv_f16": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_int, _c.c_int, _c.c_int64,
                                        _c.c_double, _c.c_void_p, _c.c_size_t, _c.c_void_p]),
    "teal_sparse_qkv_gemv_f16": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_int, _c.c_int, _c.c_int64,
                                            _c.c_int, _c.c_double, _c.c_double, _c.c_double, _c.c_void_p,
                                            _c.c_size_t, _c.c_void_p]),
    "teal_sparse_gate_up_f16": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_int, _c.c_int,
                                           _c.c_int64, _c.c_int64, _c.c_double, _c.c_double, _c.c_void_p,
                                           _c.c_size_t, _c.c_void_p]),
    "teal_norm_sparse_qkv_gemv_f16": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_double, _c.c_void_p, _c.c_void_p, _c.c_void_p,
                                                 _c.c_void_p, _c.c_int, _c.c_int, _c.c_int64, _c.c_int, _c.c_double,
                                                 _c.c_double, _c.c_double, _c.c_void_p, _c.c_size_t, _c.c_void_p]),
    "teal_norm_sparse_gate_up_f16": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_double, _c.c_void_p, _c.c_void_p, _c.c_void_p,
                                                _c.c_void_p, _c.c_void_p, _c.c_int, _c.c_int, _c.c_int64, _c.c_int64,
                                                _c.c_double, _c.c_double, _c.c_void_p, _c.c_size_t, _c.c_void_p]),
    "teal_sparse_gemv_residual_f16": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_int,
                                                 _c.c_int, _c.c_int64, _c.c_double, _c.c_void_p, _c.c_size_t,
                                                 _c.c_void_p]),
    "teal_embed_f16": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_int, _c.c_void_p]),
    "teal_attn_decode_f16": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p,
                                        _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_void_p]),
    "teal_dense_gemv_f16": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_double, _c.c_void_p, _c.c_void_p, _c.c_void_p,
                                       _c.c_int, _c.c_int, _c.c_int64, _c.c_void_p]),
    "teal_residual_add_f16": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_int, _c.c_void_p]),
    "teal_tp_allreduce_residual_f16": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_int, _c.c_int, _c.c_void_p, _c.c_void_p,
                                                 _c.c_void_p, _c.c_int, _c.c_void_p]),
    "teal_sample_topk_f16": (_c.c_int, [_c.c_void_p, _c.c_int, _c.c_double, _c.c_int, _c.c_void_p, _c.c_uint64, _c.c_void_p, _c.c_void_p,
                                        _c.c_void_p, _c.c_void_p]),
    # (x | gate_up_acc, wt, resid_x, ssq_out, Z, N, ldw, thr, clear, clear_n, peers, epochs, rank, world, ws, ws_bytes, stream)
    "teal_sparse_gemv_tp_exchange_f16": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_int, _c.c_int,
                                                    _c.c_int64, _c.c_double, _c.c_void_p, _c.c_int, _c.c_void_p, _c.c_void_p, _c.c_int,
                                                    _c.c_int, _c.c_void_p, _c.c_size_t, _c.c_void_p]),
    # deferred epilogue (include/teal_b200.h): the consumer kernel rounds the producer's fp32 sums
    "teal_norm_sparse_qkv_gemv_acc_f16": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_double, _c.c_void_p, _c.c_void_p, _c.c_void_p,
                                                     _c.c_int, _c.c_int, _c.c_int64, _c.c_int, _c.c_double, _c.c_double, _c.c_double,
                                                     _c.c_void_p]),
    "teal_norm_sparse_gate_up_acc_f16": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_double, _c.c_void_p, _c.c_void_p, _c.c_void_p,
                                                    _c.c_void_p, _c.c_int, _c.c_int, _c.c_int64, _c.c_int64, _c.c_double, _c.c_double,
                                                    _c.c_void_p]),
    "teal_attn_decode_acc_f16": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p,
                                            _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_void_p]),
    "teal_sparse_gemv_residual_ex_f16": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p,
                                                    _c.c_int, _c.c_int, _c.c_int64, _c.c_double, _c.c_void_p, _c.c_int, _c.c_void_p,
                                                    _c.c_size_t, _c.c_void_p]),
    "teal_advance": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_int, _c.c_void_p]),
    "teal_survivors_f16": (_c.c_int, [_c.c_void_p, _c.c_int, _c.c_double, _c.c_void_p, _c.c_void_p, _c.c_void_p]),
    "teal_sparse_gemv_f16_simple": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_int, _c.c_int,
                                               _c.c_int64, _c.c_double, _c.c_void_p]),
    "teal_set_config": (_c.c_int, [_c.c_int]),
    "teal_set_pdl": (_c.c_int, [_c.c_int]),
    "teal_set_bytes_probe": (_c.c_int, [_c.c_void_p]),
    "teal_set_debug_timeline": (_c.c_int, [_c.c_void_p, _c.c_int]),
}



_lib = None


class TealLibraryError(RuntimeError):
    pass


def declared_symbols() -> List[str]:
    """Every function include/teal_b200.h declares (used by the CPU-side export test)."""
    with open(HEADER) as f:
        text = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    return sorted(set(re.findall(r"\b(teal_[a-z0-9_]+)\s*\(", text)))


def register_signatures(sigs: Dict[str, tuple]) -> None:
    _SIGS.update(sigs)
    if _lib is not None:
        _bind(_lib)


def _bind(lib) -> None:
    for name, (res, args) in _SIGS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args


def header_abi_version() -> int:
    """``TEAL_B200_ABI_VERSION`` as written in include/teal_b200.h (the loaded library must agree)."""
    with open(HEADER) as f:
        return int(re.search(r"#define\s+TEAL_B200_ABI_VERSION\s+(\d+)", f.read()).group(1))


def load():
    """Loads (building in-tree first when nvcc is there and the sources are newer) and returns the ctypes library."""
    global _lib
    if _lib is not None:
        return _lib
    from . import build as _build
    if _build.needs_build():
        try:
            _build.build()
        except RuntimeError as e:  # no nvcc on this box: a stale library must not be used silently
            if os.path.exists(LIB_PATH):
                raise TealLibraryError(f"{LIB_PATH} is older than its sources and cannot be rebuilt here: {e}") from e
            raise TealLibraryError(f"{LIB_PATH} is missing and cannot be built here: {e}") from e
    try:
        lib = ctypes.CDLL(LIB_PATH)
    except OSError as e:  # fail loudly: there is no fallback path
        raise TealLibraryError(f"cannot load {LIB_PATH}: {e}") from e
    _bind(lib)
    if lib.teal_abi_version() != header_abi_version():
        raise TealLibraryError(f"libteal_b200 ABI version {lib.teal_abi_version()} != header {header_abi_version()}")
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().teal_last_error().decode()
        if rc in (-1, -2):
            raise AssertionError(f"{what}: {msg}")  # reference raises AssertionError on shape/layout
        raise TealLibraryError(f"{what} failed ({rc}): {msg}")
