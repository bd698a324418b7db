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
    "teal_norm_sparse_gate_up_chain_f16": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_double, _c.c_void_p, _c.c_void_p, _c.c_void_p,
                                                      _c.c_void_p, _c.c_int, _c.c_int, _c.c_int64, _c.c_int64, _c.c_double, _c.c_double,
                                                      _c.c_int, _c.c_void_p, _c.c_size_t, _c.c_void_p]),
    "teal_sparse_gemv_residual_chain_f16": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_int,
                                                       _c.c_int, _c.c_int64, _c.c_double, _c.c_int, _c.c_void_p, _c.c_size_t, _c.c_void_p]),
    "teal_chain_consumer_ctas": (_c.c_int, [_c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_double, _c.c_double, _c.c_double]),
    "teal_norm_sparse_qkv_gemv_chain_f16": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_double, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_int,
                                                       _c.c_int, _c.c_int64, _c.c_int, _c.c_double, _c.c_double, _c.c_double, _c.c_int,
                                                       _c.c_int, _c.c_int, _c.c_int, _c.c_void_p, _c.c_size_t, _c.c_void_p]),
    "teal_attn_decode_chain_f16": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_int,
                                              _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_double, _c.c_double, _c.c_double,
                                              _c.c_void_p, _c.c_size_t, _c.c_void_p]),
    "teal_sparse_o_proj_chain_f16": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_int, _c.c_int,
                                                _c.c_int64, _c.c_double, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_void_p, _c.c_size_t,
                                                _c.c_void_p]),
    "teal_sparse_gemv_tp_exchange_f16": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_int, _c.c_int, _c.c_int64,
                                                    _c.c_double, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_int, _c.c_int, _c.c_void_p,
                                                    _c.c_size_t, _c.c_void_p]),
    "teal_advance": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_int, _c.c_void_p]),
    "teal_survivors_f16": (_c.c_int, [_c.c_void_p, _c.c_int, _c.c_double, _c.c_void_p, _c.c_void_p, _c.c_void_p]),
    "teal_sparse_gemv_f16_simple": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_int, _c.c_int,
                                               _c.c_int64, _c.c_double, _c.c_void_p]),
    "teal_set_config": (_c.c_int, [_c.c_int]),
    "teal_set_tuning": (_c.c_int, [_c.c_int, _c.c_int]),
    "teal_set_pdl": (_c.c_int, [_c.c_int]),
    "teal_set_bytes_probe": (_c.c_int, [_c.c_void_p]),
    "teal_set_debug_timestamps": (_c.c_int, [_c.c_void_p]),
    "teal_set_debug_timeline": (_c.c_int, [_c.c_void_p, _c.c_int]),
}



class LayerDesc(_c.Structure):
    """``teal_layer_desc`` of include/teal_b200.h."""
    _fields_ = [("wqkv", _c.c_void_p), ("wo", _c.c_void_p), ("w1", _c.c_void_p), ("w3", _c.c_void_p), ("w2", _c.c_void_p),
                ("attn_norm", _c.c_void_p), ("ffn_norm", _c.c_void_p), ("k_cache", _c.c_void_p), ("v_cache", _c.c_void_p),
                ("ld_qkv", _c.c_int64), ("ld_o", _c.c_int64), ("ld_1", _c.c_int64), ("ld_3", _c.c_int64),
                ("ld_2", _c.c_int64), ("thr", _c.c_double * 7)]


class StepDesc(_c.Structure):
    """``teal_step_desc`` of include/teal_b200.h."""
    _fields_ = [("dim", _c.c_int32), ("n_head", _c.c_int32), ("n_kv_head", _c.c_int32), ("head_dim", _c.c_int32),
                ("inter", _c.c_int32), ("vocab", _c.c_int32), ("n_layer", _c.c_int32), ("max_seq", _c.c_int32),
                ("eps", _c.c_double),
                ("tok_emb", _c.c_void_p), ("token", _c.c_void_p), ("ssq_init", _c.c_void_p), ("lm_head", _c.c_void_p),
                ("final_norm", _c.c_void_p), ("ld_head", _c.c_int64), ("freqs_cis", _c.c_void_p), ("pos", _c.c_void_p),
                ("out_tokens", _c.c_void_p), ("argmax_packed", _c.c_void_p), ("x", _c.c_void_p), ("qkv", _c.c_void_p),
                ("att", _c.c_void_p), ("h", _c.c_void_p), ("logits", _c.c_void_p), ("scratch", _c.c_void_p),
                ("layers", _c.POINTER(LayerDesc)),
                ("out_cap", _c.c_int32), ("rows_small", _c.c_int32), ("rows_large", _c.c_int32),
                ("group_qkv", _c.c_int32), ("group_gate_up", _c.c_int32), ("head_rows", _c.c_int32),
                ("wide_melems", _c.c_int32), ("ctas_per_sm", _c.c_int32), ("poll_ns", _c.c_int32), ("opts", _c.c_int32), ("debug_times", _c.c_void_p), ("warps", _c.c_int32),
                ("stagger_ns", _c.c_int32), ("rows_kind", _c.c_int32 * 4), ("tiles_per_item", _c.c_int32), ("reserved2", _c.c_int32)]


class ItemInfo(_c.Structure):
    """``teal_item_info`` of include/teal_b200.h."""
    _fields_ = [(n, _c.c_int32) for n in ("phase", "kind", "layer", "item", "ctile", "kchunk", "k_begin", "k_end",
                                           "pub_index", "pub_flag", "pub_cell", "arrivals", "dep_kind", "dep_cell",
                                           "dep_per_epoch", "n_deps")] + [("deps", _c.c_int32 * 32), ("n_pubs", _c.c_int32),
                                                                            ("pub_indices", _c.c_int32 * 4),
                                                                            ("pub_flags", _c.c_int32 * 4),
                                                                            ("ctiles", _c.c_int32 * 4)]


_SIGS.update({
    "teal_step_program_sizes": (_c.c_int, [_c.POINTER(StepDesc), _c.POINTER(_c.c_size_t), _c.POINTER(_c.c_size_t)]),
    "teal_step_program_plan": (_c.c_int, [_c.POINTER(StepDesc), _c.c_void_p, _c.c_size_t]),
    "teal_step_program_run": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p]),
    "teal_step_program_item": (_c.c_int, [_c.c_void_p, _c.c_int, _c.POINTER(ItemInfo)]),
    "teal_step_program_items": (_c.c_int, [_c.c_void_p]),
})

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


def load():
    """Loads (building in-tree if the sources are newer) and returns the ctypes library."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        from . import build as _build
        _build.build()
    try:
        lib = ctypes.CDLL(LIB_PATH)
    except OSError as e:  # fail loudly: there is no fallback path
        raise TealLibraryError(f"cannot load {LIB_PATH}: {e}") from e
    _bind(lib)
    if lib.teal_abi_version() != 2:
        raise TealLibraryError("libteal_b200 ABI version mismatch")
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().teal_last_error().decode()
        if rc in (-1, -2):
            raise AssertionError(f"{what}: {msg}")  # reference raises AssertionError on shape/layout
        raise TealLibraryError(f"{what} failed ({rc}): {msg}")
