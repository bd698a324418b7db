"""Operator surface of the sparse decode path: ``torch.ops.teal.sparse_gemv`` and
``torch.ops.teal.sparse_qkv_gemv`` backed by libteal_b200 (hand-written sm_100a CUDA).

Mirrors the reference's op layer so that its call sites keep working unchanged:

* ``SparseGEMV`` / ``SparseQKVGEMV`` with ``initialize(name, target)``, ``.operator(compiled)``,
  ``.meta``, ``.forward``, ``.register()`` -- kernels/sparse_gemv.py:249-298 and
  kernels/compile_wrapper.py:124-208 (``BaseKernel``) in the reference.
* schemas are exactly the strings compile_wrapper.py:151-162 infers (tests/golden/op_schemas.txt).
* S == 1 -> fused threshold + gather-GEMV kernel; S > 1 (prefill) -> dense ``x @ W.T`` with no
  sparsification (kernels/sparse_gemv.py:271,298).
* weights must be the "column major" view (stride(1) > 1, kernels/sparse_gemv.py:106,210).

There is no Triton, no autotune warm-up and no CPU fallback: a CUDA tensor is required and the
shared library must load, otherwise the op raises.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass
from operator import attrgetter
from typing import Callable, Dict, Tuple

import torch

from . import _lib

NAMESPACE = "teal"  # kernels/compile_wrapper.py:19

_MAX_OUT_COLS = 1 << 17
_workspaces: Dict[Tuple[int, int], torch.Tensor] = {}


_alloc_streams: Dict[int, "torch.cuda.Stream"] = {}
_ws_pools: Dict[int, list] = {}   # device -> free workspace slots (allocated 8 at a time)
_WS_SLOTS = 8


def workspace(device: torch.device, stream_ptr: int) -> torch.Tensor:
    """Per (device, stream) accumulator/counter workspace; zeroed once, kernels keep it zeroed.
    Slots come from a small per-device pool that is allocated 8 at a time on a private side stream: a stream that shows up
    for the first time INSIDE a CUDA-graph capture (torch.compile(mode="reduce-overhead") records on its own stream after
    an eager warm-up) takes a ready slot without allocating, and allocations made while another stream is current are not
    redirected into a graph memory pool (cudagraph trees insist that every live tensor of their pool is a graph output)."""
    dev = device.index if device.index is not None else torch.cuda.current_device()
    key = (dev, stream_ptr)
    ws = _workspaces.get(key)
    if ws is None:
        pool = _ws_pools.setdefault(dev, [])
        if not pool:
            nbytes = (_lib.load().teal_workspace_bytes(_MAX_OUT_COLS) + 255) // 256 * 256
            side = _alloc_streams.get(dev)
            if side is None:
                side = _alloc_streams[dev] = torch.cuda.Stream(device=dev)
            with torch.cuda.stream(side):
                block = torch.zeros(_WS_SLOTS * nbytes, dtype=torch.uint8, device=device)
            side.synchronize()
            pool.extend(block[i * nbytes:(i + 1) * nbytes] for i in range(_WS_SLOTS))
        ws = _workspaces[key] = pool.pop()
    return ws


def _stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def _check_x(x: torch.Tensor, Z: int) -> torch.Tensor:
    if not x.is_cuda:
        raise _lib.TealLibraryError("teal_b200 ops need CUDA tensors (there is no CPU fallback)")
    assert x.dim() == 3 and x.shape[2] == Z, "hidden size mismatch"       # sparse_gemv.py:103,207
    assert x.shape[0] == 1, "batch/beam must be 1 (sparse_gemv.py:73)"
    return x.contiguous()                                                   # sparse_gemv.py:104


def _check_w(weight: torch.Tensor) -> int:
    assert weight.dim() == 2 and weight.stride(1) > 1 and weight.stride(0) == 1, \
        "weight should be column major"                                     # sparse_gemv.py:106,210
    assert weight.dtype == torch.float16, "weights must be fp16 (README.md:108)"
    return weight.stride(1)


def _finish(out: torch.Tensor, x_dtype: torch.dtype) -> torch.Tensor:
    if x_dtype is not torch.float16:                                        # sparse_gemv.py:138-140
        print("Warning: incuring dtype conversion overhead since input dtype is not torch.float16. "
              f"Detected dtype: {x_dtype}. ")
        return out.to(dtype=x_dtype)
    return out


def splitk_sparse_gemv(x: torch.Tensor, weight: torch.Tensor, threshold: float, sparsity_bin: int = 0) -> torch.Tensor:
    """Decode-time ``y = sparse(x) @ weight.T`` (name kept from kernels/sparse_gemv.py:87).
    x [1,1,Z] fp16, weight logical [N,Z] with stride (1, ldw).  ``sparsity_bin`` is accepted and
    ignored (it only keyed the reference's Triton autotune cache)."""
    N, Z = weight.shape
    x = _check_x(x, Z)
    ldw = _check_w(weight)
    xin = x if x.dtype is torch.float16 else x.to(torch.float16)
    out = torch.empty(x.shape[0], x.shape[1], N, device=x.device, dtype=torch.float16)
    sp = _stream_ptr()
    ws = workspace(x.device, sp)
    rc = _lib.load().teal_sparse_gemv_f16(xin.data_ptr(), weight.data_ptr(), out.data_ptr(), Z, N, ldw,
                                          float(threshold), ws.data_ptr(), ws.numel(), sp)
    _lib.check(rc, "teal_sparse_gemv_f16")
    return _finish(out, x.dtype)


def qkv_gemv(x: torch.Tensor, weight: torch.Tensor, threshold_q: float, threshold_k: float, threshold_v: float,
             sparsity_bin: int, kv_size: int) -> torch.Tensor:
    """Fused QKV decode GEMV with a threshold per output region (kernels/sparse_gemv.py:196-237)."""
    N, Z = weight.shape
    x = _check_x(x, Z)
    ldw = _check_w(weight)
    xin = x if x.dtype is torch.float16 else x.to(torch.float16)
    out = torch.empty(x.shape[0], x.shape[1], N, device=x.device, dtype=torch.float16)
    sp = _stream_ptr()
    ws = workspace(x.device, sp)
    rc = _lib.load().teal_sparse_qkv_gemv_f16(xin.data_ptr(), weight.data_ptr(), out.data_ptr(), Z, N, ldw,
                                              int(kv_size), float(threshold_q), float(threshold_k),
                                              float(threshold_v), ws.data_ptr(), ws.numel(), sp)
    _lib.check(rc, "teal_sparse_qkv_gemv_f16")
    return _finish(out, x.dtype)


def sparse_gate_up(x: torch.Tensor, w1: torch.Tensor, w3: torch.Tensor, threshold_gate: float,
                   threshold_up: float) -> torch.Tensor:
    """``silu(gemv1(x, w1, t_gate)) * gemv1(x, w3, t_up)`` in one kernel (gpt-fast/model.py:259)."""
    I, Z = w1.shape
    assert w3.shape == w1.shape
    x = _check_x(x, Z)
    assert x.shape[1] == 1 and x.dtype is torch.float16
    ld1, ld3 = _check_w(w1), _check_w(w3)
    out = torch.empty(1, 1, I, device=x.device, dtype=torch.float16)
    sp = _stream_ptr()
    ws = workspace(x.device, sp)
    rc = _lib.load().teal_sparse_gate_up_f16(x.data_ptr(), w1.data_ptr(), w3.data_ptr(), out.data_ptr(), Z, I,
                                             ld1, ld3, float(threshold_gate), float(threshold_up),
                                             ws.data_ptr(), ws.numel(), sp)
    _lib.check(rc, "teal_sparse_gate_up_f16")
    return out


def norm_qkv_gemv(x, norm_w, eps, weight, tq, tk, tv, kv_size, want_xn=False, ssq_in=None):
    """attention_norm + fused QKV GEMV in one launch (decode pipeline building block); ``ssq_in``: the producer's
    sum-of-squares partials of x (else computed by one extra tiny launch)."""
    N, Z = weight.shape
    x = _check_x(x, Z)
    ldw = _check_w(weight)
    out = torch.empty(1, 1, N, device=x.device, dtype=torch.float16)
    xn = torch.empty(1, 1, Z, device=x.device, dtype=torch.float16) if want_xn else None
    sp = _stream_ptr()
    ws = workspace(x.device, sp)
    rc = _lib.load().teal_norm_sparse_qkv_gemv_f16(x.data_ptr(), norm_w.data_ptr(), float(eps),
                                                   ssq_in.data_ptr() if ssq_in is not None else None, weight.data_ptr(),
                                                   out.data_ptr(), xn.data_ptr() if want_xn else None, Z, N, ldw,
                                                   int(kv_size), float(tq), float(tk), float(tv), ws.data_ptr(),
                                                   ws.numel(), sp)
    _lib.check(rc, "teal_norm_sparse_qkv_gemv_f16")
    return (out, xn) if want_xn else out


def norm_gate_up(x, norm_w, eps, w1, w3, t_gate, t_up, want_xn=False):
    """ffn_norm + gate/up GEMVs + SiLU*mul in one launch."""
    I, Z = w1.shape
    x = _check_x(x, Z)
    ld1, ld3 = _check_w(w1), _check_w(w3)
    out = torch.empty(1, 1, I, device=x.device, dtype=torch.float16)
    xn = torch.empty(1, 1, Z, device=x.device, dtype=torch.float16) if want_xn else None
    sp = _stream_ptr()
    ws = workspace(x.device, sp)
    rc = _lib.load().teal_norm_sparse_gate_up_f16(x.data_ptr(), norm_w.data_ptr(), float(eps), None, w1.data_ptr(),
                                                  w3.data_ptr(), out.data_ptr(), xn.data_ptr() if want_xn else None,
                                                  Z, I, ld1, ld3, float(t_gate), float(t_up), ws.data_ptr(),
                                                  ws.numel(), sp)
    _lib.check(rc, "teal_norm_sparse_gate_up_f16")
    return (out, xn) if want_xn else out


def gemv_residual(x, weight, resid, threshold, ssq_out=None):
    """resid + sparse_gemv(x, weight) with the add fused into the kernel's finaliser; ``ssq_out`` (fp32[SSQ_FLOATS], optional)
    receives the per-tile sums of squares of the result for the next fused RMSNorm."""
    N, Z = weight.shape
    x = _check_x(x, Z)
    ldw = _check_w(weight)
    out = torch.empty(1, 1, N, device=x.device, dtype=torch.float16)
    sp = _stream_ptr()
    ws = workspace(x.device, sp)
    rc = _lib.load().teal_sparse_gemv_residual_f16(x.data_ptr(), weight.data_ptr(), resid.data_ptr(), out.data_ptr(),
                                                   ssq_out.data_ptr() if ssq_out is not None else None, Z, N, ldw, float(threshold), ws.data_ptr(), ws.numel(), sp)
    _lib.check(rc, "teal_sparse_gemv_residual_f16")
    return out


def survivors(x: torch.Tensor, threshold: float) -> torch.Tensor:
    """Sorted surviving lane indices, computed by the device compaction code (parity hook)."""
    xf = x.reshape(-1).contiguous()
    assert xf.is_cuda and xf.dtype is torch.float16
    idx = torch.empty(xf.numel(), dtype=torch.int32, device=x.device)
    cnt = torch.zeros(1, dtype=torch.int32, device=x.device)
    rc = _lib.load().teal_survivors_f16(xf.data_ptr(), xf.numel(), float(threshold), idx.data_ptr(), cnt.data_ptr(),
                                        _stream_ptr())
    _lib.check(rc, "teal_survivors_f16")
    return idx[: int(cnt.item())]


def sparse_gemv_simple(x: torch.Tensor, weight: torch.Tensor, threshold: float) -> torch.Tensor:
    N, Z = weight.shape
    x = _check_x(x, Z)
    ldw = _check_w(weight)
    out = torch.empty(1, 1, N, device=x.device, dtype=torch.float16)
    rc = _lib.load().teal_sparse_gemv_f16_simple(x.data_ptr(), weight.data_ptr(), out.data_ptr(), Z, N, ldw,
                                                 float(threshold), _stream_ptr())
    _lib.check(rc, "teal_sparse_gemv_f16_simple")
    return out


def to_column_major(weight: torch.Tensor) -> torch.Tensor:
    """``W.T.contiguous().T``: the re-layout gpt-fast/generate.py:296-317 applies to every sparse
    projection (storage becomes [Z][N], N contiguous)."""
    return weight.T.contiguous().T


# ------------------------------------------------------------------------------------ op classes
@dataclass
class BaseKernel:
    """Registration helper with the reference's public surface (compile_wrapper.py:124-208)."""
    name: str
    target: str
    schema: str

    SCHEMA = ""

    @classmethod
    def initialize(cls, name: str, target: str, **kwargs) -> "BaseKernel":
        return cls(name, target, cls.schematize())

    @classmethod
    def schematize(cls) -> str:
        return cls.SCHEMA

    @property
    def is_registered(self) -> bool:
        return hasattr(getattr(torch.ops, NAMESPACE), self.name)

    def operator(self, compiled: bool = False) -> Callable:
        if compiled:
            self.register()
            return attrgetter(f"{NAMESPACE}.{self.name}")(torch.ops)
        return self.forward

    def meta(self, *args, **kwargs):
        raise NotImplementedError(f"abstract implementation `meta` in {self.__class__.__name__} required for registration")

    def forward(self, *args, **kwargs):
        raise NotImplementedError(f"concrete implementation `forward` in {self.__class__.__name__} required for registration")

    def register(self) -> None:
        if not self.is_registered:
            qualname = f"{NAMESPACE}::{self.name}"
            torch.library.define(qualname, self.schema)
            torch.library.register_fake(qualname, self.meta)
            torch.library.impl(qualname, self.target, func=self.forward)
        if torch.cuda.is_available():  # model-load time: make sure the workspace slot pool exists before any graph capture /
            workspace(torch.device("cuda", torch.cuda.current_device()), _stream_ptr())  # torch.compile warm-up can need a slot


class SparseGEMV(BaseKernel):
    """kernels/sparse_gemv.py:249-271."""
    SCHEMA = "(Tensor hidden_states, Tensor weights, float threshold, int sparsity_bin) -> Tensor"

    def meta(self, hidden_states, weights, threshold, sparsity_bin):
        return hidden_states.new_empty((hidden_states.size(0), hidden_states.size(1), weights.size(0)))

    def forward(self, hidden_states, weights, threshold, sparsity_bin):
        if hidden_states.shape[1] == 1:
            return splitk_sparse_gemv(hidden_states, weights, threshold, sparsity_bin)
        return torch.matmul(hidden_states, weights.T)  # prefill: dense, no sparsification


class SparseQKVGEMV(BaseKernel):
    """kernels/sparse_gemv.py:274-298."""
    SCHEMA = ("(Tensor x, Tensor weight, float threshold_q, float threshold_k, float threshold_v, "
              "int sparsity_bin, int kv_size) -> Tensor")

    def meta(self, x, weight, threshold_q, threshold_k, threshold_v, sparsity_bin, kv_size):
        return x.new_empty(x.shape[0], x.shape[1], weight.shape[0])

    def forward(self, x, weight, threshold_q, threshold_k, threshold_v, sparsity_bin, kv_size):
        if x.shape[1] == 1:
            return qkv_gemv(x, weight, threshold_q, threshold_k, threshold_v, sparsity_bin, kv_size)
        return torch.matmul(x, weight.T)


def set_config(config_id: int = 0, variant: int = -1, lanes_per_warp: int = 0) -> None:
    """Benchmark / parity-test knob of the calling thread (``teal_set_config``): force the kernel variant (0..4) and / or
    the input lanes one warp owns; ``set_config()`` restores the by-size default."""
    if variant >= 0:
        config_id = (variant + 1) + 16 * (lanes_per_warp // 32)
    _lib.check(_lib.load().teal_set_config(int(config_id)), "teal_set_config")


NUM_VARIANTS = 5  # kernel variants of csrc/gemv.cu (tile width x warps per CTA)
SSQ_PARTS, SSQ_REPLICAS = 64, 32  # TEAL_SSQ_PARTS / TEAL_SSQ_REPLICAS of include/teal_b200.h
SSQ_FLOATS = SSQ_PARTS * SSQ_REPLICAS  # floats in a sum-of-squares buffer (replica r at [r * SSQ_PARTS])

