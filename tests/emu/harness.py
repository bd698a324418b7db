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
        _lib.teal_emu_set_max_workers.argtypes = [ctypes.c_int]
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
        scratch = 64 * 32 * 4                  # the trailing sum-of-squares scratch may hold values
        assert not self.buf[: self.nbytes - scratch].any(), "workspace not left zeroed"


def set_config(cfg):
    check(lib().teal_set_config(int(cfg)), "teal_set_config")


def set_variant(variant=-1, lanes_per_warp=0):
    """teal_set_config: force kernel variant / input lanes per warp (-1 / 0 = the by-size default)."""
    set_config(0 if variant < 0 else (variant + 1) + 16 * (lanes_per_warp // 32))


class EmuDecoder:
    """The buffers and call sequence of teal_b200.model.SparseDecoder (`_enqueue_step`) on numpy arrays against the
    emulation library: one decode step = the same C-ABI calls in the same order."""
    PROJS = ("q", "k", "v", "o", "gate", "up", "down")

    def __init__(self, cfg: dict, sd: dict, thresholds, max_seq: int = 64):
        import torch
        from oracle import teal_oracle as O
        self.cfg, self.L = cfg, lib()
        self.dim, self.n_head, self.n_kv = cfg["dim"], cfg["n_head"], cfg["n_local_heads"]
        self.hd = self.dim // self.n_head
        self.kv = self.n_kv * self.hd
        self.nqkv = self.dim + 2 * self.kv
        self.inter, self.vocab, self.eps, self.max_seq = cfg["intermediate_size"], cfg["vocab_size"], cfg["norm_eps"], max_seq
        self.thresholds = thresholds
        f16 = np.float16
        npy = lambda t: t.detach().to(torch.float16).numpy()
        self.tok_emb = aligned(sd["tok_embeddings.weight"].shape, f16, npy(sd["tok_embeddings.weight"]))
        self.lm_head = aligned(sd["output.weight"].shape, f16, npy(sd["output.weight"]))
        self.final_norm = aligned((self.dim,), f16, npy(sd["norm.weight"]))
        self.layers = []
        for i in range(cfg["n_layer"]):
            p = f"layers.{i}."
            self.layers.append(dict(
                wqkv=column_major(npy(sd[p + "attention.wqkv.weight"])), wo=column_major(npy(sd[p + "attention.wo.weight"])),
                w1=column_major(npy(sd[p + "feed_forward.w1.weight"])), w3=column_major(npy(sd[p + "feed_forward.w3.weight"])),
                w2=column_major(npy(sd[p + "feed_forward.w2.weight"])),
                attn_norm=aligned((self.dim,), f16, npy(sd[p + "attention_norm.weight"])),
                ffn_norm=aligned((self.dim,), f16, npy(sd[p + "ffn_norm.weight"])),
                k_cache=aligned((self.n_kv, max_seq, self.hd), f16, 0), v_cache=aligned((self.n_kv, max_seq, self.hd), f16, 0)))
        self.freqs = aligned((max_seq, self.hd // 2, 2), f16, npy(O.precompute_freqs_cis(max_seq, self.hd, cfg["rope_base"])))
        self.token, self.pos = aligned((1,), np.int32, 0), aligned((1,), np.int32, 0)
        self.out_tokens, self.argmax = aligned((256,), np.int32, 0), aligned((1,), np.uint64, 0)
        self.x, self.qkv, self.att = aligned((self.dim,), f16, 0), aligned((self.nqkv,), f16, 0), aligned((self.dim,), f16, 0)
        self.h, self.logits = aligned((self.inter,), f16, 0), aligned((self.vocab,), f16, 0)
        self.ssq = aligned((64 * 32,), np.float32, 0)   # TEAL_SSQ_FLOATS
        self.ws = Workspace(1 << 17)   # the product's size (teal_b200/ops.py)
        # deferred epilogue (SparseDecoder.deferred_epilogue): fp32 sums [q|k|v][gate|up] rounded by their consumers
        self.deferred = True
        self.defer_acc = aligned((self.nqkv + 2 * self.inter,), np.float32, 0)
        self.sampling = None   # dict(temperature, top_k, seed) as in SparseDecoder.set_sampling

    def reset(self, token):
        self.token[0], self.pos[0], self.argmax[0] = token, 0, 0
        self.defer_acc[...] = 0
        for W in self.layers:
            W["k_cache"][...] = 0
            W["v_cache"][...] = 0

    def step(self, sparse=True):
        """SparseDecoder._enqueue_step (world == 1): embed | per layer norm+qkv, attention, o+residual, norm+gate/up,
        down+residual | lm_head + arg-max | advance."""
        L, ws = self.L, self.ws
        wsp, wsn, ssq = ptr(ws.buf), ws.nbytes, ptr(self.ssq)
        check(L.teal_embed_f16(ptr(self.tok_emb), ptr(self.token), ptr(self.x), ssq, self.dim, None), "embed")
        for li, W in enumerate(self.layers):
            t = self.thresholds[li] if (sparse and self.thresholds[li] is not None) else {p: -1.0 for p in self.PROJS}
            if self.deferred:
                qkv_acc, gu_acc = ptr(self.defer_acc[: self.nqkv]), ptr(self.defer_acc[self.nqkv:])
                assert not self.defer_acc[: self.nqkv].any(), "q|k|v sums not cleared before their producer"
                check(L.teal_norm_sparse_qkv_gemv_acc_f16(ptr(self.x), ptr(W["attn_norm"]), self.eps, ssq, ptr(W["wqkv"]), qkv_acc,
                                                          self.dim, self.nqkv, self.nqkv, self.kv, t["q"], t["k"], t["v"], None), "norm_qkv")
                check(L.teal_attn_decode_acc_f16(qkv_acc, ptr(self.freqs), ptr(W["k_cache"]), ptr(W["v_cache"]), ptr(self.pos), ptr(self.att),
                                                 self.n_head, self.n_kv, self.hd, self.max_seq, None), "attn")
                check(L.teal_sparse_gemv_residual_ex_f16(ptr(self.att), None, ptr(W["wo"]), ptr(self.x), ptr(self.x), ssq, self.dim, self.dim,
                                                         self.dim, t["o"], ptr(self.defer_acc), self.defer_acc.size, wsp, wsn, None), "o_proj")
                assert not self.defer_acc.any(), "the o projection clears every deferred sum"
                check(L.teal_norm_sparse_gate_up_acc_f16(ptr(self.x), ptr(W["ffn_norm"]), self.eps, ssq, ptr(W["w1"]), ptr(W["w3"]), gu_acc,
                                                         self.dim, self.inter, self.inter, self.inter, t["gate"], t["up"], None), "gate_up")
                check(L.teal_sparse_gemv_residual_ex_f16(None, gu_acc, ptr(W["w2"]), ptr(self.x), ptr(self.x), ssq, self.inter, self.dim,
                                                         self.dim, t["down"], None, 0, wsp, wsn, None), "down_proj")
                continue
            check(L.teal_norm_sparse_qkv_gemv_f16(ptr(self.x), ptr(W["attn_norm"]), self.eps, ssq, ptr(W["wqkv"]), ptr(self.qkv), None,
                                                  self.dim, self.nqkv, self.nqkv, self.kv, t["q"], t["k"], t["v"], wsp, wsn, None), "norm_qkv")
            check(L.teal_attn_decode_f16(ptr(self.qkv), ptr(self.freqs), ptr(W["k_cache"]), ptr(W["v_cache"]), ptr(self.pos), ptr(self.att),
                                         self.n_head, self.n_kv, self.hd, self.max_seq, None), "attn")
            check(L.teal_sparse_gemv_residual_f16(ptr(self.att), ptr(W["wo"]), ptr(self.x), ptr(self.x), ssq, self.dim, self.dim, self.dim,
                                                  t["o"], wsp, wsn, None), "o_proj")
            check(L.teal_norm_sparse_gate_up_f16(ptr(self.x), ptr(W["ffn_norm"]), self.eps, ssq, ptr(W["w1"]), ptr(W["w3"]), ptr(self.h), None,
                                                 self.dim, self.inter, self.inter, self.inter, t["gate"], t["up"], wsp, wsn, None), "gate_up")
            check(L.teal_sparse_gemv_residual_f16(ptr(self.h), ptr(W["w2"]), ptr(self.x), ptr(self.x), ssq, self.inter, self.dim, self.dim,
                                                  t["down"], wsp, wsn, None), "down_proj")
        check(L.teal_dense_gemv_f16(ptr(self.x), ptr(self.final_norm), self.eps, ptr(self.lm_head), ptr(self.logits), ptr(self.argmax),
                                    self.dim, self.vocab, self.dim, None), "lm_head")
        if self.sampling is not None:
            sm = self.sampling
            check(L.teal_sample_topk_f16(ptr(self.logits), self.vocab, sm["temperature"], sm["top_k"], None, sm["seed"], ptr(self.pos),
                                         ptr(self.argmax), None, None), "sample")
        check(L.teal_advance(ptr(self.argmax), ptr(self.token), ptr(self.pos), ptr(self.out_tokens), self.out_tokens.size, None), "advance")
        ws.assert_clean()
