"""gpt-fast-shaped nn.Module model whose sparse call sites are the reference's:
``layer.attention.gemv1(x, wqkv.weight, thresh_q, thresh_k, thresh_v, sparsity_bin, kv_size)``,
``layer.attention.gemv2(y, wo.weight, thresh_o, sparsity_bin)``, ``layer.feed_forward.gemv1 / gemv2`` -- i.e. what
gpt-fast/model.py:163-190,258-259 does after ``monkeypatch_layer`` (generate.py:266-323), but with
``torch.ops.teal.*`` backed by libteal_b200.  This module exists to show (and test) that the two ops are a drop-in
for the reference's call sites; the fast path of this repo is ``teal_b200.model.SparseDecoder``.
"""
from __future__ import annotations

import types
from typing import Dict, List, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from .model import ModelArgs, find_multiple, precompute_freqs_cis
from .ops import SparseGEMV, SparseQKVGEMV, to_column_major


class RMSNorm(nn.Module):
    def __init__(self, dim: int, eps: float = 1e-5):
        super().__init__()
        self.eps = eps
        self.weight = nn.Parameter(torch.ones(dim))

    def forward(self, x):
        xf = x.float()
        return (xf * torch.rsqrt(torch.mean(xf * xf, dim=-1, keepdim=True) + self.eps)).type_as(x) * self.weight


def apply_rotary_emb(x, freqs_cis):
    xs = x.float().reshape(*x.shape[:-1], -1, 2)
    fc = freqs_cis.view(1, xs.size(1), 1, xs.size(3), 2)
    out = torch.stack([xs[..., 0] * fc[..., 0] - xs[..., 1] * fc[..., 1],
                       xs[..., 1] * fc[..., 0] + xs[..., 0] * fc[..., 1]], -1)
    return out.flatten(3).type_as(x)


class KVCache(nn.Module):
    def __init__(self, max_seq, n_heads, head_dim, dtype):
        super().__init__()
        self.register_buffer("k_cache", torch.zeros(1, n_heads, max_seq, head_dim, dtype=dtype))
        self.register_buffer("v_cache", torch.zeros(1, n_heads, max_seq, head_dim, dtype=dtype))

    def update(self, input_pos, k, v):
        self.k_cache[:, :, input_pos] = k
        self.v_cache[:, :, input_pos] = v
        return self.k_cache, self.v_cache


def _sparse_attn_forward(self, x, freqs_cis, mask, input_pos):
    bsz, seqlen, _ = x.shape
    kv_size = self.n_local_heads * self.head_dim
    q, k, v = self.gemv1(x, self.wqkv.weight, self.thresh_q, self.thresh_k, self.thresh_v, self.sparsity_bin,
                         kv_size).split([self.dim, kv_size, kv_size], dim=-1)
    return self._attend(q, k, v, bsz, seqlen, freqs_cis, mask, input_pos, sparse=True)


def _sparse_ffn_forward(self, x):
    return self.gemv2(F.silu(self.gemv1(x, self.w1.weight, self.thresh_gate, self.sparsity_bin)) *
                      self.gemv1(x, self.w3.weight, self.thresh_up, self.sparsity_bin),
                      self.w2.weight, self.thresh_down, self.sparsity_bin)


class Attention(nn.Module):
    def __init__(self, a: ModelArgs):
        super().__init__()
        self.wqkv = nn.Linear(a.dim, (a.n_head + 2 * a.n_local_heads) * a.head_dim, bias=False)
        self.wo = nn.Linear(a.dim, a.dim, bias=False)
        self.kv_cache: Optional[KVCache] = None
        self.n_head, self.head_dim, self.n_local_heads, self.dim = a.n_head, a.head_dim, a.n_local_heads, a.dim

    def apply_monkeypatch(self):
        self.old_forward = self.forward
        self.forward = types.MethodType(_sparse_attn_forward, self)

    def _attend(self, q, k, v, bsz, seqlen, freqs_cis, mask, input_pos, sparse):
        q = apply_rotary_emb(q.view(bsz, seqlen, self.n_head, self.head_dim), freqs_cis)
        k = apply_rotary_emb(k.view(bsz, seqlen, self.n_local_heads, self.head_dim), freqs_cis)
        v = v.view(bsz, seqlen, self.n_local_heads, self.head_dim)
        q, k, v = (t.transpose(1, 2) for t in (q, k, v))
        if self.kv_cache is not None:
            k, v = self.kv_cache.update(input_pos, k, v)
        rep = self.n_head // self.n_local_heads
        y = F.scaled_dot_product_attention(q, k.repeat_interleave(rep, dim=1), v.repeat_interleave(rep, dim=1),
                                           attn_mask=mask, dropout_p=0.0)
        y = y.transpose(1, 2).contiguous().view(bsz, seqlen, self.dim)
        return self.gemv2(y, self.wo.weight, self.thresh_o, self.sparsity_bin) if sparse else self.wo(y)

    def forward(self, x, freqs_cis, mask, input_pos):
        bsz, seqlen, _ = x.shape
        kv_size = self.n_local_heads * self.head_dim
        q, k, v = self.wqkv(x).split([self.dim, kv_size, kv_size], dim=-1)
        return self._attend(q, k, v, bsz, seqlen, freqs_cis, mask, input_pos, sparse=False)


class FeedForward(nn.Module):
    def __init__(self, a: ModelArgs):
        super().__init__()
        self.w1 = nn.Linear(a.dim, a.intermediate_size, bias=False)
        self.w3 = nn.Linear(a.dim, a.intermediate_size, bias=False)
        self.w2 = nn.Linear(a.intermediate_size, a.dim, bias=False)

    def apply_monkeypatch(self):
        self.old_forward = self.forward
        self.forward = types.MethodType(_sparse_ffn_forward, self)

    def forward(self, x):
        return self.w2(F.silu(self.w1(x)) * self.w3(x))


class TransformerBlock(nn.Module):
    def __init__(self, a: ModelArgs):
        super().__init__()
        self.attention, self.feed_forward = Attention(a), FeedForward(a)
        self.ffn_norm, self.attention_norm = RMSNorm(a.dim, a.norm_eps), RMSNorm(a.dim, a.norm_eps)

    def forward(self, x, input_pos, freqs_cis, mask):
        h = x + self.attention(self.attention_norm(x), freqs_cis, mask, input_pos)
        return h + self.feed_forward(self.ffn_norm(h))


class Transformer(nn.Module):
    def __init__(self, config: ModelArgs):
        super().__init__()
        self.config = config
        self.tok_embeddings = nn.Embedding(config.vocab_size, config.dim)
        self.layers = nn.ModuleList(TransformerBlock(config) for _ in range(config.n_layer))
        self.norm = RMSNorm(config.dim, config.norm_eps)
        self.output = nn.Linear(config.dim, config.vocab_size, bias=False)
        self.freqs_cis = None
        self.causal_mask = None

    def setup_caches(self, max_batch_size: int, max_seq_length: int):
        max_seq_length = find_multiple(max_seq_length, 8)
        dtype, dev = self.output.weight.dtype, self.output.weight.device
        for b in self.layers:
            b.attention.kv_cache = KVCache(max_seq_length, self.config.n_local_heads, self.config.head_dim, dtype).to(dev)
        self.freqs_cis = precompute_freqs_cis(self.config.block_size, self.config.head_dim, self.config.rope_base, dtype).to(dev)
        self.causal_mask = torch.tril(torch.ones(max_seq_length, max_seq_length, dtype=torch.bool, device=dev))

    def forward(self, idx, input_pos):
        mask = self.causal_mask[None, None, input_pos]
        freqs_cis = self.freqs_cis[input_pos]
        x = self.tok_embeddings(idx)
        for layer in self.layers:
            x = layer(x, input_pos, freqs_cis, mask)
        return self.output(self.norm(x))


def install_sparsity(model: Transformer, thresholds: List[Dict[str, float]], device: str = "cuda") -> Transformer:
    """What ``monkeypatch_layer`` does (generate.py:289-321) given per-layer {q,k,v,o,gate,up,down} thresholds
    (from ``teal_b200.distribution.layer_thresholds`` for real histograms): install the ops, re-lay the five weights
    out column-major and swap the forwards."""
    for layer, t in zip(model.layers, thresholds):
        ff, at = layer.feed_forward, layer.attention
        ff.gemv1_kernel = SparseGEMV.initialize("sparse_gemv", device)
        ff.gemv1 = ff.gemv1_kernel.operator(True)
        ff.gemv2_kernel = SparseGEMV.initialize("sparse_gemv", device)
        ff.gemv2 = ff.gemv2_kernel.operator(True)
        ff.thresh_up, ff.thresh_gate, ff.thresh_down, ff.sparsity_bin = t["up"], t["gate"], t["down"], 0
        at.gemv1_kernel = SparseQKVGEMV.initialize("sparse_qkv_gemv", device)
        at.gemv1 = at.gemv1_kernel.operator(True)
        at.gemv2_kernel = SparseGEMV.initialize("sparse_gemv", device)
        at.gemv2 = at.gemv2_kernel.operator(True)
        at.thresh_q, at.thresh_k, at.thresh_v, at.thresh_o, at.sparsity_bin = t["q"], t["k"], t["v"], t["o"], 0
        for lin in (ff.w1, ff.w3, ff.w2, at.wqkv, at.wo):
            lin.weight.data = to_column_major(lin.weight.data)
        ff.apply_monkeypatch()
        at.apply_monkeypatch()
    return model.eval()
