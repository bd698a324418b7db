"""TEST INFRASTRUCTURE (only tests/ and oracle/gen_golden_greedy.py import this): a tiny decoder layer with the calling
convention and the attribute layout the reference's greedy optimiser drives (teal/greedyopt.py:53-84: `layer.mlp.sparse_fns[...]`,
`layer.self_attn.sparse_fns[...]`, `layer(hidden, attention_mask, position_ids, past_key_value, output_attentions,
use_cache, cache_position)[0]`), with the thresholded projections placed where teal/mlp.py:37-57 and
teal/self_attn.py:46-156 place them.  The SparsifyFn class is a parameter: the golden generator passes the REFERENCE's
(utils/utils.py:9-55) with the reference's Distribution, the parity test passes teal_b200.hf.SparsifyFn with
teal_b200.distribution.Distribution -- same weights, same inputs, same torch ops on the CPU, so the two runs of the
greedy loop must produce the same CSV to the last digit."""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

DIM, HEADS, INTER, SEQ, BATCH = 64, 4, 160, 24, 3
PROJS = ["q", "k", "v", "o", "gate", "up", "down"]


class _Norm(nn.Module):
    def __init__(self, g):
        super().__init__()
        self.weight = nn.Parameter(1 + 0.1 * torch.randn(DIM, generator=g))

    def forward(self, x):
        return x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + 1e-5) * self.weight


def _linear(n, z, g):
    lin = nn.Linear(z, n, bias=False)
    lin.weight.data = torch.randn(n, z, generator=g) / math.sqrt(z)
    return lin


class _Attn(nn.Module):
    def __init__(self, g, fns):
        super().__init__()
        self.q_proj, self.k_proj, self.v_proj, self.o_proj = (_linear(DIM, DIM, g) for _ in range(4))
        self.sparse_fns = nn.ModuleDict({p: fns[p] for p in ("q", "k", "v", "o")})

    def forward(self, x):
        b, s, _ = x.shape
        hd = DIM // HEADS
        q = self.q_proj(self.sparse_fns["q"](x)).view(b, s, HEADS, hd).transpose(1, 2)
        k = self.k_proj(self.sparse_fns["k"](x)).view(b, s, HEADS, hd).transpose(1, 2)
        v = self.v_proj(self.sparse_fns["v"](x)).view(b, s, HEADS, hd).transpose(1, 2)
        att = F.scaled_dot_product_attention(q, k, v, is_causal=True).transpose(1, 2).reshape(b, s, DIM)
        return self.o_proj(self.sparse_fns["o"](att))


class _Mlp(nn.Module):
    def __init__(self, g, fns):
        super().__init__()
        self.gate_proj, self.up_proj, self.down_proj = _linear(INTER, DIM, g), _linear(INTER, DIM, g), _linear(DIM, INTER, g)
        self.sparse_fns = nn.ModuleDict({p: fns[p] for p in ("gate", "up", "down")})

    def forward(self, x):
        h = F.silu(self.gate_proj(self.sparse_fns["gate"](x))) * self.up_proj(self.sparse_fns["up"](x))
        return self.down_proj(self.sparse_fns["down"](h))


class Layer(nn.Module):
    def __init__(self, fns, seed=0):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.input_layernorm, self.post_attention_layernorm = _Norm(g), _Norm(g)
        self.self_attn, self.mlp = _Attn(g, fns), _Mlp(g, fns)

    def forward(self, hidden_states, attention_mask=None, position_ids=None, past_key_value=None, output_attentions=False,
                use_cache=False, cache_position=None):
        h = hidden_states + self.self_attn(self.input_layernorm(hidden_states))
        return (h + self.mlp(self.post_attention_layernorm(h)),)


def inputs(seed=1):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(BATCH, SEQ, DIM, generator=g)


def site_activations(layer: Layer, x: torch.Tensor):
    """The tensors entering the projections (h1 / h2 of the attention and the MLP block): what the histograms are built of."""
    with torch.no_grad():
        a1 = layer.input_layernorm(x)
        b, s, _ = x.shape
        hd = DIM // HEADS
        q, k, v = (p(a1).view(b, s, HEADS, hd).transpose(1, 2) for p in (layer.self_attn.q_proj, layer.self_attn.k_proj, layer.self_attn.v_proj))
        a2 = F.scaled_dot_product_attention(q, k, v, is_causal=True).transpose(1, 2).reshape(b, s, DIM)
        h = x + layer.self_attn.o_proj(a2)
        m1 = layer.post_attention_layernorm(h)
        m2 = F.silu(layer.mlp.gate_proj(m1)) * layer.mlp.up_proj(m1)
    return {"self_attn": {"h1": a1, "h2": a2}, "mlp": {"h1": m1, "h2": m2}}
