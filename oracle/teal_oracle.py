"""CPU oracle for the TEAL activation-sparse decode path.

TEST INFRASTRUCTURE ONLY.  Nothing in ``teal_b200/`` may import this module; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs use it, and only as the checker (or as the timed CPU baseline), never as the product path.

Every function is a restatement (plain PyTorch-CPU / numpy) of a piece of the reference,
cited as ``file:line`` relative to ``/root/reference``.  Parity pinning: the reference ships no
tests or golden vectors, so this oracle is pinned against outputs of the
reference's own Python code imported in the build container (``oracle/gen_golden.py`` ->
``tests/golden/*.npz``; checked by ``tests/test_oracle_golden.py``).

Mask semantics pinned here: the reference evaluates
``x.abs().gt(threshold)`` on an fp16 tensor with a Python-float threshold
(utils/utils.py:51-52); torch demotes the scalar to fp16 (round-to-nearest-even) before the
compare, so the effective threshold is ``fp16(threshold)`` and the compare is strict ``>``.
NaN lanes fail the compare but ``False * nan`` is still nan, so a NaN lane poisons every output
(the Triton kernel behaves the same: kernels/sparse_gemv.py:73-78 multiplies the loaded-as-zero
row by x0 = nan).
"""
from __future__ import annotations

import csv
import os
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

PROJS = ["q", "k", "v", "o", "gate", "up", "down"]  # utils/utils.py:246


# --------------------------------------------------------------------------- thresholds
def round_threshold_fp16(threshold: float) -> float:
    """Effective threshold of ``x.abs().gt(threshold)`` for fp16 ``x`` (utils/utils.py:51-52)."""
    with np.errstate(over="ignore"):
        return float(np.float16(threshold))


def sparsify_apply(x: torch.Tensor, threshold: float) -> torch.Tensor:
    """``SparsifyFn.apply``: utils/utils.py:51-52 (verbatim semantics, any dtype)."""
    return x.abs().gt(threshold) * x


def survivor_mask(x: torch.Tensor, threshold: float) -> torch.Tensor:
    """Boolean keep-mask of utils/utils.py:52 (``x.abs().gt(threshold)``)."""
    return x.abs().gt(threshold)


def survivor_indices(x: torch.Tensor, threshold: float) -> np.ndarray:
    """Sorted indices m with |x[m]| > threshold (flattened last dim)."""
    return np.flatnonzero(survivor_mask(x.reshape(-1), threshold).numpy()).astype(np.int32)


def sparsify_forward(x: torch.Tensor, threshold: float, apply_prefill: bool = True) -> torch.Tensor:
    """``SparsifyFn.forward``: utils/utils.py:32-49 (prefill half-sequence rule + decode)."""
    if x.size(1) > 1 and apply_prefill:
        half = x.size(1) // 2
        last = sparsify_apply(x[:, -half:, :], threshold)
        return torch.cat((x[:, :-half, :], last), dim=1)
    if x.size(1) > 1 and not apply_prefill:
        return x
    assert x.size(1) == 1, "supposedly x is decode only"
    return sparsify_apply(x, threshold)


# --------------------------------------------------------------------------- projections
def sparse_linear(x: torch.Tensor, weight: torch.Tensor, threshold: float) -> torch.Tensor:
    """Reference threshold-then-matmul for one projection.

    HF path: ``proj(sparse_fns[p](x))`` teal/mlp.py:50-56, teal/self_attn.py:77-84,150-151.
    gpt-fast path computes the same thing with the Triton kernel
    (kernels/sparse_gemv.py:60-83).  ``weight`` is the logical ``[N, Z]`` nn.Linear weight.
    """
    return F.linear(sparsify_apply(x, threshold), weight)


def sparse_linear_f32(x: torch.Tensor, weight: torch.Tensor, threshold: float) -> torch.Tensor:
    """Same op with the mask taken in x's dtype but the matmul accumulated in fp32 and NOT
    rounded -- the 'truth' used to state tolerances."""
    xm = sparsify_apply(x, threshold)
    return F.linear(xm.float(), weight.float())


def qkv_sparse_linear(x: torch.Tensor, wqkv: torch.Tensor, tq: float, tk: float, tv: float,
                      kv_size: int) -> torch.Tensor:
    """Fused QKV with one threshold per output region.

    Regions follow kernels/sparse_gemv.py:212-213 (``N_q = N - 2*kv_size``) and the intent of
    :167-168,181 (q / k / v columns use tq / tk / tv).  The reference selects the threshold per
    BLOCK_N tile (wrong when a tile straddles regions); the oracle
    -- like teal/self_attn.py:77-84 -- selects it per column region exactly.
    """
    n = wqkv.shape[0]
    n_q = n - 2 * kv_size
    q = sparse_linear(x, wqkv[:n_q], tq)
    k = sparse_linear(x, wqkv[n_q:n_q + kv_size], tk)
    v = sparse_linear(x, wqkv[n_q + kv_size:], tv)
    return torch.cat([q, k, v], dim=-1)


def op_sparse_gemv(x: torch.Tensor, weight: torch.Tensor, threshold: float) -> torch.Tensor:
    """``SparseGEMV.forward`` dispatch (kernels/sparse_gemv.py:262-271): S==1 sparse, else dense
    ``x @ W.T`` with NO sparsification."""
    if x.shape[1] == 1:
        return sparse_linear(x, weight, threshold)
    return torch.matmul(x, weight.T)


def op_sparse_qkv_gemv(x, wqkv, tq, tk, tv, kv_size) -> torch.Tensor:
    """``SparseQKVGEMV.forward`` dispatch (kernels/sparse_gemv.py:288-298)."""
    if x.shape[1] == 1:
        return qkv_sparse_linear(x, wqkv, tq, tk, tv, kv_size)
    return torch.matmul(x, wqkv.T)


def ffn_forward(x, w1, w3, w2, t_gate, t_up, t_down) -> torch.Tensor:
    """``_new_ffn_forward`` gpt-fast/model.py:258-259 (== teal/mlp.py:50-56)."""
    gate = op_sparse_gemv(x, w1, t_gate)
    up = op_sparse_gemv(x, w3, t_up)
    return op_sparse_gemv(F.silu(gate) * up, w2, t_down)


# --------------------------------------------------------------------------- calibration lookup
class Distribution:
    """``Distribution`` gpt-fast/distribution.py:16-26,48-66 (== utils/utils.py:72-123).

    Built from the ``counts`` / ``centers`` fp32 tensors of a ``histograms.pt`` entry.
    """

    def __init__(self, centers: torch.Tensor, counts: torch.Tensor):
        self.bin_centers = torch.as_tensor(centers, dtype=torch.float32)
        self.counts = torch.as_tensor(counts, dtype=torch.float32)
        self.total_count = self.counts.sum()                       # distribution.py:25
        self.cumulative_counts = torch.cumsum(self.counts, dim=0)   # distribution.py:26

    @classmethod
    def from_file(cls, file_path: str, hidden_type: str) -> "Distribution":
        h = torch.load(os.path.join(file_path, "histograms.pt"))     # distribution.py:21
        return cls(h[f"{hidden_type}_centers"], h[hidden_type])

    def icdf(self, q: float) -> torch.Tensor:
        """distribution.py:48-66: linear interpolation in cumulative counts, clamped."""
        target = q * self.total_count
        idx = int(torch.searchsorted(self.cumulative_counts, target))
        if idx == 0:
            return self.bin_centers[0]
        if idx == len(self.bin_centers):
            return self.bin_centers[-1]
        lo_c, hi_c = self.cumulative_counts[idx - 1], self.cumulative_counts[idx]
        lo_v, hi_v = self.bin_centers[idx - 1], self.bin_centers[idx]
        frac = (target - lo_c) / (hi_c - lo_c)
        return lo_v + frac * (hi_v - lo_v)


def sparsity_to_threshold_hf(distr: Distribution, sparsity: float) -> float:
    """``SparsifyFn.set_threshold`` utils/utils.py:28-30 (exactly 0.0 when sparsity == 0)."""
    return distr.icdf(0.5 + sparsity / 2).item() if sparsity != 0.0 else 0.0


def sparsity_to_threshold_gptfast(distr: Distribution, sparsity: float) -> float:
    """gpt-fast loader: generate.py:278-287 (``icdf(0.5 + 0.5*s)``, no zero special-case)."""
    return distr.icdf(0.5 + 0.5 * sparsity).item()


# which histogram feeds which projection: generate.py:272-287
PROJ_HIST = {"q": ("self_attn", "h1"), "k": ("self_attn", "h1"), "v": ("self_attn", "h1"),
             "o": ("self_attn", "h2"), "gate": ("mlp", "h1"), "up": ("mlp", "h1"),
             "down": ("mlp", "h2")}


def layer_thresholds(hist_path: str, layer_idx: int, sparsities: Dict[str, float]) -> Dict[str, float]:
    """Per-projection thresholds of one layer as ``monkeypatch_layer`` computes them
    (generate.py:266-287)."""
    out = {}
    for proj in PROJS:
        sub, h = PROJ_HIST[proj]
        d = Distribution.from_file(os.path.join(hist_path, f"layer-{layer_idx}", sub), h)
        out[proj] = sparsity_to_threshold_gptfast(d, sparsities[proj])
    return out


def greedy_row_lookup(rows: Sequence[Dict[str, float]], target_sparsity: float) -> Dict[str, float]:
    """One layer of ``get_layer_greedy_sparsities`` utils/utils.py:251-256: the CSV row whose
    'Effective Sparsity' is closest to the target (first such row on ties: pandas argsort is
    stable-enough here because the column is strictly increasing)."""
    eff = np.array([r["Effective Sparsity"] for r in rows], dtype=np.float64)
    i = int(np.argsort(np.abs(eff - target_sparsity), kind="stable")[0])
    return {p: float(rows[i][p]) for p in PROJS}


def read_results_csv(path: str) -> List[Dict[str, float]]:
    with open(path, newline="") as f:
        return [{k: float(v) for k, v in row.items()} for row in csv.DictReader(f)]


def get_layer_greedy_sparsities(layer_sparsities: Sequence[float], results_dir: str) -> Dict[str, List[float]]:
    """utils/utils.py:243-258."""
    out = {p: [0.0] * len(layer_sparsities) for p in PROJS}
    for layer, target in enumerate(layer_sparsities):
        rows = read_results_csv(os.path.join(results_dir, f"layer-{layer}", "results.csv"))
        sel = greedy_row_lookup(rows, target)
        for p in PROJS:
            out[p][layer] = sel[p]
    return out


# --------------------------------------------------------------------------- tensor parallel
def tp_shard(w: torch.Tensor, dim: int, rank: int, world: int) -> torch.Tensor:
    """gpt-fast/tp.py:68-70: even ``tensor_split`` along ``dim``."""
    assert w.size(dim) % world == 0
    return torch.tensor_split(w, world, dim=dim)[rank]


def tp_shard_qkv(wqkv: torch.Tensor, splits: Sequence[int], rank: int, world: int) -> torch.Tensor:
    """gpt-fast/tp.py:72-78: shard q, k, v regions separately, re-concatenate per rank."""
    q, k, v = wqkv.split(list(splits), dim=0)
    return torch.cat((tp_shard(q, 0, rank, world), tp_shard(k, 0, rank, world),
                      tp_shard(v, 0, rank, world)), dim=0)


def tp_ffn_forward(x, w1, w3, w2, t_gate, t_up, t_down, world: int) -> torch.Tensor:
    """Megatron col->row FFN under TP (tp.py:110-121): sum over ranks of the rowwise partials,
    each rank's partial rounded to fp16 before the (fp16) all-reduce sum."""
    parts = []
    for r in range(world):
        parts.append(ffn_forward(x, tp_shard(w1, 0, r, world), tp_shard(w3, 0, r, world),
                                 tp_shard(w2, 1, r, world), t_gate, t_up, t_down))
    out = parts[0].clone()
    for p in parts[1:]:
        out = out + p
    return out


# --------------------------------------------------------------------------- gpt-fast decode model
def rms_norm(x: torch.Tensor, weight: torch.Tensor, eps: float) -> torch.Tensor:
    """gpt-fast/model.py:286-291."""
    xf = x.float()
    out = (xf * torch.rsqrt(torch.mean(xf * xf, dim=-1, keepdim=True) + eps)).type_as(x)
    return out * weight


def precompute_freqs_cis(seq_len: int, n_elem: int, base: float, dtype=torch.float16) -> torch.Tensor:
    """gpt-fast/model.py:294-303."""
    freqs = 1.0 / (base ** (torch.arange(0, n_elem, 2)[: (n_elem // 2)].float() / n_elem))
    t = torch.arange(seq_len)
    freqs = torch.outer(t, freqs)
    fc = torch.polar(torch.ones_like(freqs), freqs)
    return torch.stack([fc.real, fc.imag], dim=-1).to(dtype)


def apply_rotary_emb(x: torch.Tensor, freqs_cis: torch.Tensor) -> torch.Tensor:
    """gpt-fast/model.py:306-318 (adjacent-pair rotation, computed in fp32)."""
    xs = x.float().reshape(*x.shape[:-1], -1, 2)
    fc = freqs_cis.view(1, xs.size(1), 1, xs.size(3), 2)
    out = torch.stack([xs[..., 0] * fc[..., 0] - xs[..., 1] * fc[..., 1],
                       xs[..., 1] * fc[..., 0] + xs[..., 0] * fc[..., 1]], -1)
    return out.flatten(3).type_as(x)


class OracleDecoder:
    """Batch-1 gpt-fast Transformer with TEAL's sparse forwards (gpt-fast/model.py:99-190,
    258-259) restated on CPU.  ``weights`` uses gpt-fast names: tok_embeddings.weight,
    layers.{i}.attention.{wqkv,wo}.weight, layers.{i}.feed_forward.{w1,w2,w3}.weight,
    layers.{i}.{attention_norm,ffn_norm}.weight, norm.weight, output.weight.
    ``thresholds[i]`` maps proj name -> float (None => dense layer, old_forward)."""

    def __init__(self, cfg: dict, weights: Dict[str, torch.Tensor], thresholds, max_seq: int):
        self.cfg, self.w, self.thr = cfg, weights, thresholds
        self.n_layer, self.n_head, self.dim = cfg["n_layer"], cfg["n_head"], cfg["dim"]
        self.n_local_heads = cfg["n_local_heads"]
        self.head_dim = self.dim // self.n_head
        self.eps = cfg.get("norm_eps", 1e-5)
        dtype = weights["output.weight"].dtype
        self.freqs = precompute_freqs_cis(cfg.get("block_size", 2048), self.head_dim,
                                          cfg.get("rope_base", 10000), dtype)
        self.max_seq = max_seq
        shape = (1, self.n_local_heads, max_seq, self.head_dim)      # model.py:84
        self.k_cache = [torch.zeros(shape, dtype=dtype) for _ in range(self.n_layer)]
        self.v_cache = [torch.zeros(shape, dtype=dtype) for _ in range(self.n_layer)]
        self.causal = torch.tril(torch.ones(max_seq, max_seq, dtype=torch.bool))  # model.py:131

    def forward(self, idx: torch.Tensor, input_pos: torch.Tensor) -> torch.Tensor:
        """model.py:133-143.  idx [1,S] int, input_pos [S]."""
        w = self.w
        mask = self.causal[None, None, input_pos]
        fc = self.freqs[input_pos]
        x = F.embedding(idx, w["tok_embeddings.weight"])
        kv_size = self.n_local_heads * self.head_dim
        for i in range(self.n_layer):
            p = f"layers.{i}."
            t = self.thr[i]
            h = rms_norm(x, w[p + "attention_norm.weight"], self.eps)
            b, s, _ = h.shape
            if t is None:
                qkv = F.linear(h, w[p + "attention.wqkv.weight"])
            else:
                qkv = op_sparse_qkv_gemv(h, w[p + "attention.wqkv.weight"], t["q"], t["k"], t["v"], kv_size)
            q, k, v = qkv.split([self.dim, kv_size, kv_size], dim=-1)
            q = apply_rotary_emb(q.view(b, s, self.n_head, self.head_dim), fc)
            k = apply_rotary_emb(k.view(b, s, self.n_local_heads, self.head_dim), fc)
            v = v.view(b, s, self.n_local_heads, self.head_dim)
            q, k, v = (z.transpose(1, 2) for z in (q, k, v))
            self.k_cache[i][:, :, input_pos] = k                     # model.py:93-94
            self.v_cache[i][:, :, input_pos] = v
            rep = self.n_head // self.n_local_heads
            kk = self.k_cache[i].repeat_interleave(rep, dim=1)
            vv = self.v_cache[i].repeat_interleave(rep, dim=1)
            y = F.scaled_dot_product_attention(q, kk, vv, attn_mask=mask, dropout_p=0.0)
            y = y.transpose(1, 2).contiguous().view(b, s, self.dim)
            if t is None:
                y = F.linear(y, w[p + "attention.wo.weight"])
            else:
                y = op_sparse_gemv(y, w[p + "attention.wo.weight"], t["o"])
            hres = x + y
            g = rms_norm(hres, w[p + "ffn_norm.weight"], self.eps)
            if t is None:
                f = F.linear(F.silu(F.linear(g, w[p + "feed_forward.w1.weight"])) *
                             F.linear(g, w[p + "feed_forward.w3.weight"]), w[p + "feed_forward.w2.weight"])
            else:
                f = ffn_forward(g, w[p + "feed_forward.w1.weight"], w[p + "feed_forward.w3.weight"],
                                w[p + "feed_forward.w2.weight"], t["gate"], t["up"], t["down"])
            x = hres + f
        x = rms_norm(x, w["norm.weight"], self.eps)
        return F.linear(x, w["output.weight"])                       # model.py:141-142 (dense)


# --------------------------------------------------------------------------- cpu baseline helper
# ---------------------------------------------------------------------------------------------
# Sampling (gpt-fast/generate.py:49-66)
# ---------------------------------------------------------------------------------------------
def logits_to_probs(logits: torch.Tensor, temperature: float = 1.0, top_k: Optional[int] = None) -> torch.Tensor:
    """generate.py:53-61 restated with the arithmetic made explicit for an fp16 logits vector [V]: the division is
    fp32 (fp16 tensor / Python float) rounded to fp16; values below the top_k-th largest -> -inf; soft-max in fp32
    rounded once to fp16."""
    l = (logits.float() / np.float32(max(temperature, 1e-5))).to(logits.dtype)
    if top_k is not None and top_k > 0:
        v, _ = torch.topk(l, min(top_k, l.size(-1)))
        l = torch.where(l < v[..., -1:], torch.full_like(l, -float("inf")), l)
    return torch.softmax(l.float(), dim=-1).to(logits.dtype)


def sample_topk(logits: torch.Tensor, temperature: float, top_k: Optional[int], q: torch.Tensor):
    """generate.py:49-51,63-66 with the Exponential(1) noise `q` (fp16 [V]) given: token = argmax(probs / q),
    first index on ties.  Returns (token, probs)."""
    probs = logits_to_probs(logits, temperature, top_k)
    r = (probs.float() / q.float()).to(probs.dtype)
    return int(torch.argmax(r)), probs


def philox4x32_10(c0: int, c1: int, c2: int, c3: int, k0: int, k1: int):
    """Philox4x32-10 (Salmon et al. 2011), the counter-based generator of the library's sampling kernel."""
    M = 0xFFFFFFFF
    for _ in range(10):
        p0, p1 = 0xD2511F53 * c0, 0xCD9E8D57 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & M, p1 & M, ((p0 >> 32) ^ c3 ^ k1) & M, p0 & M
        k0, k1 = (k0 + 0x9E3779B9) & M, (k1 + 0xBB67AE85) & M
    return c0, c1, c2, c3


def sample_noise(seed: int, V: int, pos: int) -> torch.Tensor:
    """The library's own Exponential(1) noise: u = ((philox(idx, pos; seed)[0] >> 8) + 1) / 2^24 in (0, 1],
    q = fp16(-log u), zero replaced by the smallest fp16 subnormal."""
    r0 = np.array([philox4x32_10(i, pos, 0, 0, seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)[0] for i in range(V)], dtype=np.uint64)
    u = ((r0 >> np.uint64(8)).astype(np.float32) + np.float32(1.0)) * np.float32(1.0 / 16777216.0)
    q = torch.from_numpy(-np.log(u)).half()
    q[q == 0] = float(np.float16(2.0 ** -24))
    return q


def time_sparse_linear_cpu(x: torch.Tensor, weight: torch.Tensor, threshold: float, reps: int,
                           warmup: int = 2) -> Tuple[float, torch.Tensor]:
    """Times the reference CPU path (mask then dense F.linear -- it does not skip rows;
    BASELINE.md section 4).  Returns (seconds per call, last output)."""
    import time
    y = None
    for _ in range(warmup):
        y = sparse_linear(x, weight, threshold)
    t0 = time.perf_counter()
    for _ in range(reps):
        y = sparse_linear(x, weight, threshold)
    return (time.perf_counter() - t0) / max(reps, 1), y
