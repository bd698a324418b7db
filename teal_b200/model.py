"""Batch-1 decode engine around the fused sparse GEMV kernels.

Mirrors the pieces of the reference's gpt-fast fork that sit on the north-star path:

* ``ModelArgs`` / ``transformer_configs`` -- gpt-fast/model.py:25-79 (shapes verbatim).
* ``SparseDecoder`` -- what ``_load_model`` + ``monkeypatch_layer`` + ``decode_one_token`` build
  (gpt-fast/generate.py:73-77,225-334; model.py:133-190,258-259), but as ONE captured CUDA graph of
  hand-written kernels per token: 5 launches per layer
  (RMSNorm+QKV-GEMV | RoPE+KV+attention | o-GEMV+residual | RMSNorm+gate/up-GEMV | SiLU*mul+down-GEMV+residual)
  + embedding, lm_head(+final norm, +argmax) and the step bookkeeping.  QKV and gate/up leave their fp32 split-K sums
  for attention / the down projection to round (deferred epilogue, include/teal_b200.h).
* weights live in the reference's "column major" layout ([Z][N] storage, generate.py:296-317).
* tensor parallelism follows gpt-fast/tp.py: wqkv/w1/w3 column-sharded (per-rank [q_r; k_r; v_r]),
  wo/w2 row-sharded, one exchange after the attention block and one after the MLP block -- fused into the o / down
  GEMV's tile finalisers over NVLink peer memory (default), or a one-shot peer-memory kernel, or ncclAllReduce.

PyTorch here is plumbing: device memory, the CUDA-graph capture and torch.distributed.
"""
from __future__ import annotations

import os

from dataclasses import dataclass
from typing import Dict, List, Optional

import numpy as np
import threading

import torch

from . import _lib, ops
from .distribution import Distribution, PROJS, threshold_for


def find_multiple(n: int, k: int) -> int:
    return n if n % k == 0 else n + k - (n % k)


@dataclass
class ModelArgs:
    block_size: int = 2048
    vocab_size: int = 32000
    n_layer: int = 32
    n_head: int = 32
    dim: int = 4096
    intermediate_size: Optional[int] = None
    n_local_heads: int = -1
    head_dim: int = 64
    rope_base: float = 10000
    norm_eps: float = 1e-5

    def __post_init__(self):
        if self.n_local_heads == -1:
            self.n_local_heads = self.n_head
        if self.intermediate_size is None:
            self.intermediate_size = find_multiple(int(2 * 4 * self.dim / 3), 256)
        self.head_dim = self.dim // self.n_head

    @classmethod
    def from_name(cls, name: str) -> "ModelArgs":
        if name in transformer_configs:
            return cls(**transformer_configs[name])
        cands = [c for c in transformer_configs if c.lower() in str(name).lower()]
        cands.sort(key=len, reverse=True)
        return cls(**transformer_configs[cands[0]])


# gpt-fast/model.py:66-79 (the shapes BASELINE.json's configs name)
transformer_configs = {
    "7B": dict(n_layer=32, n_head=32, dim=4096),
    "13B": dict(n_layer=40, n_head=40, dim=5120),
    "70B": dict(n_layer=80, n_head=64, dim=8192, n_local_heads=8, intermediate_size=28672),
    "Mistral-7B": dict(n_layer=32, n_head=32, n_local_heads=8, dim=4096, intermediate_size=14336, vocab_size=32000),
    "llama-3-8b": dict(block_size=8192, n_layer=32, n_head=32, n_local_heads=8, dim=4096, intermediate_size=14336,
                       vocab_size=128256, rope_base=500000),
    "llama-3-70b": dict(block_size=8192, n_layer=80, n_head=64, n_local_heads=8, dim=8192, intermediate_size=28672,
                        vocab_size=128256, rope_base=500000),
    # small shapes for tests / smoke
    "tiny": dict(block_size=256, n_layer=2, n_head=4, n_local_heads=2, dim=256, intermediate_size=704, vocab_size=512),
}


def precompute_freqs_cis(seq_len: int, n_elem: int, base: float, dtype=torch.float16) -> torch.Tensor:
    """Same table as gpt-fast/model.py:294-303 ((cos, sin) rounded to the model dtype)."""
    freqs = 1.0 / (base ** (torch.arange(0, n_elem, 2)[: (n_elem // 2)].float() / n_elem))
    t = torch.arange(seq_len)
    freqs = torch.outer(t, freqs)
    fc = torch.polar(torch.ones_like(freqs), freqs)
    return torch.stack([fc.real, fc.imag], dim=-1).to(dtype)


def fuse_qkv_state_dict(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """``Attention.load_hook`` (gpt-fast/model.py:209-214): checkpoints that still carry wq / wk / wv get the fused
    ``wqkv = cat([wq, wk, wv])`` the model (and the QKV kernel's region layout) expects."""
    out = dict(sd)
    for k in [k for k in sd if k.endswith("attention.wq.weight")]:
        p = k[: -len("wq.weight")]
        out[p + "wqkv.weight"] = torch.cat([out.pop(p + "wq.weight"), out.pop(p + "wk.weight"), out.pop(p + "wv.weight")])
    return out


def load_checkpoint(path: str) -> Dict[str, torch.Tensor]:
    """gpt-fast ``model.pth`` state dict (generate.py:244-247: mmap, weights only), QKV-fused."""
    sd = torch.load(path, mmap=True, weights_only=True)
    if "model" in sd and isinstance(sd["model"], dict):
        sd = sd["model"]
    return fuse_qkv_state_dict(sd)


def build_histogram(values: torch.Tensor, num_bins: int = 10000, outlier: float = 0.01):
    """Histogram in the reference's format (utils/utils.py:145-173: fp32 counts + bin centres, 1 % / 99 % outliers
    clipped into the two edge bins), computed ON THE DEVICE the activations live on (`calibration.find_histogram`: one
    sort + one searchsorted; bit-identical to the reference's histograms, tests/golden/histogram.npz).
    Returns (centers, counts) as numpy arrays for `Distribution`."""
    from .calibration import find_histogram
    counts, centers = find_histogram(values, num_bins, outlier)
    return centers.float().cpu().numpy(), counts.float().cpu().numpy()


class SparseDecoder:
    """Random-init (or state-dict) gpt-fast-shaped model decoding one token per CUDA-graph replay."""

    def __init__(self, args: ModelArgs, device="cuda", max_seq: int = 512, seed: int = 1234, rank: int = 0,
                 world: int = 1, process_group=None, state_dict: Optional[Dict[str, torch.Tensor]] = None,
                 out_cap: int = 4096):
        if not torch.cuda.is_available():
            raise _lib.TealLibraryError("SparseDecoder needs a CUDA device: the decode path has no CPU fallback")
        self.lib = _lib.load()
        self.args, self.dev, self.rank, self.world, self.pg = args, torch.device(device), rank, world, process_group
        a = args
        assert a.n_head % world == 0 and a.n_local_heads % world == 0 and a.intermediate_size % world == 0
        self.n_head_l, self.n_kv_l = a.n_head // world, a.n_local_heads // world   # tp.py:134-137
        self.hd = a.head_dim
        self.dim_l = self.n_head_l * self.hd                                      # local attention width
        self.kv_l = self.n_kv_l * self.hd
        self.nqkv_l = self.dim_l + 2 * self.kv_l
        self.inter_l = a.intermediate_size // world
        self.max_seq = find_multiple(max_seq, 8)                                  # model.py:118
        self.sparse = False
        self.thresholds: List[Optional[Dict[str, float]]] = [None] * a.n_layer
        self._build_weights(seed, state_dict)
        f16, dev = torch.float16, self.dev
        # one RoPE row per cache position: the reference caps max_seq_length at block_size (generate.py:171) -- longer
        # caches get their own rows instead of reading past the table
        self.freqs = precompute_freqs_cis(self.max_seq, self.hd, a.rope_base).to(dev).contiguous()
        self.steps_enqueued = 0  # host-side mirror of the device position (bounds check, see _check_room)
        self.k_cache = [torch.zeros(self.n_kv_l, self.max_seq, self.hd, dtype=f16, device=dev) for _ in range(a.n_layer)]
        self.v_cache = [torch.zeros(self.n_kv_l, self.max_seq, self.hd, dtype=f16, device=dev) for _ in range(a.n_layer)]
        # step state (device resident so the graph replays unchanged)
        self.token = torch.zeros(1, dtype=torch.int32, device=dev)
        self.pos = torch.zeros(1, dtype=torch.int32, device=dev)
        self.out_tokens = torch.zeros(out_cap, dtype=torch.int32, device=dev)
        self.argmax = torch.zeros(1, dtype=torch.int64, device=dev)
        self.x = torch.zeros(1, 1, a.dim, dtype=f16, device=dev)
        self.qkv = torch.zeros(1, 1, self.nqkv_l, dtype=f16, device=dev)
        self.att = torch.zeros(1, 1, self.dim_l, dtype=f16, device=dev)
        self.h = torch.zeros(1, 1, self.inter_l, dtype=f16, device=dev)
        self.part = torch.zeros(1, 1, a.dim, dtype=f16, device=dev)
        self.logits = torch.zeros(1, 1, a.vocab_size, dtype=f16, device=dev)
        self.ssq = torch.zeros(ops.SSQ_FLOATS, dtype=torch.float32, device=dev)  # replicated sum-of-squares partials of the residual stream
        self.stream = torch.cuda.Stream(device=dev)
        # the decoder's own accumulator / counter workspace (zeroed once, the kernels keep it zeroed): one decoder = one
        # stream of launches, so it never shares the per-stream slots ops.workspace() hands to ad-hoc op calls
        self.ws = torch.zeros(self.lib.teal_workspace_bytes(ops._MAX_OUT_COLS), dtype=torch.uint8, device=dev)
        self.tp_fused = False
        if world > 1:
            self._setup_peer_exchange()
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.graph_mode = None
        self.taps = None  # calibration taps
        self.force_tp_path = False  # tests: take the (plain GEMV -> all-reduce -> residual_add) route at world == 1
        # Deferred epilogue (include/teal_b200.h): the QKV and gate/up launches leave their fp32 split-K sums in
        # `defer_acc` ([q|k|v][gate|up]) and attention / the down projection round them on load; the o projection
        # re-zeroes the whole buffer once attention (the last reader) is done.  Same bits as the finaliser route
        # (tests compare the two); off for calibration taps and the cuBLAS / alternative TP exchange routes.
        self.deferred_epilogue = True
        self.defer_acc = torch.zeros(self.nqkv_l + 2 * self.inter_l, dtype=torch.float32, device=dev)
        self.ffn_inputs = None  # when a list: per-layer copies of (x, ssq) entering the MLP block (bench roofline)
        self.kernels_per_step = 0
        self.sampling = None  # None = greedy arg-max; else dict(temperature, top_k, seed): generate.py:49-66 on device

    # ------------------------------------------------------------------ TP exchange buffers
    def _setup_peer_exchange(self):
        """Symmetric (peer-mapped) memory for the TP exchange step: two fp16 partial buffers [dim] + a flag array per rank.
        TEAL_TP_ALLREDUCE = gemv (default: the exchange fused into the row-parallel GEMV's tile finalisers, push protocol) |
        fused (one-shot peer-memory kernel after a plain GEMV) | nccl (ncclAllReduce + residual add).  Every rank first agrees on what
        is available (all_reduce MIN) and only then enters the collective rendezvous, so a rank-local failure (import,
        environment) falls back to NCCL on all ranks together instead of hanging the others inside rendezvous."""
        import ctypes
        import torch.distributed as dist
        mode = os.environ.get("TEAL_TP_ALLREDUCE", "gemv")
        want = {"nccl": 0, "fused": 1, "gemv": 2}.get(mode, 0)
        symm = None
        if want:
            try:
                import torch.distributed._symmetric_memory as symm
            except Exception as e:  # noqa: BLE001
                want, self._symm_error = 0, repr(e)
        flag = torch.tensor([want], device=self.dev, dtype=torch.int32)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.pg)
        want = int(flag.item())
        self.tp_fused, self.tp_gemv = False, False
        if want == 0:
            if self.rank == 0:
                print(f"[teal_b200] TP exchange: NCCL all-reduce ({getattr(self, '_symm_error', 'TEAL_TP_ALLREDUCE=' + mode)})", flush=True)
            return
        dim, ws = self.args.dim, self.world
        pg = self.pg if self.pg is not None else dist.group.WORLD
        ok = 1
        try:  # from here on every rank makes the same collective calls; a failure is reported through the second vote
            if hasattr(symm, "enable_symm_mem_for_group"):
                try:
                    symm.enable_symm_mem_for_group(pg.group_name)
                except Exception:
                    pass
            t = symm.empty(4 * dim + 256, dtype=torch.uint8, device=self.dev)
            t.zero_()
            hdl = symm.rendezvous(t, pg)
            ptrs = [int(p) for p in hdl.buffer_ptrs]
            assert len(ptrs) == ws and all(ptrs)
            self._symm = (t, hdl)
            self.tp_bufs = [(ctypes.c_void_p * ws)(*[p + par * dim * 2 for p in ptrs]) for par in (0, 1)]
            self.tp_flags = (ctypes.c_void_p * ws)(*[p + 4 * dim for p in ptrs])
            self.tp_part = [t[par * dim * 2:(par + 1) * dim * 2].view(torch.float16).view(1, 1, dim) for par in (0, 1)]
            self.tp_epoch = torch.zeros(1, dtype=torch.int32, device=self.dev)
            if want == 2:  # per parity a receive buffer [world sources][dim / 2] 8-byte words per rank, local epochs [64]
                t2 = symm.empty(2 * ws * dim * 4, dtype=torch.uint8, device=self.dev)
                t2.zero_()
                hdl2 = symm.rendezvous(t2, pg)
                p2 = [int(p) for p in hdl2.buffer_ptrs]
                assert len(p2) == ws and all(p2)
                self._symm2 = (t2, hdl2)
                self.tp_recv = [(ctypes.c_void_p * ws)(*[p + par * ws * dim * 4 for p in p2]) for par in (0, 1)]
                self.tp_tile_epochs = [torch.zeros(64, dtype=torch.int32, device=self.dev) for _ in (0, 1)]
        except Exception as e:  # noqa: BLE001
            ok, self._symm_error = 0, repr(e)
        flag = torch.tensor([ok], device=self.dev, dtype=torch.int32)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.pg)
        torch.cuda.synchronize()
        if int(flag.item()):
            self.tp_fused, self.tp_gemv = True, want == 2
        elif self.rank == 0:
            print(f"[teal_b200] peer-memory TP exchange unavailable ({getattr(self, '_symm_error', 'a peer failed')}); "
                  "using NCCL all-reduce", flush=True)

    def _exchange(self, parity: int, sp):
        """The TP exchange step after a row-parallel GEMV whose partial is in tp_part[parity] / self.part."""
        a, lib = self.args, self.lib
        if self.tp_fused:
            _lib.check(lib.teal_tp_allreduce_residual_f16(self.tp_bufs[parity], self.tp_flags, self.rank, self.world,
                                                          self.x.data_ptr(), self.ssq.data_ptr(), self.tp_epoch.data_ptr(),
                                                          a.dim, sp), "tp_allreduce_residual")
            return 1
        self._allreduce(self.part)
        _lib.check(lib.teal_residual_add_f16(self.x.data_ptr(), self.part.data_ptr(), self.ssq.data_ptr(), a.dim, sp),
                   "residual_add")
        return 1

    # ------------------------------------------------------------------ weights
    def _colmajor_rand(self, g, n, z, std=0.02, style=None, splits=None):
        """fp16 [Z][N] storage viewed as the logical [N, Z] weight (stride (1, N)).  Under TP the FULL matrix is drawn
        (same seed on every rank => the same model everywhere) and this rank keeps its gpt-fast/tp.py shard: colwise =
        output columns (wqkv per q / k / v region, tp.py:72-78), rowwise = input lanes.  One matrix at a time, so the
        transient is one full projection (<= 470 MB for the 70B shapes), never the whole model."""
        from . import tp as _tp
        w = torch.empty(z, n, dtype=torch.float16, device=self.dev)
        w.normal_(0.0, std, generator=g)
        if self.world > 1 and style is not None:
            r, ws = self.rank, self.world
            if splits is not None:   # fused wqkv, storage [Z][q | k | v]: shard every region, re-concatenate
                w = torch.cat([_tp.shard(t, 1, r, ws) for t in w.split(list(splits), dim=1)], dim=1).contiguous()
            elif style == "colwise":
                w = _tp.shard(w, 1, r, ws).contiguous()
            else:                     # rowwise: this rank's input lanes
                w = _tp.shard(w, 0, r, ws).contiguous()
        return w.T

    def _build_weights(self, seed, sd):
        a, dev = self.args, self.dev
        g = torch.Generator(device=dev).manual_seed(seed)
        f16 = torch.float16
        r, w = self.rank, self.world

        from . import tp as _tp
        self.layers = []
        if sd is None:
            self.tok_emb = torch.empty(a.vocab_size, a.dim, dtype=f16, device=dev).normal_(0, 1.0, generator=g)
            self.lm_head = torch.empty(a.vocab_size, a.dim, dtype=f16, device=dev).normal_(0, 0.02, generator=g)
            self.final_norm = torch.ones(a.dim, dtype=f16, device=dev)
            kv = a.n_local_heads * a.head_dim
            for _ in range(a.n_layer):
                # ONE seeded model, sharded like gpt-fast/tp.py:54-156: every rank draws the same full matrices in the same
                # order and keeps its slice, so TP-N logits can be compared with the single-GPU model's
                L = dict(wqkv=self._colmajor_rand(g, a.dim + 2 * kv, a.dim, style="colwise", splits=(a.dim, kv, kv)),
                         wo=self._colmajor_rand(g, a.dim, a.dim, style="rowwise"),
                         w1=self._colmajor_rand(g, a.intermediate_size, a.dim, style="colwise"),
                         w3=self._colmajor_rand(g, a.intermediate_size, a.dim, style="colwise"),
                         w2=self._colmajor_rand(g, a.dim, a.intermediate_size, style="rowwise"),
                         attn_norm=torch.ones(a.dim, dtype=f16, device=dev), ffn_norm=torch.ones(a.dim, dtype=f16, device=dev))
                self.layers.append(L)
        else:
            cm = ops.to_column_major
            sd = _tp.shard_state_dict(sd, a.dim, a.n_head, a.n_local_heads, a.n_layer, r, w)
            self.tok_emb = sd["tok_embeddings.weight"].to(dev, f16).contiguous()
            self.lm_head = sd["output.weight"].to(dev, f16).contiguous()
            self.final_norm = sd["norm.weight"].to(dev, f16).contiguous()
            for i in range(a.n_layer):
                p = f"layers.{i}."
                L = dict(wqkv=cm(sd[p + "attention.wqkv.weight"].to(dev, f16)),
                         wo=cm(sd[p + "attention.wo.weight"].to(dev, f16)),
                         w1=cm(sd[p + "feed_forward.w1.weight"].to(dev, f16)),
                         w3=cm(sd[p + "feed_forward.w3.weight"].to(dev, f16)),
                         w2=cm(sd[p + "feed_forward.w2.weight"].to(dev, f16)),
                         attn_norm=sd[p + "attention_norm.weight"].to(dev, f16).contiguous(),
                         ffn_norm=sd[p + "ffn_norm.weight"].to(dev, f16).contiguous())
                self.layers.append(L)

    def weight_bytes(self):
        """(sparsifiable layer bytes, dense lm_head bytes) read per token on this rank."""
        a = self.args
        per_layer = 2 * (a.dim * self.nqkv_l + self.dim_l * a.dim + 2 * a.dim * self.inter_l + self.inter_l * a.dim)
        return per_layer * a.n_layer, 2 * a.vocab_size * a.dim

    # ------------------------------------------------------------------ thresholds
    def set_thresholds(self, thresholds: List[Optional[Dict[str, float]]]):
        """Per layer {q,k,v,o,gate,up,down} -> threshold (None = dense layer).  Invalidates the graph."""
        self.thresholds = list(thresholds)
        self.sparse = any(t is not None for t in thresholds)
        self.graph = None

    def calibrate(self, sparsities, n_tokens: int = 16, seed: int = 7):
        """Synthetic-model calibration through the reference's own path: run dense
        decode steps, histogram the four activation sites per layer (grab_acts -> histograms.pt),
        then threshold = Distribution.icdf(0.5 + s/2) per projection (generate.py:278-287).
        `sparsities`: float, or {proj: float}, or list (per layer) of {proj: float}."""
        a = self.args
        self.set_thresholds([None] * a.n_layer)
        acts = [dict(attn_h1=[], attn_h2=[], mlp_h1=[], mlp_h2=[]) for _ in range(a.n_layer)]
        self.taps = acts
        rng = np.random.RandomState(seed)
        self.reset(int(rng.randint(0, a.vocab_size)))
        for _ in range(n_tokens):
            self._enqueue_step(mode="dense_own")
            torch.cuda.synchronize()
            self.token.fill_(int(rng.randint(0, a.vocab_size)))  # random token stream
        self.taps = None
        site = {"q": "attn_h1", "k": "attn_h1", "v": "attn_h1", "o": "attn_h2", "gate": "mlp_h1", "up": "mlp_h1",
                "down": "mlp_h2"}
        thr = []
        for li in range(a.n_layer):
            sp = sparsities[li] if isinstance(sparsities, (list, tuple)) else sparsities
            if not isinstance(sp, dict):
                sp = {p: float(sp) for p in PROJS}
            dists = {k: Distribution(None, None, **dict(zip(("centers", "counts"), build_histogram(torch.cat(v)))))
                     for k, v in acts[li].items()}
            thr.append({p: threshold_for(dists[site[p]], sp[p]) for p in PROJS})
        self.set_thresholds(thr)
        return thr

    def set_sampling(self, temperature: float = 0.8, top_k: Optional[int] = 200, seed: int = 0):
        """Sample the next token like the reference's `sample()` (generate.py:49-66; its benchmark default is
        temperature 0.8, top_k 200 -- generate.py:353-354) instead of the greedy arg-max.  temperature == 0 / None =
        greedy (the reference itself divides by 1e-5 there and overflows fp16).  Invalidates the captured graph.
        The noise is the library's Philox stream keyed by (seed; vocabulary index, position): every TP rank draws the
        same token from its identical logits without any exchange."""
        self.sampling = None if not temperature else dict(temperature=float(temperature), top_k=int(top_k or 0), seed=int(seed))
        self.graph = None

    # ------------------------------------------------------------------ one decode step (enqueue only)
    def reset(self, first_token: int = 1):
        self.token.fill_(int(first_token))
        self.pos.zero_()
        self.argmax.zero_()
        self.defer_acc.zero_()
        self.steps_enqueued = 0

    def _allreduce(self, t):
        if self.world > 1:
            from . import tp as _tp
            _tp.all_reduce_sum(t, self.pg)

    def _enqueue_step(self, mode: str):
        """mode: 'sparse' (fused kernels with thresholds), 'dense_own' (same kernels, nothing dropped),
        'dense_cublas' (projections through torch.matmul like the reference's dense nn.Linear)."""
        a, lib, L = self.args, self.lib, self.layers
        sp = torch.cuda.current_stream().cuda_stream
        wsp, wsn = self.ws.data_ptr(), self.ws.numel()
        chk = _lib.check
        n_launch = 0
        ssq = self.ssq.data_ptr()
        chk(lib.teal_embed_f16(self.tok_emb.data_ptr(), self.token.data_ptr(), self.x.data_ptr(), ssq, a.dim, sp), "embed")
        n_launch += 1
        dense_t = {p: -1.0 for p in PROJS}
        for li in range(a.n_layer):
            W = L[li]
            t = self.thresholds[li] if (mode == "sparse" and self.thresholds[li] is not None) else dense_t
            tap = self.taps[li] if self.taps is not None else None
            one_kernel_rows = (self.world == 1 and not self.force_tp_path) or (self.tp_fused and getattr(self, "tp_gemv", False))
            defer = self.deferred_epilogue and tap is None and mode != "dense_cublas" and one_kernel_rows
            qkv_acc, gu_acc = self.defer_acc.data_ptr(), self.defer_acc.data_ptr() + 4 * self.nqkv_l
            if mode == "dense_cublas":
                xn = _rms_norm(self.x, W["attn_norm"], a.norm_eps)
                torch.matmul(xn, W["wqkv"].T, out=self.qkv)
            elif defer:
                chk(lib.teal_norm_sparse_qkv_gemv_acc_f16(self.x.data_ptr(), W["attn_norm"].data_ptr(), a.norm_eps, ssq,
                                                          W["wqkv"].data_ptr(), qkv_acc, a.dim, self.nqkv_l,
                                                          W["wqkv"].stride(1), self.kv_l, t["q"], t["k"], t["v"], sp), "norm_qkv")
            else:
                xn_out = torch.empty_like(self.x) if tap is not None else None
                chk(lib.teal_norm_sparse_qkv_gemv_f16(self.x.data_ptr(), W["attn_norm"].data_ptr(), a.norm_eps, ssq,
                                                      W["wqkv"].data_ptr(), self.qkv.data_ptr(),
                                                      xn_out.data_ptr() if tap is not None else None, a.dim, self.nqkv_l,
                                                      W["wqkv"].stride(1), self.kv_l, t["q"], t["k"], t["v"], wsp, wsn, sp),
                    "norm_qkv")
                if tap is not None:
                    tap["attn_h1"].append(xn_out.flatten().float())              # (activations stay on the device: the
            if True:  # attention block: RoPE + KV append + attention, then the o projection
                if defer:
                    chk(lib.teal_attn_decode_acc_f16(qkv_acc, self.freqs.data_ptr(), self.k_cache[li].data_ptr(),
                                                     self.v_cache[li].data_ptr(), self.pos.data_ptr(), self.att.data_ptr(),
                                                     self.n_head_l, self.n_kv_l, self.hd, self.max_seq, sp), "attn")
                else:
                    chk(lib.teal_attn_decode_f16(self.qkv.data_ptr(), self.freqs.data_ptr(), self.k_cache[li].data_ptr(),
                                                 self.v_cache[li].data_ptr(), self.pos.data_ptr(), self.att.data_ptr(),
                                                 self.n_head_l, self.n_kv_l, self.hd, self.max_seq, sp), "attn")
                if tap is not None:
                    tap["attn_h2"].append(self.att.flatten().float())            #  histograms are built there too)
                n_launch += 2
                if mode == "dense_cublas":
                    torch.matmul(self.att, W["wo"].T, out=self.part)
                    self._allreduce(self.part)
                    self.x.add_(self.part)
                    xn = _rms_norm(self.x, W["ffn_norm"], a.norm_eps)
                    g1 = torch.matmul(xn, W["w1"].T)
                    u1 = torch.matmul(xn, W["w3"].T)
                    torch.mul(torch.nn.functional.silu(g1), u1, out=self.h)
                    torch.matmul(self.h, W["w2"].T, out=self.part)
                    self._allreduce(self.part)
                    self.x.add_(self.part)
                    continue
                clear, clear_n = (self.defer_acc.data_ptr(), self.defer_acc.numel()) if defer else (None, 0)
                if self.world == 1 and not self.force_tp_path:
                    chk(lib.teal_sparse_gemv_residual_ex_f16(self.att.data_ptr(), None, W["wo"].data_ptr(), self.x.data_ptr(),
                                                             self.x.data_ptr(), ssq, self.dim_l, a.dim, W["wo"].stride(1), t["o"],
                                                             clear, clear_n, wsp, wsn, sp), "o_proj")
                elif self.tp_fused and getattr(self, "tp_gemv", False):  # GEMV + exchange + residual + ssq in one kernel
                    chk(lib.teal_sparse_gemv_tp_exchange_f16(self.att.data_ptr(), None, W["wo"].data_ptr(), self.x.data_ptr(), ssq,
                                                             self.dim_l, a.dim, W["wo"].stride(1), t["o"], clear, clear_n,
                                                             self.tp_recv[0], self.tp_tile_epochs[0].data_ptr(), self.rank,
                                                             self.world, wsp, wsn, sp), "o_proj+exchange")
                else:  # rowwise shard: partial sums, all-reduce over NVLink, then the residual add (tp.py:139-140)
                    dst = self.tp_part[0] if self.tp_fused else self.part
                    chk(lib.teal_sparse_gemv_f16(self.att.data_ptr(), W["wo"].data_ptr(), dst.data_ptr(), self.dim_l,
                                                 a.dim, W["wo"].stride(1), t["o"], wsp, wsn, sp), "o_proj")
                    n_launch += self._exchange(0, sp)
            if self.ffn_inputs is not None:
                self.ffn_inputs.append((self.x.clone(), self.ssq.clone()))
            if defer:
                chk(lib.teal_norm_sparse_gate_up_acc_f16(self.x.data_ptr(), W["ffn_norm"].data_ptr(), a.norm_eps, ssq,
                                                         W["w1"].data_ptr(), W["w3"].data_ptr(), gu_acc, a.dim, self.inter_l,
                                                         W["w1"].stride(1), W["w3"].stride(1), t["gate"], t["up"], sp), "gate_up")
            else:
                xn_out = torch.empty_like(self.x) if tap is not None else None
                chk(lib.teal_norm_sparse_gate_up_f16(self.x.data_ptr(), W["ffn_norm"].data_ptr(), a.norm_eps, ssq,
                                                     W["w1"].data_ptr(), W["w3"].data_ptr(), self.h.data_ptr(),
                                                     xn_out.data_ptr() if tap is not None else None, a.dim, self.inter_l,
                                                     W["w1"].stride(1), W["w3"].stride(1), t["gate"], t["up"], wsp, wsn, sp),
                    "gate_up")
            if tap is not None:
                tap["mlp_h1"].append(xn_out.flatten().float())
                tap["mlp_h2"].append(self.h.flatten().float())
            h_in, gu_in = (None, gu_acc) if defer else (self.h.data_ptr(), None)
            if self.world == 1 and not self.force_tp_path:
                chk(lib.teal_sparse_gemv_residual_ex_f16(h_in, gu_in, W["w2"].data_ptr(), self.x.data_ptr(),
                                                         self.x.data_ptr(), ssq, self.inter_l, a.dim, W["w2"].stride(1),
                                                         t["down"], None, 0, wsp, wsn, sp), "down_proj")
            elif self.tp_fused and getattr(self, "tp_gemv", False):
                chk(lib.teal_sparse_gemv_tp_exchange_f16(h_in, gu_in, W["w2"].data_ptr(), self.x.data_ptr(), ssq, self.inter_l, a.dim,
                                                         W["w2"].stride(1), t["down"], None, 0, self.tp_recv[1],
                                                         self.tp_tile_epochs[1].data_ptr(), self.rank, self.world, wsp, wsn, sp),
                    "down_proj+exchange")
            else:
                dst = self.tp_part[1] if self.tp_fused else self.part
                chk(lib.teal_sparse_gemv_f16(self.h.data_ptr(), W["w2"].data_ptr(), dst.data_ptr(), self.inter_l,
                                             a.dim, W["w2"].stride(1), t["down"], wsp, wsn, sp), "down_proj")
                n_launch += self._exchange(1, sp)
            n_launch += 3
        chk(lib.teal_dense_gemv_f16(self.x.data_ptr(), self.final_norm.data_ptr(), a.norm_eps, self.lm_head.data_ptr(),
                                    self.logits.data_ptr(), self.argmax.data_ptr(), a.dim, a.vocab_size, a.dim, sp),
            "lm_head")
        if self.sampling is not None:  # replaces the greedy arg-max in the packed cell that `advance` consumes
            sm = self.sampling
            chk(lib.teal_sample_topk_f16(self.logits.data_ptr(), a.vocab_size, sm["temperature"], sm["top_k"], None, sm["seed"],
                                         self.pos.data_ptr(), self.argmax.data_ptr(), None, sp), "sample")
            n_launch += 1
        chk(lib.teal_advance(self.argmax.data_ptr(), self.token.data_ptr(), self.pos.data_ptr(),
                             self.out_tokens.data_ptr(), self.out_tokens.numel(), sp), "advance")
        self.kernels_per_step = n_launch + 2

    # ------------------------------------------------------------------ public API
    def _check_room(self, n_new: int):
        """The device position advances without the host looking; the host mirrors it so that a step can never run past
        the KV cache / RoPE table / output buffer (the reference sizes its caches from the request, generate.py:171-176)."""
        if self.steps_enqueued + n_new > self.max_seq:
            raise ValueError(f"decode position would reach {self.steps_enqueued + n_new} > max_seq {self.max_seq}: "
                             "construct the decoder with a larger max_seq or reset()")
        if self.steps_enqueued + n_new > self.out_tokens.numel():
            raise ValueError(f"{self.steps_enqueued + n_new} tokens do not fit out_cap {self.out_tokens.numel()}")

    def capture(self, mode: str = "sparse", pdl: bool = True):
        """Capture one decode step into a CUDA graph (generate.py:420's reduce-overhead equivalent): one launch per
        projection group in a programmatic-dependent-launch chain.  The PDL switch is launch context of the calling
        thread (teal_set_pdl is thread-local), so decoders captured from different threads do not interfere."""
        if mode == "sparse" and not self.sparse:
            raise ValueError("no thresholds installed: call calibrate() / set_thresholds() or use a dense mode")
        self.lib.teal_set_pdl(1 if (pdl and mode != "dense_cublas") else 0)
        try:
            with _CAPTURE_LOCK, torch.cuda.stream(self.stream):
                saved = (self.token.clone(), self.pos.clone())
                self._enqueue_step(mode)  # warm-up outside capture (allocations, lazy init)
                self.stream.synchronize()
                self.token.copy_(saved[0]); self.pos.copy_(saved[1]); self.argmax.zero_()
                g = torch.cuda.CUDAGraph()
                # thread-local capture errors: other host threads (a second decoder capturing, the NCCL watchdog polling
                # events under TP) must not invalidate this capture
                with torch.cuda.graph(g, stream=self.stream, capture_error_mode="thread_local"):
                    self._enqueue_step(mode)
        finally:
            self.lib.teal_set_pdl(0)
        self.stream.synchronize()   # (stream-level: a device-wide synchronize is an error while another thread captures)
        self.graph, self.graph_mode = g, mode
        return g

    def decode_one_token(self):
        """Replay the captured step (device-side token/position; no host sync)."""
        self._check_room(1)
        self.graph.replay()
        self.steps_enqueued += 1

    def prefill(self, prompt, mode: str = "dense_own") -> None:
        """Feed a prompt (gpt-fast/generate.py:68-71: `prefill`).  The reference runs the prompt through the DENSE branch of
        the ops (S > 1 -> `x @ W.T`, no sparsification: kernels/sparse_gemv.py:271,298); here the prompt tokens go through
        the same decode step one position at a time with nothing dropped (`dense_own`), which is that computation up to
        fp16 summation order, fills the KV cache, and leaves the token sampled after the last prompt position in
        `self.token`.  The benchmark context is short (generate.py:171), so the prompt is not a throughput path."""
        toks = [int(t) for t in (prompt.tolist() if hasattr(prompt, "tolist") else prompt)]
        assert toks, "empty prompt"
        self.reset(toks[0])
        self._check_room(len(toks))
        with torch.cuda.stream(self.stream):
            for t in toks:
                self.token.fill_(t)        # teacher forcing: the step's own sample is overwritten by the next prompt token
                self._enqueue_step(mode)
        self.stream.synchronize()
        self.steps_enqueued = len(toks)

    def generate(self, prompt, n_new: int, mode: str = "sparse") -> torch.Tensor:
        """`prompt`: one token id (decode from it) or a sequence of ids (generate.py:149-217: prefill, then decode).
        Returns the `n_new` tokens sampled after the prompt's last position."""
        if self.graph is None or self.graph_mode != mode:
            self.capture(mode)
        many = not isinstance(prompt, int)
        if many and len(prompt) > 1:
            self.prefill(prompt)           # leaves the first new token in self.token, pos = len(prompt)
            first = self.token.clone()
            base = len(prompt)
            n_rest = n_new - 1
        else:
            self.reset(int(prompt[0]) if many else int(prompt))
            first, base, n_rest = None, 0, n_new
        self._check_room(n_rest)
        for _ in range(n_rest):
            self.graph.replay()
        self.steps_enqueued += n_rest
        torch.cuda.current_stream().synchronize()   # the replays' stream only (see capture())
        out = self.out_tokens[base:base + n_rest].clone()
        return out if first is None else torch.cat([first.to(out.dtype), out])


# Graph captures are serialised process-wide: torch's `torch.cuda.graph` entry synchronises the device, collects garbage
# and empties the allocator cache, which is not safe while another thread's capture is in flight.  Replays are not affected.
_CAPTURE_LOCK = threading.RLock()


def _rms_norm(x, w, eps):
    xf = x.float()
    return (xf * torch.rsqrt(torch.mean(xf * xf, dim=-1, keepdim=True) + eps)).to(x.dtype) * w
