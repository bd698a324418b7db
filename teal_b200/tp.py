"""Tensor-parallel topology of the decode path (host side, pure torch -- testable on CPU with gloo).

Same topology as the reference's gpt-fast/tp.py: wqkv / w1 / w3 are column-sharded (``colwise``, the fused
wqkv per q/k/v region so each rank holds ``[q_r; k_r; v_r]``, tp.py:72-78), wo / w2 are row-sharded
(``rowwise``), heads and kv-heads are divided by the world size (tp.py:134-137), embeddings, norms and the
lm_head are replicated, and the only communication is one all-reduce(sum) of ``dim`` fp16 values after the
attention block and one after the MLP block (tp.py:120-121,139-140).
"""
from __future__ import annotations

import os
from typing import Dict, Optional, Sequence

import torch


def _get_rank() -> int:
    return int(os.environ.get("LOCAL_RANK", "0"))


def _get_world_size() -> int:
    return int(os.environ.get("LOCAL_WORLD_SIZE", os.environ.get("WORLD_SIZE", "1")))


def maybe_init_dist(backend: str = "nccl") -> Optional[int]:
    """gpt-fast/tp.py:36-51: process group from the torchrun environment; None when world < 2."""
    import torch.distributed as dist
    world = _get_world_size()
    if world < 2:
        return None
    rank = _get_rank()
    if backend == "nccl":
        torch.cuda.set_device(rank)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if not dist.is_initialized():
        dist.init_process_group(backend=backend, rank=int(os.environ.get("RANK", rank)),
                                world_size=int(os.environ.get("WORLD_SIZE", world)))
    return rank


def shard(x: torch.Tensor, dim: int, rank: int, world: int) -> torch.Tensor:
    """tp.py:68-70."""
    assert x.size(dim) % world == 0, "cannot shard evenly"
    return torch.tensor_split(x, world, dim=dim)[rank]


def shard_linear(weight: torch.Tensor, style: str, rank: int, world: int) -> torch.Tensor:
    """nn.Linear weight [out, in]: colwise shards out_features (dim 0), rowwise in_features (dim 1) (tp.py:60-66)."""
    return shard(weight, {"colwise": 0, "rowwise": 1}[style], rank, world)


def shard_qkv(wqkv: torch.Tensor, splits: Sequence[int], rank: int, world: int) -> torch.Tensor:
    """tp.py:72-78: shard q, k and v regions separately and re-concatenate per rank."""
    q, k, v = wqkv.split(list(splits), dim=0)
    return torch.cat((shard(q, 0, rank, world), shard(k, 0, rank, world), shard(v, 0, rank, world)), dim=0)


def shard_state_dict(sd: Dict[str, torch.Tensor], dim: int, n_head: int, n_local_heads: int, n_layer: int,
                     rank: int, world: int) -> Dict[str, torch.Tensor]:
    """Per-rank view of a gpt-fast state dict (apply_tp, tp.py:151-156)."""
    if world == 1:
        return dict(sd)
    hd = dim // n_head
    kv = n_local_heads * hd
    out = dict(sd)
    for i in range(n_layer):
        p = f"layers.{i}."
        out[p + "attention.wqkv.weight"] = shard_qkv(sd[p + "attention.wqkv.weight"], [dim, kv, kv], rank, world)
        out[p + "attention.wo.weight"] = shard_linear(sd[p + "attention.wo.weight"], "rowwise", rank, world)
        out[p + "feed_forward.w1.weight"] = shard_linear(sd[p + "feed_forward.w1.weight"], "colwise", rank, world)
        out[p + "feed_forward.w3.weight"] = shard_linear(sd[p + "feed_forward.w3.weight"], "colwise", rank, world)
        out[p + "feed_forward.w2.weight"] = shard_linear(sd[p + "feed_forward.w2.weight"], "rowwise", rank, world)
    return out


def all_reduce_sum(t: torch.Tensor, group=None) -> torch.Tensor:
    """The one collective of the path (tp.py:120-121,139-140)."""
    import torch.distributed as dist
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t
