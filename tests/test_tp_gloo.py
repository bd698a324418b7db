"""N > 1 host logic on CPU: world_size-2 gloo processes shard a gpt-fast state dict with the product's own
teal_b200/tp.py, run the (oracle) sparse blocks on their shards, all-reduce the row-parallel partials, and must
reproduce the unsharded oracle -- the contract the GPU TP path relies on."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import teal_oracle as O
from teal_b200 import tp

CFG = dict(block_size=64, vocab_size=96, n_layer=2, n_head=4, dim=64, n_local_heads=2, intermediate_size=160,
           rope_base=10000, norm_eps=1e-5)
THR = [dict(q=0.5, k=0.5, v=0.5, o=0.02, gate=0.5, up=0.4, down=0.01)] * 2


def _sd(cfg, seed=3):
    g = torch.Generator().manual_seed(seed)
    hd = cfg["dim"] // cfg["n_head"]
    nqkv = (cfg["n_head"] + 2 * cfg["n_local_heads"]) * hd
    sd = {"tok_embeddings.weight": torch.randn(cfg["vocab_size"], cfg["dim"], generator=g),
          "output.weight": torch.randn(cfg["vocab_size"], cfg["dim"], generator=g) * 0.1,
          "norm.weight": torch.ones(cfg["dim"])}
    for i in range(cfg["n_layer"]):
        p = f"layers.{i}."
        sd[p + "attention.wqkv.weight"] = torch.randn(nqkv, cfg["dim"], generator=g) * 0.1
        sd[p + "attention.wo.weight"] = torch.randn(cfg["dim"], cfg["dim"], generator=g) * 0.1
        sd[p + "feed_forward.w1.weight"] = torch.randn(cfg["intermediate_size"], cfg["dim"], generator=g) * 0.1
        sd[p + "feed_forward.w3.weight"] = torch.randn(cfg["intermediate_size"], cfg["dim"], generator=g) * 0.1
        sd[p + "feed_forward.w2.weight"] = torch.randn(cfg["dim"], cfg["intermediate_size"], generator=g) * 0.1
        sd[p + "attention_norm.weight"] = torch.ones(cfg["dim"])
        sd[p + "ffn_norm.weight"] = torch.ones(cfg["dim"])
    return {k: v.half() for k, v in sd.items()}


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank), LOCAL_WORLD_SIZE=str(world))
    assert tp.maybe_init_dist("gloo") == rank
    sd = _sd(CFG)
    loc = tp.shard_state_dict(sd, CFG["dim"], CFG["n_head"], CFG["n_local_heads"], CFG["n_layer"], rank, world)
    hd = CFG["dim"] // CFG["n_head"]
    x = torch.randn(1, 1, CFG["dim"], generator=torch.Generator().manual_seed(11)).half()
    t = THR[0]
    # col-parallel QKV: every rank thresholds the replicated x identically and gets its [q_r; k_r; v_r] slice
    kv_l = CFG["n_local_heads"] // world * hd
    qkv_l = O.op_sparse_qkv_gemv(x, loc["layers.0.attention.wqkv.weight"], t["q"], t["k"], t["v"], kv_l)
    # row-parallel o-proj on this rank's slice of a fake attention output, then the all-reduce (tp.py:139-140)
    y_full = torch.randn(1, 1, CFG["dim"], generator=torch.Generator().manual_seed(12)).half() * 0.1
    y_loc = tp.shard(y_full, 2, rank, world)
    part = O.op_sparse_gemv(y_loc, loc["layers.0.attention.wo.weight"], t["o"]).float()
    tp.all_reduce_sum(part)
    # MLP block: col-parallel gate/up, row-parallel down, all-reduce (tp.py:110-121)
    f = O.ffn_forward(x, loc["layers.0.feed_forward.w1.weight"], loc["layers.0.feed_forward.w3.weight"],
                      loc["layers.0.feed_forward.w2.weight"], t["gate"], t["up"], t["down"]).float()
    tp.all_reduce_sum(f)
    q.put((rank, qkv_l.numpy(), part.numpy(), f.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_tp2_matches_unsharded_oracle():
    world = 2
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    sd = _sd(CFG)
    hd = CFG["dim"] // CFG["n_head"]
    kv = CFG["n_local_heads"] * hd
    x = torch.randn(1, 1, CFG["dim"], generator=torch.Generator().manual_seed(11)).half()
    t = THR[0]
    full_qkv = O.op_sparse_qkv_gemv(x, sd["layers.0.attention.wqkv.weight"], t["q"], t["k"], t["v"], kv)
    q_f, k_f, v_f = full_qkv.split([CFG["dim"], kv, kv], dim=-1)
    for r, qkv_l, _, _ in res:
        want = torch.cat([tp.shard(q_f, 2, r, world), tp.shard(k_f, 2, r, world), tp.shard(v_f, 2, r, world)], dim=-1)
        assert np.array_equal(qkv_l, want.numpy())  # column shards are bit-identical slices
    y_full = torch.randn(1, 1, CFG["dim"], generator=torch.Generator().manual_seed(12)).half() * 0.1
    want_o = O.op_sparse_gemv(y_full, sd["layers.0.attention.wo.weight"], t["o"]).float().numpy()
    want_f = O.ffn_forward(x, sd["layers.0.feed_forward.w1.weight"], sd["layers.0.feed_forward.w3.weight"],
                           sd["layers.0.feed_forward.w2.weight"], t["gate"], t["up"], t["down"]).float().numpy()
    for _, _, part, f in res:
        assert np.allclose(part, want_o, rtol=1e-2, atol=2e-3)  # ws fp16 roundings of the partial sums
        assert np.allclose(f, want_f, rtol=1e-2, atol=2e-3)
    assert np.array_equal(res[0][2], res[1][2]) and np.array_equal(res[0][3], res[1][3])  # all ranks agree


def test_library_exports_every_declared_symbol():
    """C-ABI load + symbol check without a GPU (no compute calls)."""
    import ctypes
    from teal_b200 import _lib, build
    lib = ctypes.CDLL(build.build())
    decl = _lib.declared_symbols()
    assert len(decl) >= 20
    for s in decl:
        assert hasattr(lib, s), s
    assert lib.teal_abi_version() == _lib.header_abi_version() and lib.teal_compiled_arch() == 100
    lib.teal_workspace_bytes.restype = ctypes.c_size_t
    assert lib.teal_workspace_bytes(4096) >= 4096 * 4
    # every declared symbol has a ctypes signature in the Python binding
    assert set(decl) <= set(_lib._SIGS), set(decl) - set(_lib._SIGS)


def test_ops_fail_loudly_without_cuda():
    import pytest
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import teal_b200.ops as ops
    from teal_b200 import _lib
    x = torch.zeros(1, 1, 64, dtype=torch.float16)
    W = torch.zeros(64, 32, dtype=torch.float16).T
    with pytest.raises(_lib.TealLibraryError):
        ops.splitk_sparse_gemv(x, W, 0.5, 0)
    from teal_b200.model import ModelArgs, SparseDecoder
    with pytest.raises(_lib.TealLibraryError):
        SparseDecoder(ModelArgs.from_name("tiny"))
