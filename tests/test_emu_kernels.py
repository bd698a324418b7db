"""The CUDA kernels' LOGIC on the CPU: the production .cu sources compiled as C++ against tests/emu/cuda_emu.h (fibers
for threads, real atomics, warp collectives through per-warp slots) and driven through the same C ABI, checked against
the oracle.  This is how kernels that have not run on a B200 yet (generation-2 row-chunk kernel, experimental
step-program paths) are exercised in the CPU suite; the hardware-validated kernels run through the same emulation
first, which validates the emulator itself.  It proves indexing / masks / pipelines / protocols, not the memory model
and not speed.  The -m gpu tests remain the parity tests proper."""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import teal_oracle as O

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu"))
import harness as H  # noqa: E402

V1_CFGS = [30, 31, 32, 33, 35, 38, 40]      # hardware-validated row-chunk variants (emulator self-check)
GEN2_CFGS = [42, 43, 44, 45, 46, 47, 48, 49, 50, 51, 52, 53]   # 50+: warp-wise completion


@pytest.fixture(autouse=True)
def _defaults():
    H.lib()
    H.set_config(0)
    H.set_tuning()
    H.lib().teal_emu_set_sm_count(4)
    H.lib().teal_emu_set_concurrent(0)
    yield
    H.set_config(0)
    H.set_tuning()


def _mk(Z, N, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    x = (torch.randn(1, 1, Z, generator=g) * scale).half()
    W = (torch.randn(N, Z, generator=g) * 0.05).half()
    return x, W


def _tol(y32):
    rms = float(y32.float().pow(2).mean().sqrt())
    return y32.abs() * 2.0 ** -10 + 2.5e-4 * max(rms, 1e-6) + 1e-7


def _gemv(x, W, thr, ws):
    N, Z = W.shape
    Wc = H.column_major(W.numpy())
    y = H.aligned((N,), np.float16, 0)
    xa = H.aligned((Z,), np.float16, x.reshape(-1).numpy())
    H.check(H.lib().teal_sparse_gemv_f16(H.ptr(xa), H.ptr(Wc), H.ptr(y), Z, N, N, thr, H.ptr(ws.buf), ws.nbytes, None), "gemv")
    ws.assert_clean()
    return torch.from_numpy(y.copy())


def _check_gemv(Z, N, thr, seed=0):
    x, W = _mk(Z, N, seed)
    ws = H.Workspace()
    y = _gemv(x, W, thr, ws)
    want = O.sparse_linear_f32(x, W, thr).reshape(-1)
    err = (y.float() - want).abs()
    assert (err <= _tol(want)).all(), (Z, N, thr, float(err.max()))


SHAPES = [(200, 264), (512, 1024), (1000, 776), (96, 8), (2048, 520)]


@pytest.mark.parametrize("cfg", V1_CFGS)
def test_emulator_reproduces_the_validated_kernels(cfg):
    H.set_config(cfg)
    for Z, N in SHAPES[:3]:
        _check_gemv(Z, N, 0.6745, seed=cfg)


@pytest.mark.parametrize("cfg", GEN2_CFGS)
@pytest.mark.parametrize("spec", [(0, 1), (4, 1), (8, 2), (32, 1)])
def test_gen2_gemv_matches_oracle(cfg, spec):
    H.set_config(cfg)
    H.set_tuning(generation=2, spec_rows=spec[0], spec_mode=spec[1])
    for Z, N in SHAPES:
        _check_gemv(Z, N, 0.6745, seed=cfg + spec[0])


@pytest.mark.parametrize("cfg", [43, 45, 46, 50, 53])
@pytest.mark.parametrize("thr", [0.0, 100.0, 0.05, 2.0])
def test_gen2_sparsity_extremes(cfg, thr):
    H.set_config(cfg)
    H.set_tuning(generation=2, spec_rows=4)
    _check_gemv(1000, 776, thr, seed=3)


@pytest.mark.parametrize("gen", [1, 2, 3])
def test_default_dispatch_and_survivors_bit_exact(gen):
    H.set_tuning(generation=gen)
    for Z, thr in [(8, 0.5), (200, 0.3), (4096, 0.6745), (4096, 0.0), (4096, 100.0), (11008, 0.0203768)]:
        g = torch.Generator().manual_seed(Z)
        x = torch.randn(1, 1, Z, generator=g).half()
        if thr < 0.1 and thr > 0:
            x = (x.float() * 0.03).half()
        xa = H.aligned((Z,), np.float16, x.reshape(-1).numpy())
        idx = H.aligned((Z,), np.int32, 0)
        cnt = H.aligned((1,), np.int32, 0)
        H.check(H.lib().teal_survivors_f16(H.ptr(xa), Z, thr, H.ptr(idx), H.ptr(cnt), None), "survivors")
        want = O.survivor_indices(x, thr)
        assert int(cnt[0]) == len(want) and np.array_equal(idx[: len(want)], want), (gen, Z, thr)
    _check_gemv(512, 1024, 0.6745)


def _tie_vector():
    # |x| == fp16(t) with fp16(t) > t: the oracle (fp16 compare) drops it
    t = 0.02038193
    t16 = float(torch.tensor(t, dtype=torch.float16))
    x = torch.tensor([t16, -t16, t16 * 1.001, 0.0, float("nan"), -0.0, 1.0, -1.0], dtype=torch.float16).reshape(1, 1, -1)
    return x, t


@pytest.mark.parametrize("gen", [1, 2, 3])
def test_tie_rule_and_nan(gen):
    H.set_tuning(generation=gen)
    x, t = _tie_vector()
    Z = x.numel()
    xa = H.aligned((Z,), np.float16, x.reshape(-1).numpy())
    idx = H.aligned((Z,), np.int32, 0)
    cnt = H.aligned((1,), np.int32, 0)
    H.check(H.lib().teal_survivors_f16(H.ptr(xa), Z, t, H.ptr(idx), H.ptr(cnt), None), "survivors")
    got = idx[: int(cnt[0])].tolist()
    assert got == [2, 4, 6, 7]          # the two ties and the zeros are dropped, NaN is gathered (poisons y like mask*x)


@pytest.mark.parametrize("gen", [1, 2, 3])
@pytest.mark.parametrize("uniform", [True, False])
def test_qkv_region_thresholds(gen, uniform):
    H.set_tuning(generation=gen, spec_rows=4 if gen >= 2 else 0)
    Z, N, kv = 256, 640, 128                       # N_q = 384: the k/v regions do not start on a 256-column tile
    x, W = _mk(Z, N, 5)
    tq, tk, tv = (0.5, 0.5, 0.5) if uniform else (0.3, 0.8, 1.2)
    Wc = H.column_major(W.numpy())
    y = H.aligned((N,), np.float16, 0)
    xa = H.aligned((Z,), np.float16, x.reshape(-1).numpy())
    ws = H.Workspace()
    H.check(H.lib().teal_sparse_qkv_gemv_f16(H.ptr(xa), H.ptr(Wc), H.ptr(y), Z, N, N, kv, tq, tk, tv, H.ptr(ws.buf), ws.nbytes, None), "qkv")
    ws.assert_clean()
    want = torch.cat([O.sparse_linear_f32(x, W[: N - 2 * kv], tq).reshape(-1), O.sparse_linear_f32(x, W[N - 2 * kv: N - kv], tk).reshape(-1),
                      O.sparse_linear_f32(x, W[N - kv:], tv).reshape(-1)])
    assert ((torch.from_numpy(y.copy()).float() - want).abs() <= _tol(want)).all()


@pytest.mark.parametrize("gen", [1, 2, 3])
def test_fused_layer_half_steps(gen):
    """RMSNorm+gate/up+SiLU*mul -> down+residual (+ sum of squares) -> RMSNorm+QKV from the partials: the decode step's
    GEMV chain (gpt-fast/model.py:159-161,258-259,286-291) on one small layer."""
    H.set_tuning(generation=gen, spec_rows=8 if gen >= 2 else 0, spec_mode=2)
    L = H.lib()
    dim, inter, eps = 256, 776, 1e-5
    g = torch.Generator().manual_seed(11)
    resid = torch.randn(1, 1, dim, generator=g).half()
    nw = (1 + 0.1 * torch.randn(dim, generator=g)).half()
    W1, W3 = [(torch.randn(inter, dim, generator=g) * 0.05).half() for _ in range(2)]
    W2 = (torch.randn(dim, inter, generator=g) * 0.05).half()
    tg, tu, td = 0.4, 0.6, 0.02
    ws = H.Workspace()
    xa = H.aligned((dim,), np.float16, resid.reshape(-1).numpy())
    nwa = H.aligned((dim,), np.float16, nw.numpy())
    h = H.aligned((inter,), np.float16, 0)
    xn = H.aligned((dim,), np.float16, 0)
    W1c, W3c, W2c = H.column_major(W1.numpy()), H.column_major(W3.numpy()), H.column_major(W2.numpy())
    # ssq_in = NULL: the library computes the partials itself (ssq_kernel) into the workspace scratch
    H.check(L.teal_norm_sparse_gate_up_f16(H.ptr(xa), H.ptr(nwa), eps, None, H.ptr(W1c), H.ptr(W3c), H.ptr(h), H.ptr(xn), dim, inter,
                                           inter, inter, tg, tu, H.ptr(ws.buf), ws.nbytes, None), "norm_gate_up")
    ws.assert_clean()
    xn_ref = O.rms_norm(resid, nw, eps)
    assert np.array_equal(xn.view(np.uint16), xn_ref.reshape(-1).numpy().view(np.uint16)) or \
        float((torch.from_numpy(xn.copy()).float() - xn_ref.reshape(-1).float()).abs().max()) <= 2e-3
    xn_t = torch.from_numpy(xn.copy()).reshape(1, 1, -1)
    g32, u32 = O.sparse_linear_f32(xn_t, W1, tg), O.sparse_linear_f32(xn_t, W3, tu)
    h_ref = (torch.nn.functional.silu(g32.half()) * u32.half()).reshape(-1)           # fp16 silu, fp16 product (model.py:259)
    h_t = torch.from_numpy(h.copy())
    assert ((h_t.float() - h_ref.float()).abs() <= h_ref.float().abs() * 2.0 ** -8 + 2e-3).all()
    # down + residual + sum of squares
    y = H.aligned((dim,), np.float16, 0)
    ssq = H.aligned((64,), np.float32, 7.0)                                            # stale garbage must be overwritten
    H.check(L.teal_sparse_gemv_residual_f16(H.ptr(h), H.ptr(W2c), H.ptr(xa), H.ptr(y), H.ptr(ssq), inter, dim, dim, td,
                                            H.ptr(ws.buf), ws.nbytes, None), "down_residual")
    ws.assert_clean()
    d32 = O.sparse_linear_f32(h_t.reshape(1, 1, -1), W2, td).reshape(-1)
    want = (resid.reshape(-1) + d32.half()).float()
    y_t = torch.from_numpy(y.copy())
    assert ((y_t.float() - want).abs() <= want.abs() * 2.0 ** -9 + 2e-3).all()
    assert abs(float(ssq.sum()) - float((y_t.float() ** 2).sum())) <= 1e-4 * float((y_t.float() ** 2).sum())
    # consumer: RMSNorm from the partials + QKV
    N, kv = 384, 64
    Wq = (torch.randn(N, dim, generator=g) * 0.05).half()
    Wqc = H.column_major(Wq.numpy())
    q = H.aligned((N,), np.float16, 0)
    xn2 = H.aligned((dim,), np.float16, 0)
    H.check(L.teal_norm_sparse_qkv_gemv_f16(H.ptr(y), H.ptr(nwa), eps, H.ptr(ssq), H.ptr(Wqc), H.ptr(q), H.ptr(xn2), dim, N, N, kv,
                                            0.5, 0.5, 0.5, H.ptr(ws.buf), ws.nbytes, None), "norm_qkv")
    ws.assert_clean()
    xn2_ref = O.rms_norm(y_t.reshape(1, 1, -1), nw, eps).reshape(-1)
    assert float((torch.from_numpy(xn2.copy()).float() - xn2_ref.float()).abs().max()) <= 4e-3
    q_ref = O.sparse_linear_f32(torch.from_numpy(xn2.copy()).reshape(1, 1, -1), Wq, 0.5).reshape(-1)
    assert ((torch.from_numpy(q.copy()).float() - q_ref).abs() <= _tol(q_ref)).all()


# ------------------------------------------------------------------------------------------------ whole decode step
CFG = dict(block_size=256, vocab_size=512, n_layer=2, n_head=4, dim=256, n_local_heads=2, intermediate_size=704,
           rope_base=10000, norm_eps=1e-5)
THR = [dict(q=0.55, k=0.4, v=0.7, o=0.02, gate=0.6, up=0.5, down=0.01),
       dict(q=0.6, k=0.6, v=0.6, o=0.03, gate=0.5, up=0.5, down=0.02)]
TOKS = [3, 17, 42, 8]


def _state_dict(cfg, seed=0):   # the model of tests/test_decoder_gpu.py
    g = torch.Generator().manual_seed(seed)
    hd = cfg["dim"] // cfg["n_head"]
    nqkv = (cfg["n_head"] + 2 * cfg["n_local_heads"]) * hd
    sd = {"tok_embeddings.weight": torch.randn(cfg["vocab_size"], cfg["dim"], generator=g),
          "output.weight": torch.randn(cfg["vocab_size"], cfg["dim"], generator=g) * 0.05,
          "norm.weight": 1 + 0.1 * torch.randn(cfg["dim"], generator=g)}
    for i in range(cfg["n_layer"]):
        p = f"layers.{i}."
        sd[p + "attention.wqkv.weight"] = torch.randn(nqkv, cfg["dim"], generator=g) * 0.05
        sd[p + "attention.wo.weight"] = torch.randn(cfg["dim"], cfg["dim"], generator=g) * 0.05
        sd[p + "feed_forward.w1.weight"] = torch.randn(cfg["intermediate_size"], cfg["dim"], generator=g) * 0.05
        sd[p + "feed_forward.w3.weight"] = torch.randn(cfg["intermediate_size"], cfg["dim"], generator=g) * 0.05
        sd[p + "feed_forward.w2.weight"] = torch.randn(cfg["dim"], cfg["intermediate_size"], generator=g) * 0.05
        sd[p + "attention_norm.weight"] = 1 + 0.1 * torch.randn(cfg["dim"], generator=g)
        sd[p + "ffn_norm.weight"] = 1 + 0.1 * torch.randn(cfg["dim"], generator=g)
    return {k: v.half() for k, v in sd.items()}


_ORACLE_LOGITS = {}


def _oracle_logits(sparse):
    if sparse not in _ORACLE_LOGITS:
        orc = O.OracleDecoder(CFG, _state_dict(CFG), THR if sparse else [None] * CFG["n_layer"], max_seq=64)
        with torch.no_grad():
            _ORACLE_LOGITS[sparse] = [orc.forward(torch.tensor([[t]]), torch.tensor([i])).float().reshape(-1) for i, t in enumerate(TOKS)]
    return _ORACLE_LOGITS[sparse]


def _run_and_compare(dec, sparse, step):
    want = _oracle_logits(sparse)
    dec.reset(TOKS[0])
    for i, t in enumerate(TOKS):
        dec.token[0] = t
        step()
        got = torch.from_numpy(dec.logits.astype(np.float32))
        scale = float(want[i].abs().max())
        tol = (6e-2 if sparse else 2e-2) * scale + 2e-2             # the bound of tests/test_step_program_gpu.py
        assert float((got - want[i]).abs().max()) <= tol, (i, float((got - want[i]).abs().max()), tol)
        assert int(dec.pos[0]) == i + 1 and int(dec.out_tokens[i]) == int(dec.token[0])
        assert int(dec.token[0]) == int(torch.argmax(got))         # fused arg-max == torch.argmax of the fp16 logits


@pytest.mark.parametrize("gen,sparse", [(1, False), (1, True), (2, False), (2, True), (3, False), (3, True), (20, True)])
def test_decode_step_launch_per_projection(gen, sparse):
    if gen == 20:   # generation 2 with conversions + FFMA2 instead of FHFMA
        H.set_tuning(generation=2, fhfma=0)
    else:
        H.set_tuning(generation=gen, spec_rows=4 if gen >= 2 else 0)
    dec = H.EmuDecoder(CFG, _state_dict(CFG), THR)
    _run_and_compare(dec, sparse, lambda: dec.step(sparse))


PROGRAM_KNOBS = [
    dict(),                                                        # hardware-validated default (emulator self-check)
    dict(rows_small=32, rows_large=64), dict(group_qkv=1, group_gate_up=2), dict(wide_melems=1), dict(head_rows=7),
    # written after the GPU budget was spent -- never run on hardware:
    dict(warps=2), dict(warps=2, rows_small=32), dict(opts=2), dict(opts=3, stagger_ns=100), dict(warps=2, opts=3, stagger_ns=50),
    dict(rows_qkv=32, rows_o=128, rows_gate_up=64, rows_down=64), dict(opts=4), dict(opts=5), dict(warps=2, opts=7, stagger_ns=50),
    dict(tiles_per_item=2), dict(tiles_per_item=4, warps=2, rows_small=64, rows_large=64),
    dict(tiles_per_item=3, group_qkv=5, group_gate_up=6), dict(tiles_per_item=2, wide_melems=1, opts=4),
]


@pytest.mark.parametrize("knobs", PROGRAM_KNOBS, ids=lambda k: ",".join(f"{a}={b}" for a, b in k.items()) or "default")
def test_step_program_one_kernel_per_token(knobs):
    H.lib().teal_emu_set_sm_count(2)
    H.lib().teal_emu_set_concurrent(1)          # persistent kernel: all CTAs alive together, they wait on each other
    dec = H.EmuDecoder(CFG, _state_dict(CFG), THR)
    dec.build_program(True, **knobs)
    _run_and_compare(dec, True, dec.step_program)


def test_decode_step_with_device_sampling():
    """lm_head -> sample_topk -> advance (SparseDecoder.set_sampling): the sampled token is the oracle's `sample_topk` of the
    kernel's own logits under the library's noise stream, and it is what the next step embeds."""
    dec = H.EmuDecoder(CFG, _state_dict(CFG), THR)
    dec.sampling = dict(temperature=0.8, top_k=20, seed=77)
    dec.reset(5)
    dec.token[0] = 5
    for i in range(4):
        fed = int(dec.token[0])
        dec.step(True)
        logits = torch.from_numpy(dec.logits.copy())
        want, _ = O.sample_topk(logits, 0.8, 20, O.sample_noise(77, CFG["vocab_size"], i))
        assert int(dec.token[0]) == want and int(dec.out_tokens[i]) == want, (i, fed)
        assert int(dec.pos[0]) == i + 1


def test_timeline_twins_record_every_cta_and_do_not_change_results():
    """teal_set_debug_timeline: the TIMED twins of the two default kernels write monotone stamps for every CTA into one
    slot per launch; results are those of the production instantiation."""
    L = H.lib()
    slots, slot_ints = 3, 4096 * 16
    buf = H.aligned((slots * slot_ints,), np.int64, 0)
    x, W = _mk(512, 1024, 9)
    ws = H.Workspace()
    y_plain = _gemv(x, W, 0.6745, ws)
    H.check(L.teal_set_debug_timeline(H.ptr(buf), slots), "timeline")
    try:
        y_timed = _gemv(x, W, 0.6745, ws)          # small launch -> variant 2 twin, slot 0
        _check_gemv(1000, 776, 0.5, seed=4)         # slot 1
    finally:
        H.check(L.teal_set_debug_timeline(None, 0), "timeline off")
    assert torch.equal(y_plain, y_timed) or float((y_plain.float() - y_timed.float()).abs().max()) < 2e-3
    t = buf.reshape(slots, 4096, 16)
    for s, (Z, N) in enumerate([(512, 1024), (1000, 776)]):
        used = t[s][t[s][:, 0] > 0]
        ctiles = (N + 255) // 256
        assert len(used) % ctiles == 0 and len(used) >= ctiles * (Z // 64)     # every CTA of the grid reported
        for a, b in ((0, 1), (1, 2), (2, 4), (4, 5), (5, 6)):
            assert (used[:, a] <= used[:, b]).all(), (s, a, b)
        assert int(used[:, 10].sum()) == ctiles                                 # exactly one last arrival per column tile
        assert ((used[:, 7] > 0) == (used[:, 10] == 1)).all()                    # only the finaliser stamps [7]
    assert not t[2].any()                                                        # third slot untouched
    # the analysis of tools/rowchunk_timeline.py digests these stamps
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import rowchunk_timeline as RT
    r = RT.analyse(t[1])
    assert r["ctas"] == len(t[1][t[1][:, 0] > 0]) and r["span_us"] >= 0 and r["split_us_p50"]["finalise(last only)"] is not None


@pytest.mark.parametrize("gen", [1, 2, 3])
@pytest.mark.parametrize("rows", [32, 96, 256, 512])
def test_forced_k_chunk_sizes(gen, rows):
    """TEAL_TUNE_ROWS_*: any multiple of 32 input lanes per CTA (several lane groups per warp, ragged last chunk)."""
    H.set_tuning(generation=gen, rows_wide=rows, rows_narrow=rows)
    _check_gemv(1000, 776, 0.6745, seed=rows)          # narrow variant
    dec = H.EmuDecoder(CFG, _state_dict(CFG), THR)
    _run_and_compare(dec, True, lambda: dec.step(True))


@pytest.mark.parametrize("world", [2, 4])
def test_tp_peer_memory_allreduce_protocol(world):
    """The fused one-shot all-reduce + residual + sum-of-squares kernel (gpt-fast/tp.py:120-121,139-140 + model.py:160-161)
    with `world` emulated ranks running concurrently on shared host memory: every rank publishes its epoch into every
    peer's flag array, waits for the peers, sums the partials in rank order -- replicas stay bit-identical over many
    steps on the two alternating buffers."""
    import threading
    L = H.lib()
    dim, steps = 512, 6
    g = torch.Generator().manual_seed(world)
    parts = [[H.aligned((dim,), np.float16, 0) for _ in range(world)] for _ in range(2)]      # [parity][rank]
    flags = [H.aligned((world,), np.uint32, 0) for _ in range(world)]
    xs = [H.aligned((dim,), np.float16, 0) for _ in range(world)]
    x0 = (torch.randn(dim, generator=g)).half().numpy()
    for r in range(world):
        xs[r][...] = x0
    ssq = [H.aligned((64,), np.float32, 0) for _ in range(world)]
    epoch = [H.aligned((1,), np.uint32, 0) for _ in range(world)]
    vals = [[(torch.randn(dim, generator=g) * 0.5).half().numpy() for _ in range(world)] for _ in range(steps)]
    import ctypes
    errors = []

    def rank_main(r):
        try:
            for s in range(steps):
                par = s & 1
                parts[par][r][...] = vals[s][r]                       # "the row-parallel GEMV wrote its partial"
                pp = (ctypes.c_void_p * world)(*[parts[par][q].ctypes.data for q in range(world)])
                pf = (ctypes.c_void_p * world)(*[flags[q].ctypes.data for q in range(world)])
                H.check(L.teal_tp_allreduce_residual_f16(pp, pf, r, world, H.ptr(xs[r]), H.ptr(ssq[r]), H.ptr(epoch[r]), dim, None), "tp")
        except Exception as e:  # noqa: BLE001
            errors.append(e)

    threads = [threading.Thread(target=rank_main, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=120)
    assert not errors and not any(t.is_alive() for t in threads), errors
    want = torch.from_numpy(x0.copy())
    for s in range(steps):
        tot = torch.zeros(dim)
        for r in range(world):
            tot += torch.from_numpy(vals[s][r]).float()              # fp32 sum in rank order, one rounding
        want = (want + tot.half())                                   # fp16 residual add
    for r in range(world):
        assert np.array_equal(xs[r].view(np.uint16), xs[0].view(np.uint16))      # replicas bit-identical
        assert np.array_equal(xs[r].view(np.uint16), want.numpy().view(np.uint16))
        assert abs(float(ssq[r].sum()) - float((want.float() ** 2).sum())) <= 1e-3 * float((want.float() ** 2).sum())


@pytest.mark.parametrize("world,Z,N,cfg", [(2, 256, 512, 0), (4, 512, 776, 0), (2, 2048, 1024, 0), (2, 512, 1032, 31), (4, 1024, 512, 31)])
def test_row_parallel_gemv_fused_with_its_tp_exchange(world, Z, N, cfg):
    """teal_sparse_gemv_tp_exchange_f16 with `world` emulated ranks running concurrently: each rank gathers its Z/world
    lanes, the tile finalisers exchange partials over 'peer' memory -- x ends bit-identical on every rank and equals
    residual + fp16(sum of the ranks' fp16 partials) of the oracle; two alternating buffer sets over several steps."""
    import ctypes
    import threading
    L = H.lib()
    L.teal_emu_set_concurrent(1)          # every CTA alive: finalisers wait for peers while the rest of the grid runs
    H.set_config(cfg)                     # 0: the narrow default variant's TP twin, 31: the wide one's
    steps, Zs = 4, Z // world
    g = torch.Generator().manual_seed(world * 1000 + Z)
    W = (torch.randn(N, Z, generator=g) * 0.05).half()
    Wc = [H.column_major(W[:, r * Zs:(r + 1) * Zs].numpy()) for r in range(world)]           # row-parallel shards (tp.py:117)
    x0 = torch.randn(N, generator=g).half()
    hs = [torch.randn(1, 1, Z, generator=g).half() for _ in range(steps)]                     # replicated input of each step
    thr = 0.5
    parts = [[H.aligned((N,), np.float16, 0) for _ in range(world)] for _ in range(2)]
    flags = [[H.aligned((world * 64,), np.uint32, 0) for _ in range(world)] for _ in range(2)]
    epochs = [[H.aligned((64,), np.uint32, 0) for _ in range(world)] for _ in range(2)]
    xs = [H.aligned((N,), np.float16, x0.numpy()) for _ in range(world)]
    ssq = [H.aligned((64,), np.float32, 0) for _ in range(world)]
    errors = []

    def rank_main(r):
        try:
            ws = H.Workspace()
            for s in range(steps):
                par = s & 1
                xa = H.aligned((Zs,), np.float16, hs[s].reshape(-1)[r * Zs:(r + 1) * Zs].numpy())
                pp = (ctypes.c_void_p * world)(*[parts[par][q].ctypes.data for q in range(world)])
                pf = (ctypes.c_void_p * world)(*[flags[par][q].ctypes.data for q in range(world)])
                H.check(L.teal_sparse_gemv_tp_exchange_f16(H.ptr(xa), H.ptr(Wc[r]), H.ptr(xs[r]), H.ptr(ssq[r]), Zs, N, N, thr, pp, pf,
                                                           H.ptr(epochs[par][r]), r, world, H.ptr(ws.buf), ws.nbytes, None), "tp gemv")
                ws.assert_clean()
        except Exception as e:  # noqa: BLE001
            errors.append(e)

    threads = [threading.Thread(target=rank_main, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=300)
    assert not errors and not any(t.is_alive() for t in threads), errors
    want = x0.clone()
    for s in range(steps):
        tot = torch.zeros(N)
        for r in range(world):
            sl = slice(r * Zs, (r + 1) * Zs)
            tot += O.sparse_linear_f32(hs[s][..., sl], W[:, sl], thr).reshape(-1).half().float()   # the rank's fp16 partial
        want = want + tot.half()
    for r in range(world):
        assert np.array_equal(xs[r].view(np.uint16), xs[0].view(np.uint16)), r                     # replicas bit-identical
    got = torch.from_numpy(xs[0].copy()).float()
    assert ((got - want.float()).abs() <= want.float().abs() * 2.0 ** -8 + 8e-3).all(), float((got - want.float()).abs().max())
    assert abs(float(ssq[0].sum()) - float((got ** 2).sum())) <= 1e-3 * float((got ** 2).sum())


@pytest.mark.parametrize("opts", [1, 2, 3, 4, 5, 6, 7])
def test_generation1_one_change_at_a_time(opts):
    """TEAL_TUNE_V1_OPTS: early x / norm-weight requests, FHFMA, 32-bit row offsets on the validated kernel structure."""
    H.set_tuning(v1_opts=opts, rows_narrow=96 if opts & 1 else 0)     # (96: a warp owns a ragged second lane group)
    for Z, N in SHAPES[:4]:
        _check_gemv(Z, N, 0.6745, seed=opts)
    dec = H.EmuDecoder(CFG, _state_dict(CFG), THR)
    _run_and_compare(dec, True, lambda: dec.step(True))


def _chain_case(dim, inter, seed):
    g = torch.Generator().manual_seed(seed)
    resid = torch.randn(1, 1, dim, generator=g).half()
    nw = (1 + 0.1 * torch.randn(dim, generator=g)).half()
    W1, W3 = [(torch.randn(inter, dim, generator=g) * 0.05).half() for _ in range(2)]
    W2 = (torch.randn(dim, inter, generator=g) * 0.05).half()
    return resid, nw, W1, W3, W2


@pytest.mark.parametrize("dim,inter,wide_kelems,concurrent", [(256, 704, 0, False), (256, 704, 0, True), (256, 1000, 0, True),
                                                               (512, 1536, 100, True), (512, 1400, 100, False), (256, 704, 300, True)])
def test_flag_chained_mlp_pair(dim, inter, wide_kelems, concurrent):
    """teal_norm_sparse_gate_up_chain_f16 -> teal_sparse_gemv_residual_chain_f16: the down projection waits for h tile by
    tile instead of for the whole gate/up grid.  `concurrent`: the two launches really overlap (two host threads; the
    consumer is started FIRST and spins until the producer's finalisers release its tiles).  wide_kelems lowers the
    wide-variant threshold so that the 512-column twins and mixed tile widths are exercised at emulation sizes.
    Result == the unchained pair's (oracle tolerance), ready cells back to zero, workspace clean."""
    import threading
    L = H.lib()
    L.teal_emu_set_concurrent(1)
    H.set_tuning(wide_kelems=wide_kelems)
    resid, nw, W1, W3, W2 = _chain_case(dim, inter, dim + inter)
    tg, tu, td, eps = 0.4, 0.6, 0.02, 1e-5
    ws = H.Workspace(1 << 17)                             # the product's size (teal_b200/ops.py): room for the second accumulator region
    xa = H.aligned((dim,), np.float16, resid.reshape(-1).numpy())
    nwa = H.aligned((dim,), np.float16, nw.numpy())
    W1c, W3c, W2c = H.column_major(W1.numpy()), H.column_major(W3.numpy()), H.column_major(W2.numpy())
    ssq = H.aligned((64,), np.float32, 0)
    ssq[0] = float((resid.float() ** 2).sum())
    for rep in range(2):                                  # twice on the same workspace: the cells reset themselves
        h = H.aligned((inter,), np.float16, 0)
        y = H.aligned((dim,), np.float16, 0)
        ssq_out = H.aligned((64,), np.float32, 5.0)
        errors = []

        def produce():
            try:
                H.check(L.teal_norm_sparse_gate_up_chain_f16(H.ptr(xa), H.ptr(nwa), eps, H.ptr(ssq), H.ptr(W1c), H.ptr(W3c), H.ptr(h), dim, inter,
                                                             inter, inter, tg, tu, 0, H.ptr(ws.buf), ws.nbytes, None), "gate_up_chain")
            except Exception as e:  # noqa: BLE001
                errors.append(e)

        def consume():
            try:
                H.check(L.teal_sparse_gemv_residual_chain_f16(H.ptr(h), H.ptr(W2c), H.ptr(xa), H.ptr(y), H.ptr(ssq_out), inter, dim, dim, td, 0,
                                                              H.ptr(ws.buf), ws.nbytes, None), "down_chain")
            except Exception as e:  # noqa: BLE001
                errors.append(e)

        if concurrent:
            tc, tp = threading.Thread(target=consume), threading.Thread(target=produce)
            tc.start()
            import time
            time.sleep(0.05)                              # the consumer grid is alive and polling before any h tile exists
            tp.start()
            tp.join(timeout=120)
            tc.join(timeout=120)
            assert not tc.is_alive() and not tp.is_alive()
        else:
            produce()
            consume()
        assert not errors, errors
        ws.assert_clean()                                 # counters, both accumulator regions and the ready cells are zero again
        xn = O.rms_norm(resid, nw, eps)
        g32, u32 = O.sparse_linear_f32(xn, W1, tg), O.sparse_linear_f32(xn, W3, tu)
        h_ref = (torch.nn.functional.silu(g32.half()) * u32.half()).reshape(-1)
        h_t = torch.from_numpy(h.copy())
        assert ((h_t.float() - h_ref.float()).abs() <= h_ref.float().abs() * 2.0 ** -8 + 2e-3).all()
        d32 = O.sparse_linear_f32(h_t.reshape(1, 1, -1), W2, td).reshape(-1)
        want = (resid.reshape(-1) + d32.half()).float()
        y_t = torch.from_numpy(y.copy()).float()
        assert ((y_t - want).abs() <= want.abs() * 2.0 ** -9 + 2e-3).all(), (rep, float((y_t - want).abs().max()))
        assert abs(float(ssq_out.sum()) - float((y_t ** 2).sum())) <= 1e-4 * float((y_t ** 2).sum())


@pytest.mark.parametrize("wide_kelems", [0, 150])
def test_decode_step_with_the_flag_chained_mlp_pair(wide_kelems):
    H.set_tuning(wide_kelems=wide_kelems)
    dec = H.EmuDecoder(CFG, _state_dict(CFG), THR)
    dec.chain_mlp = True
    _run_and_compare(dec, True, lambda: dec.step(True))


@pytest.mark.parametrize("uniform_qkv,concurrent,wide_kelems", [(True, False, 0), (False, False, 0), (True, True, 0), (False, True, 0),
                                                                (False, True, 60), (True, True, 200)])
def test_decode_step_with_flag_chained_blocks(uniform_qkv, concurrent, wide_kelems):
    """Both blocks of every layer flag-chained: QKV -> attention -> o-proj (tiles / heads) and gate/up -> down (h tiles).
    `concurrent`: the launches of a block run in concurrent host threads, last launch first, so every consumer grid is
    polling before its producer starts.  3-threshold QKV (tiles per q | k | v region) and uniform QKV (one region, slices
    share tiles); wide_kelems moves the narrow / wide variant boundary so that 512-column tiles hold several heads."""
    H.set_tuning(wide_kelems=wide_kelems)
    H.lib().teal_emu_set_concurrent(1 if concurrent else 0)
    thr = [dict(t, k=t["q"], v=t["q"]) for t in THR] if uniform_qkv else THR
    dec = H.EmuDecoder(CFG, _state_dict(CFG), thr)
    dec.chain_attn = dec.chain_mlp = True
    dec.concurrent_chain = concurrent
    if uniform_qkv:
        orc = O.OracleDecoder(CFG, _state_dict(CFG), thr, max_seq=64)
        dec.reset(TOKS[0])
        for i, t in enumerate(TOKS):
            dec.token[0] = t
            dec.step(True)
            with torch.no_grad():
                want = orc.forward(torch.tensor([[t]]), torch.tensor([i])).float().reshape(-1)
            got = torch.from_numpy(dec.logits.astype(np.float32))
            assert float((got - want).abs().max()) <= 6e-2 * float(want.abs().max()) + 2e-2, i
    else:
        _run_and_compare(dec, True, lambda: dec.step(True))


@pytest.mark.parametrize("uniform_qkv,concurrent,wide_kelems", [(False, False, 0), (True, True, 0), (False, True, 0), (False, True, 100)])
def test_decode_step_fully_flag_chained(uniform_qkv, concurrent, wide_kelems):
    """chain_norms on top of the two chained blocks: the RMSNorm points are hand-off cells too (o-proj's last finaliser releases
    the residual stream to the gate/up CTAs, down's to the next layer's QKV CTAs), so inside the layer stack no launch waits for
    a whole grid.  `concurrent`: the five launches of a layer run in concurrent host threads, the last one started first."""
    H.set_tuning(wide_kelems=wide_kelems)
    H.lib().teal_emu_set_concurrent(1 if concurrent else 0)
    thr = [dict(t, k=t["q"], v=t["q"]) for t in THR] if uniform_qkv else THR
    dec = H.EmuDecoder(CFG, _state_dict(CFG), thr)
    dec.chain_attn = dec.chain_mlp = dec.chain_norms = True
    dec.concurrent_chain = concurrent
    orc = O.OracleDecoder(CFG, _state_dict(CFG), thr, max_seq=64)
    dec.reset(TOKS[0])
    for i, t in enumerate(TOKS):
        dec.token[0] = t
        dec.step(True)                     # (asserts that the workspace -- counters, accumulators, every hand-off cell -- is zero again)
        with torch.no_grad():
            want = orc.forward(torch.tensor([[t]]), torch.tensor([i])).float().reshape(-1)
        got = torch.from_numpy(dec.logits.astype(np.float32))
        assert float((got - want).abs().max()) <= 6e-2 * float(want.abs().max()) + 2e-2, i


def test_timeline_of_a_chained_step():
    """The chained twins stamp through the same hook (decided at run time): slot i = i-th row-chunk launch of the step."""
    L = H.lib()
    dec = H.EmuDecoder(CFG, _state_dict(CFG), THR)
    dec.chain_attn = dec.chain_mlp = dec.chain_norms = True
    n = 4 * CFG["n_layer"]
    buf = H.aligned((n * 4096 * 16,), np.int64, 0)
    H.check(L.teal_set_debug_timeline(H.ptr(buf), n), "timeline")
    try:
        dec.reset(3)
        dec.token[0] = 3
        dec.step(True)
    finally:
        H.check(L.teal_set_debug_timeline(None, 0), "timeline off")
    t = buf.reshape(n, 4096, 16)
    for s in range(n):
        used = t[s][t[s][:, 0] > 0]
        assert len(used) > 0 and (used[:, 0] <= used[:, 1]).all() and (used[:, 1] <= used[:, 4]).all(), s
        assert int(used[:, 10].sum()) >= 1


@pytest.mark.parametrize("Z,N", [(4096, 11008), (11008, 4096), (4096, 4096), (4096, 6144)])
def test_baseline_shapes_at_full_size(Z, N):
    """BASELINE.json config 1 (Llama-2-7B gate_proj 4096 -> 11008, batch 1, 50 %) and its sibling projections at their real
    sizes through the default dispatch on the emulation (the wide- and narrow-tile production kernels, multi-wave grids)."""
    H.lib().teal_emu_set_sm_count(8)
    _check_gemv(Z, N, 0.6745, seed=Z + N)


def test_llama3_8b_mlp_block_at_full_size_chained_and_unchained():
    """One Llama-3-8B MLP block (4096 -> 2 x 14336 -> 4096) at full size: fused RMSNorm + gate/up + SiLU*mul -> down +
    residual, unchained and flag-chained with the two launches overlapping (consumer first); both against the oracle."""
    import threading
    import time
    L = H.lib()
    L.teal_emu_set_sm_count(8)
    dim, inter, eps, tg, tu, td = 4096, 14336, 1e-5, 0.6745, 0.6745, 0.01
    resid, nw, W1, W3, W2 = _chain_case(dim, inter, 7)
    W1, W3, W2 = (W1 * 0.4).half(), (W3 * 0.4).half(), (W2 * 0.4).half()
    ws = H.Workspace(1 << 17)
    xa0 = resid.reshape(-1).numpy()
    nwa = H.aligned((dim,), np.float16, nw.numpy())
    W1c, W3c, W2c = H.column_major(W1.numpy()), H.column_major(W3.numpy()), H.column_major(W2.numpy())
    ssq = H.aligned((64,), np.float32, 0)
    ssq[0] = float((resid.float() ** 2).sum())
    xn = O.rms_norm(resid, nw, eps)
    h_ref = (torch.nn.functional.silu(O.sparse_linear_f32(xn, W1, tg).half()) * O.sparse_linear_f32(xn, W3, tu).half()).reshape(-1)
    for chained in (False, True):
        xa = H.aligned((dim,), np.float16, xa0)
        h, y, ssq_out = H.aligned((inter,), np.float16, 0), H.aligned((dim,), np.float16, 0), H.aligned((64,), np.float32, 0)
        if chained:
            calls = [lambda: H.check(L.teal_norm_sparse_gate_up_chain_f16(H.ptr(xa), H.ptr(nwa), eps, H.ptr(ssq), H.ptr(W1c), H.ptr(W3c), H.ptr(h),
                                                                          dim, inter, inter, inter, tg, tu, 0, H.ptr(ws.buf), ws.nbytes, None), "p"),
                     lambda: H.check(L.teal_sparse_gemv_residual_chain_f16(H.ptr(h), H.ptr(W2c), H.ptr(xa), H.ptr(y), H.ptr(ssq_out), inter, dim,
                                                                           dim, td, 0, H.ptr(ws.buf), ws.nbytes, None), "c")]
            errs = []

            def guarded(c):
                try:
                    c()
                except Exception as e:  # noqa: BLE001
                    errs.append(e)
            tc, tp = threading.Thread(target=guarded, args=(calls[1],)), threading.Thread(target=guarded, args=(calls[0],))
            tc.start()
            time.sleep(0.1)
            tp.start()
            tp.join(timeout=300)
            tc.join(timeout=300)
            assert not errs and not tc.is_alive()
        else:
            H.check(L.teal_norm_sparse_gate_up_f16(H.ptr(xa), H.ptr(nwa), eps, H.ptr(ssq), H.ptr(W1c), H.ptr(W3c), H.ptr(h), None, dim, inter,
                                                   inter, inter, tg, tu, H.ptr(ws.buf), ws.nbytes, None), "gate_up")
            H.check(L.teal_sparse_gemv_residual_f16(H.ptr(h), H.ptr(W2c), H.ptr(xa), H.ptr(y), H.ptr(ssq_out), inter, dim, dim, td,
                                                    H.ptr(ws.buf), ws.nbytes, None), "down")
        ws.assert_clean()
        h_t = torch.from_numpy(h.copy())
        assert ((h_t.float() - h_ref.float()).abs() <= h_ref.float().abs() * 2.0 ** -8 + 2e-3).all(), chained
        want = (resid.reshape(-1) + O.sparse_linear_f32(h_t.reshape(1, 1, -1), W2, td).reshape(-1).half()).float()
        y_t = torch.from_numpy(y.copy()).float()
        assert ((y_t - want).abs() <= want.abs() * 2.0 ** -9 + 2e-3).all(), chained
