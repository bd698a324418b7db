"""The CUDA kernels' LOGIC on the CPU: the production .cu sources compiled as C++ against tests/emu/cuda_emu.h (fibers
for threads, real atomics, warp collectives through per-warp slots) and driven through the same C ABI, checked against
the oracle.  The CPU suite exercises the kernel logic (indexing, masks, lists, the arrival / finaliser protocol, the TP
exchange with concurrent emulated ranks) through the same entry points the GPU tests use.  It proves indexing / masks / pipelines / protocols, not the memory model
and not speed.  The -m gpu tests remain the parity tests proper."""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import teal_oracle as O

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu"))
import harness as H  # noqa: E402

VARIANTS = [0, 1, 2, 3, 4]          # kernel variants of csrc/gemv.cu (tile width x warps per CTA)
LANES = [0, 32, 96, 256, 512]       # input lanes per warp (0 = chosen by launch size)


@pytest.fixture(autouse=True)
def _defaults():
    H.lib()
    H.set_config(0)
    H.lib().teal_emu_set_sm_count(4)
    H.lib().teal_emu_set_concurrent(0)
    H.lib().teal_emu_set_max_workers(0)
    yield
    H.set_config(0)
    H.lib().teal_emu_set_max_workers(0)


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


@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("lanes", LANES)
def test_every_variant_and_k_chunk_matches_oracle(variant, lanes):
    """All kernel variants x forced lanes per warp (several 32-lane groups per warp, more than one request pass, ragged
    last chunk, partial last column tile) against the oracle."""
    H.set_variant(variant, lanes)
    for i, (Z, N) in enumerate(SHAPES):
        _check_gemv(Z, N, 0.6745, seed=i)


@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("thr", [0.0, 100.0, 0.05, 2.0])
def test_sparsity_extremes(variant, thr):
    H.set_variant(variant, 64)
    _check_gemv(1000, 776, thr, seed=3)


def test_default_dispatch_and_survivors_bit_exact():
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
        assert int(cnt[0]) == len(want) and np.array_equal(idx[: len(want)], want), (Z, thr)
    _check_gemv(512, 1024, 0.6745)


def _tie_vector():
    # |x| == fp16(t) with fp16(t) > t: the oracle (fp16 compare) drops it
    t = 0.02038193
    t16 = float(torch.tensor(t, dtype=torch.float16))
    x = torch.tensor([t16, -t16, t16 * 1.001, 0.0, float("nan"), -0.0, 1.0, -1.0], dtype=torch.float16).reshape(1, 1, -1)
    return x, t


def test_tie_rule_and_nan():
    x, t = _tie_vector()
    Z = x.numel()
    xa = H.aligned((Z,), np.float16, x.reshape(-1).numpy())
    idx = H.aligned((Z,), np.int32, 0)
    cnt = H.aligned((1,), np.int32, 0)
    H.check(H.lib().teal_survivors_f16(H.ptr(xa), Z, t, H.ptr(idx), H.ptr(cnt), None), "survivors")
    got = idx[: int(cnt[0])].tolist()
    assert got == [2, 4, 6, 7]          # the two ties and the zeros are dropped, NaN is gathered (poisons y like mask*x)


@pytest.mark.parametrize("uniform", [True, False])
def test_qkv_region_thresholds(uniform):
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


@pytest.mark.parametrize("variant,lanes", [(-1, 0), (2, 32), (1, 96)])
def test_fused_layer_half_steps(variant, lanes):
    """RMSNorm+gate/up+SiLU*mul -> down+residual (+ sum of squares) -> RMSNorm+QKV from the partials: the decode step's
    GEMV chain (gpt-fast/model.py:159-161,258-259,286-291) on one small layer."""
    H.set_variant(variant, lanes)
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
    ssq = H.aligned((64 * 32,), np.float32, 7.0)                                            # stale garbage must be overwritten
    H.check(L.teal_sparse_gemv_residual_f16(H.ptr(h), H.ptr(W2c), H.ptr(xa), H.ptr(y), H.ptr(ssq), inter, dim, dim, td,
                                            H.ptr(ws.buf), ws.nbytes, None), "down_residual")
    ws.assert_clean()
    d32 = O.sparse_linear_f32(h_t.reshape(1, 1, -1), W2, td).reshape(-1)
    want = (resid.reshape(-1) + d32.half()).float()
    y_t = torch.from_numpy(y.copy())
    assert ((y_t.float() - want).abs() <= want.abs() * 2.0 ** -9 + 2e-3).all()
    ssq_r = ssq.reshape(32, 64)
    assert all(np.array_equal(ssq_r[r], ssq_r[0]) for r in range(32))                  # every replica fully rewritten, identical
    assert abs(float(ssq_r[0].sum()) - float((y_t.float() ** 2).sum())) <= 1e-4 * float((y_t.float() ** 2).sum())
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


@pytest.mark.parametrize("variant,lanes,sparse", [(-1, 0, False), (-1, 0, True), (2, 32, True), (1, 96, True), (4, 64, True)])
def test_decode_step_launch_per_projection(variant, lanes, sparse):
    """SparseDecoder._enqueue_step's call sequence (embed | 5 launches per layer | lm_head + arg-max | advance) on the
    emulation against OracleDecoder (pinned to the reference's gpt-fast Transformer by tests/golden/gptfast_dense.npz)."""
    H.set_variant(variant, lanes)
    dec = H.EmuDecoder(CFG, _state_dict(CFG), THR)
    _run_and_compare(dec, sparse, lambda: dec.step(sparse))


@pytest.mark.parametrize("variant,lanes,sparse", [(-1, 0, True), (-1, 0, False), (2, 32, True), (4, 64, True)])
def test_deferred_epilogue_is_bit_identical_to_the_finaliser_route(variant, lanes, sparse):
    """The decoder's default route (QKV and gate/up leave fp32 sums, attention and the down projection round them on load,
    the o projection clears the buffer) against the route where every launch finalises its own tiles: same thresholds
    decisions, same fp16 roundings => the same logits, KV cache and tokens, bit for bit (fp32 split-K sums depend on the
    order of the REDs, so the emulation runs the CTAs one after the other here: same order on both routes)."""
    H.set_variant(variant, lanes)
    H.lib().teal_emu_set_max_workers(1)
    a, b = H.EmuDecoder(CFG, _state_dict(CFG), THR), H.EmuDecoder(CFG, _state_dict(CFG), THR)
    b.deferred = False
    a.reset(TOKS[0]); b.reset(TOKS[0])
    for t in TOKS:
        a.token[0] = b.token[0] = t
        a.step(sparse); b.step(sparse)
        assert np.array_equal(a.logits.view(np.uint16), b.logits.view(np.uint16))
        assert np.array_equal(a.x.view(np.uint16), b.x.view(np.uint16)) and int(a.token[0]) == int(b.token[0])
    for Wa, Wb in zip(a.layers, b.layers):
        assert np.array_equal(Wa["k_cache"].view(np.uint16), Wb["k_cache"].view(np.uint16))
        assert np.array_equal(Wa["v_cache"].view(np.uint16), Wb["v_cache"].view(np.uint16))
    assert not a.defer_acc[: a.nqkv].any()              # q|k|v sums cleared; gate/up sums stay until the next o projection


def test_deferred_pieces_against_their_finalised_twins():
    """Each deferred entry point against the fused one it replaces, on ragged shapes: the fp32 sums rounded once are the
    fp16 outputs; silu(gate) * up computed by the consumer equals the pair finaliser's h; bad arguments are refused."""
    L = H.lib()
    L.teal_emu_set_max_workers(1)                                   # fixed RED order: the comparisons below are bit-exact
    dim, inter, N, kv, eps = 520, 1112, 776, 128, 1e-5
    g = torch.Generator().manual_seed(23)
    x = H.aligned((dim,), np.float16, torch.randn(dim, generator=g).half().numpy())
    nw = H.aligned((dim,), np.float16, (1 + 0.1 * torch.randn(dim, generator=g)).half().numpy())
    Wq = H.column_major((torch.randn(N, dim, generator=g) * 0.05).half().numpy())
    W1, W3 = [H.column_major((torch.randn(inter, dim, generator=g) * 0.05).half().numpy()) for _ in range(2)]
    W2 = H.column_major((torch.randn(dim, inter, generator=g) * 0.05).half().numpy())
    ws = H.Workspace()
    ssq = H.aligned((64 * 32,), np.float32, 0)
    H.check(L.teal_norm_sparse_qkv_gemv_f16(H.ptr(x), H.ptr(nw), eps, None, H.ptr(Wq), H.ptr(H.aligned((N,), np.float16, 0)), None, dim, N, N,
                                            kv, 0.5, 0.5, 0.5, H.ptr(ws.buf), ws.nbytes, None), "ssq via scratch")
    ssq[...] = ws.buf[ws.nbytes - 64 * 32 * 4:].view(np.float32)    # the partials the library computed for x
    for tq, tk, tv in ((0.5, 0.5, 0.5), (0.6, 0.3, 0.9)):
        q16, qacc = H.aligned((N,), np.float16, 0), H.aligned((N,), np.float32, 0)
        H.check(L.teal_norm_sparse_qkv_gemv_f16(H.ptr(x), H.ptr(nw), eps, H.ptr(ssq), H.ptr(Wq), H.ptr(q16), None, dim, N, N, kv, tq, tk, tv,
                                                H.ptr(ws.buf), ws.nbytes, None), "qkv")
        H.check(L.teal_norm_sparse_qkv_gemv_acc_f16(H.ptr(x), H.ptr(nw), eps, H.ptr(ssq), H.ptr(Wq), H.ptr(qacc), dim, N, N, kv, tq, tk, tv, None), "qkv acc")
        assert np.array_equal(qacc.astype(np.float16).view(np.uint16), q16.view(np.uint16))
    h16, gu = H.aligned((inter,), np.float16, 0), H.aligned((2 * inter,), np.float32, 0)
    H.check(L.teal_norm_sparse_gate_up_f16(H.ptr(x), H.ptr(nw), eps, H.ptr(ssq), H.ptr(W1), H.ptr(W3), H.ptr(h16), None, dim, inter, inter, inter,
                                           0.4, 0.6, H.ptr(ws.buf), ws.nbytes, None), "gate_up")
    H.check(L.teal_norm_sparse_gate_up_acc_f16(H.ptr(x), H.ptr(nw), eps, H.ptr(ssq), H.ptr(W1), H.ptr(W3), H.ptr(gu), dim, inter, inter, inter,
                                               0.4, 0.6, None), "gate_up acc")
    ws.assert_clean()
    resid = H.aligned((dim,), np.float16, torch.randn(dim, generator=g).half().numpy())
    ya, yb = H.aligned((dim,), np.float16, 0), H.aligned((dim,), np.float16, 0)
    sa, sb = H.aligned((64 * 32,), np.float32, 3.0), H.aligned((64 * 32,), np.float32, 5.0)
    junk = H.aligned((1000,), np.float32, 9.0)
    H.check(L.teal_sparse_gemv_residual_f16(H.ptr(h16), H.ptr(W2), H.ptr(resid), H.ptr(ya), H.ptr(sa), inter, dim, dim, 0.02,
                                            H.ptr(ws.buf), ws.nbytes, None), "down")
    H.check(L.teal_sparse_gemv_residual_ex_f16(None, H.ptr(gu), H.ptr(W2), H.ptr(resid), H.ptr(yb), H.ptr(sb), inter, dim, dim, 0.02,
                                               H.ptr(junk), 1000, H.ptr(ws.buf), ws.nbytes, None), "down from gate/up sums")
    ws.assert_clean()
    assert np.array_equal(ya.view(np.uint16), yb.view(np.uint16)) and np.array_equal(sa, sb)
    assert not junk.any()                                           # the clear range is zeroed, all of it
    # refusals: both / neither input, misaligned or ragged clear range, deferred launch without its buffer
    bad = [L.teal_sparse_gemv_residual_ex_f16(H.ptr(h16), H.ptr(gu), H.ptr(W2), H.ptr(resid), H.ptr(yb), None, inter, dim, dim, 0.02, None, 0,
                                              H.ptr(ws.buf), ws.nbytes, None),
           L.teal_sparse_gemv_residual_ex_f16(None, None, H.ptr(W2), H.ptr(resid), H.ptr(yb), None, inter, dim, dim, 0.02, None, 0,
                                              H.ptr(ws.buf), ws.nbytes, None),
           L.teal_sparse_gemv_residual_ex_f16(H.ptr(h16), None, H.ptr(W2), H.ptr(resid), H.ptr(yb), None, inter, dim, dim, 0.02, H.ptr(junk), 999,
                                              H.ptr(ws.buf), ws.nbytes, None),
           L.teal_sparse_gemv_residual_ex_f16(H.ptr(h16), None, H.ptr(W2), H.ptr(resid), H.ptr(yb), None, inter, dim, dim, 0.02, H.ptr(junk[1:]), 996,
                                              H.ptr(ws.buf), ws.nbytes, None),
           L.teal_sparse_gemv_residual_ex_f16(None, H.ptr(gu), H.ptr(W2), H.ptr(resid), H.ptr(yb), None, inter, dim, dim, 0.02, H.ptr(gu), 2 * inter,
                                              H.ptr(ws.buf), ws.nbytes, None),      # a launch may not clear its own input
           L.teal_norm_sparse_qkv_gemv_acc_f16(H.ptr(x), H.ptr(nw), eps, H.ptr(ssq), H.ptr(Wq), None, dim, N, N, kv, 0.5, 0.5, 0.5, None),
           L.teal_norm_sparse_gate_up_acc_f16(H.ptr(x), H.ptr(nw), eps, None, H.ptr(W1), H.ptr(W3), H.ptr(gu), dim, inter, inter, inter, 0.4, 0.6, None),
           L.teal_attn_decode_acc_f16(None, None, None, None, None, None, 4, 2, 64, 64, None)]
    assert all(rc != 0 for rc in bad), bad
    ws.assert_clean()


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
    """teal_set_debug_timeline: the TIMED twin of the kernel write monotone stamps for every CTA into one
    slot per launch; results are those of the production instantiation."""
    L = H.lib()
    slots, slot_ints = 4, 4096 * 16
    buf = H.aligned((slots * slot_ints,), np.int64, 0)
    x, W = _mk(512, 1024, 9)
    ws = H.Workspace()
    y_plain = _gemv(x, W, 0.6745, ws)
    H.check(L.teal_set_debug_timeline(H.ptr(buf), slots), "timeline")
    try:
        y_timed = _gemv(x, W, 0.6745, ws)          # slot 0
        _check_gemv(1000, 776, 0.5, seed=4)         # slot 1
        # slot 2: a deferred launch (no arrival, no finaliser)
        g = torch.Generator().manual_seed(2)
        xd = H.aligned((512,), np.float16, torch.randn(512, generator=g).half().numpy())
        nwd = H.aligned((512,), np.float16, 1.0)
        ssqd = H.aligned((64 * 32,), np.float32, 0)
        ssqd.reshape(32, 64)[:, 0] = float((xd.astype(np.float32) ** 2).sum())
        accd = H.aligned((1024,), np.float32, 0)
        H.check(L.teal_norm_sparse_qkv_gemv_acc_f16(H.ptr(xd), H.ptr(nwd), 1e-5, H.ptr(ssqd), H.ptr(H.column_major(W.numpy())), H.ptr(accd),
                                                    512, 1024, 1024, 128, 0.5, 0.5, 0.5, None), "deferred qkv")
    finally:
        H.check(L.teal_set_debug_timeline(None, 0), "timeline off")
    assert torch.equal(y_plain, y_timed) or float((y_plain.float() - y_timed.float()).abs().max()) < 2e-3
    t = buf.reshape(slots, 4096, 16)
    for s, (Z, N) in enumerate([(512, 1024), (1000, 776)]):
        used = t[s][t[s][:, 0] > 0]
        ctiles = (N + 255) // 256
        assert len(used) % ctiles == 0 and len(used) >= ctiles                  # every CTA of the grid reported
        for a, b in ((0, 1), (1, 2), (2, 4), (4, 5), (5, 6)):
            assert (used[:, a] <= used[:, b]).all(), (s, a, b)
        assert int(used[:, 10].sum()) == ctiles                                 # exactly one last arrival per column tile
        assert ((used[:, 7] > 0) == (used[:, 10] == 1)).all()                    # only the finaliser stamps [7]
    used = t[2][t[2][:, 0] > 0]
    assert len(used) >= 4 and (used[:, 4] <= used[:, 5]).all() and (used[:, 6] == used[:, 5]).all()
    assert not used[:, 7].any() and not used[:, 10].any() and accd.any()         # deferred: nobody finalises
    assert not t[3].any()                                                        # fourth slot untouched
    # the analysis of tools/gemv_timeline.py digests these stamps
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import gemv_timeline as RT
    r = RT.analyse(t[1])
    assert r["ctas"] == len(t[1][t[1][:, 0] > 0]) and r["span_us"] >= 0 and r["split_us_p50"]["finalise(last only)"] is not None
    r = RT.analyse(t[2])
    assert r["split_us_p50"]["finalise(last only)"] is None and r["split_us_p50"]["fence_arrive"] == 0


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
    ssq = [H.aligned((64 * 32,), np.float32, 0) for _ in range(world)]
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
        assert abs(float(ssq[r][:64].sum()) - float((want.float() ** 2).sum())) <= 1e-3 * float((want.float() ** 2).sum())


@pytest.mark.parametrize("world,Z,N,cfg,from_sums", [(2, 256, 512, 0, False), (4, 512, 776, 0, False), (2, 2048, 1024, 0, False),
                                                     (2, 512, 1032, 3 + 16, False), (4, 1024, 512, 2 + 32, False),
                                                     (2, 512, 776, 0, True)])
def test_row_parallel_gemv_fused_with_its_tp_exchange(world, Z, N, cfg, from_sums):
    """teal_sparse_gemv_tp_exchange_f16 with `world` emulated ranks running concurrently: each rank gathers its Z/world
    lanes, the tile finalisers push their partials as {data, epoch} words into every rank's receive buffer -- x ends bit-identical on every rank and equals
    residual + fp16(sum of the ranks' fp16 partials) of the oracle; two alternating buffer sets over several steps."""
    import ctypes
    import threading
    L = H.lib()
    L.teal_emu_set_concurrent(1)          # every CTA alive: finalisers wait for peers while the rest of the grid runs
    H.set_config(cfg)                     # 0: the default variant's TP twin; the 512-column and the 4-warp twins with forced K chunks
    steps, Zs = 4, Z // world
    g = torch.Generator().manual_seed(world * 1000 + Z)
    W = (torch.randn(N, Z, generator=g) * 0.05).half()
    Wc = [H.column_major(W[:, r * Zs:(r + 1) * Zs].numpy()) for r in range(world)]           # row-parallel shards (tp.py:117)
    x0 = torch.randn(N, generator=g).half()
    hs = [torch.randn(1, 1, Z, generator=g).half() for _ in range(steps)]                     # replicated input of each step
    gus = None
    if from_sums:   # the down projection behind a deferred gate/up launch: the input arrives as fp32 (gate | up) sums, the
        #             launch also clears a deferred buffer (what the o projection does in the decode step)
        gus = [(torch.randn(Z, generator=g), torch.randn(Z, generator=g)) for _ in range(steps)]
        hs = [(torch.nn.functional.silu(gg.half()) * uu.half()).reshape(1, 1, Z) for gg, uu in gus]
    junk = [H.aligned((256,), np.float32, 1.0) for _ in range(world)]
    thr = 0.5
    recv = [[H.aligned((world * N // 2, 2), np.uint32, 0) for _ in range(world)] for _ in range(2)]   # [parity][rank]: {2 x fp16, epoch} words
    epochs = [[H.aligned((64,), np.uint32, 0) for _ in range(world)] for _ in range(2)]
    xs = [H.aligned((N,), np.float16, x0.numpy()) for _ in range(world)]
    ssq = [H.aligned((64 * 32,), np.float32, 0) for _ in range(world)]
    errors = []

    def rank_main(r):
        try:
            ws = H.Workspace()
            for s in range(steps):
                par = s & 1
                xa = H.aligned((Zs,), np.float16, hs[s].reshape(-1)[r * Zs:(r + 1) * Zs].numpy())
                pp = (ctypes.c_void_p * world)(*[recv[par][q].ctypes.data for q in range(world)])
                gu = None
                if from_sums:
                    gu = H.aligned((2 * Zs,), np.float32, np.concatenate([gus[s][0][r * Zs:(r + 1) * Zs].numpy(), gus[s][1][r * Zs:(r + 1) * Zs].numpy()]))
                    junk[r][...] = 1.0
                H.check(L.teal_sparse_gemv_tp_exchange_f16(None if from_sums else H.ptr(xa), H.ptr(gu), H.ptr(Wc[r]), H.ptr(xs[r]), H.ptr(ssq[r]),
                                                           Zs, N, N, thr, H.ptr(junk[r]) if from_sums else None, 256 if from_sums else 0,
                                                           pp, H.ptr(epochs[par][r]), r, world, H.ptr(ws.buf), ws.nbytes, None), "tp gemv")
                assert not from_sums or not junk[r].any()
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
    assert abs(float(ssq[0][:64].sum()) - float((got ** 2).sum())) <= 1e-3 * float((got ** 2).sum())
    assert np.array_equal(ssq[0][:64], ssq[0][64 * 31:])


@pytest.mark.parametrize("Z,N", [(4096, 11008), (11008, 4096), (4096, 4096), (4096, 6144)])
def test_baseline_shapes_at_full_size(Z, N):
    """BASELINE.json config 1 (Llama-2-7B gate_proj 4096 -> 11008, batch 1, 50 %) and its sibling projections at their real
    sizes through the default dispatch on the emulation (single-wave K chunks chosen by the launcher)."""
    H.lib().teal_emu_set_sm_count(8)
    _check_gemv(Z, N, 0.6745, seed=Z + N)


@pytest.mark.parametrize("hd,n_head,n_kv", [(128, 4, 1), (64, 4, 2)])
def test_attention_over_several_sweeps_of_the_cache(hd, n_head, n_kv):
    """attn_decode_kernel at positions on both sides of every sweep boundary (a sweep = 16 warps x 8 positions = 128 cache
    rows; the first sweep's rows, the position and its RoPE row are requested before the PDL wait, later sweeps inside
    the loop): against RoPE + SDPA of the oracle over a cache of random rows, for both input forms (fp16 q|k|v, and the
    fp32 sums of a deferred QKV launch), with the appended cache row bit-exact."""
    L = H.lib()
    S = 320
    g = torch.Generator().manual_seed(hd)
    freqs_t = O.precompute_freqs_cis(S, hd, 10000)
    freqs = H.aligned(tuple(freqs_t.shape), np.float16, freqs_t.numpy())
    nq = (n_head + 2 * n_kv) * hd
    for P in (0, 1, 15, 16, 17, 127, 128, 129, 200, 255, 256, 257, 319):
        kc_t, vc_t = torch.randn(n_kv, S, hd, generator=g).half(), torch.randn(n_kv, S, hd, generator=g).half()
        kc_t[:, P:], vc_t[:, P:] = 0, 0                                     # rows at and beyond the position are not cache yet
        qkv_t = torch.randn(nq, generator=g).half()
        q, k, v = qkv_t.split([n_head * hd, n_kv * hd, n_kv * hd])
        fc = freqs_t[P:P + 1]
        qr = O.apply_rotary_emb(q.view(1, 1, n_head, hd), fc).transpose(1, 2)
        kr = O.apply_rotary_emb(k.view(1, 1, n_kv, hd), fc).transpose(1, 2)
        kref, vref = kc_t.clone(), vc_t.clone()
        kref[:, P], vref[:, P] = kr[0, :, 0], v.view(n_kv, hd)
        kk = kref[None].repeat_interleave(n_head // n_kv, dim=1)[:, :, : P + 1].float()
        vv = vref[None].repeat_interleave(n_head // n_kv, dim=1)[:, :, : P + 1].float()
        want = torch.nn.functional.scaled_dot_product_attention(qr.float(), kk, vv).transpose(1, 2).reshape(-1)
        pos = H.aligned((1,), np.int32, P)
        outs = []
        for form in ("f16", "acc"):
            kc, vc = H.aligned((n_kv, S, hd), np.float16, kc_t.numpy()), H.aligned((n_kv, S, hd), np.float16, vc_t.numpy())
            out = H.aligned((n_head * hd,), np.float16, 0)
            if form == "f16":
                src = H.aligned((nq,), np.float16, qkv_t.numpy())
                H.check(L.teal_attn_decode_f16(H.ptr(src), H.ptr(freqs), H.ptr(kc), H.ptr(vc), H.ptr(pos), H.ptr(out), n_head, n_kv, hd, S, None), "attn")
            else:
                src = H.aligned((nq,), np.float32, qkv_t.float().numpy())     # sums that round to the same fp16 values
                H.check(L.teal_attn_decode_acc_f16(H.ptr(src), H.ptr(freqs), H.ptr(kc), H.ptr(vc), H.ptr(pos), H.ptr(out), n_head, n_kv, hd, S, None), "attn acc")
            got = torch.from_numpy(out.copy()).float()
            assert torch.allclose(got, want, rtol=2e-2, atol=2e-3), (P, form, float((got - want).abs().max()))
            assert np.array_equal(kc.view(np.uint16), kref.numpy().view(np.uint16)) and np.array_equal(vc.view(np.uint16), vref.numpy().view(np.uint16))
            outs.append(out.copy())
        assert np.array_equal(outs[0].view(np.uint16), outs[1].view(np.uint16)), P        # the two input forms agree bit for bit


@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("lanes", [0, 32, 160])
def test_deferred_gate_up_then_down_every_variant(variant, lanes):
    """RMSNorm + gate/up (deferred) -> down-from-sums + residual on ragged shapes for every kernel variant and several K
    chunks, against the same two launches with the pair finaliser in between (fixed RED order => bit-identical x and sum
    of squares), and against the oracle's fp32 truth."""
    L = H.lib()
    L.teal_emu_set_max_workers(1)
    H.set_variant(variant, lanes)
    dim, inter, eps = 392, 1000, 1e-5
    g = torch.Generator().manual_seed(100 + variant)
    xt = torch.randn(1, 1, dim, generator=g).half()
    nwt = (1 + 0.1 * torch.randn(dim, generator=g)).half()
    W1t, W3t = [(torch.randn(inter, dim, generator=g) * 0.05).half() for _ in range(2)]
    W2t = (torch.randn(dim, inter, generator=g) * 0.05).half()
    x, nw = H.aligned((dim,), np.float16, xt.reshape(-1).numpy()), H.aligned((dim,), np.float16, nwt.numpy())
    W1, W3, W2 = H.column_major(W1t.numpy()), H.column_major(W3t.numpy()), H.column_major(W2t.numpy())
    ssq = H.aligned((64 * 32,), np.float32, 0)
    ssq.reshape(32, 64)[:, 0] = float((xt.float() ** 2).sum())
    ws = H.Workspace()
    h16, gu = H.aligned((inter,), np.float16, 0), H.aligned((2 * inter,), np.float32, 0)
    ya, yb = H.aligned((dim,), np.float16, 0), H.aligned((dim,), np.float16, 0)
    sa, sb = H.aligned((64 * 32,), np.float32, 0), H.aligned((64 * 32,), np.float32, 0)
    H.check(L.teal_norm_sparse_gate_up_f16(H.ptr(x), H.ptr(nw), eps, H.ptr(ssq), H.ptr(W1), H.ptr(W3), H.ptr(h16), None, dim, inter, inter, inter,
                                           0.5, 0.7, H.ptr(ws.buf), ws.nbytes, None), "gate_up")
    H.check(L.teal_sparse_gemv_residual_f16(H.ptr(h16), H.ptr(W2), H.ptr(x), H.ptr(ya), H.ptr(sa), inter, dim, dim, 0.03,
                                            H.ptr(ws.buf), ws.nbytes, None), "down")
    H.check(L.teal_norm_sparse_gate_up_acc_f16(H.ptr(x), H.ptr(nw), eps, H.ptr(ssq), H.ptr(W1), H.ptr(W3), H.ptr(gu), dim, inter, inter, inter,
                                               0.5, 0.7, None), "gate_up acc")
    H.check(L.teal_sparse_gemv_residual_ex_f16(None, H.ptr(gu), H.ptr(W2), H.ptr(x), H.ptr(yb), H.ptr(sb), inter, dim, dim, 0.03,
                                               None, 0, H.ptr(ws.buf), ws.nbytes, None), "down from sums")
    ws.assert_clean()
    assert np.array_equal(ya.view(np.uint16), yb.view(np.uint16)) and np.array_equal(sa, sb)
    # (the sums stay: a launch never clears its own input -- another CTA may still have to read the slice; in the decode
    #  step the NEXT layer's o projection clears them)
    assert gu.any()
    xn = O.rms_norm(xt, nwt, eps)
    h_ref = (torch.nn.functional.silu(O.sparse_linear_f32(xn, W1t, 0.5).half()) * O.sparse_linear_f32(xn, W3t, 0.7).half())
    want = (xt.reshape(-1) + O.sparse_linear_f32(h_ref.reshape(1, 1, -1), W2t, 0.03).reshape(-1).half()).float()
    got = torch.from_numpy(ya.copy()).float()
    assert ((got - want).abs() <= want.abs() * 2.0 ** -7 + 2e-2).all(), float((got - want).abs().max())
