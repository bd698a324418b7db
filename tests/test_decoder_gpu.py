"""Decode-step parity: the CUDA-graph decoder (fused kernels) against the CPU OracleDecoder, which is
itself pinned to the reference's gpt-fast Transformer by tests/golden/gptfast_dense.npz."""
import numpy as np
import pytest
import torch

from oracle import teal_oracle as O

pytestmark = pytest.mark.gpu

CFG = dict(block_size=256, vocab_size=512, n_layer=2, n_head=4, dim=256, n_local_heads=2, intermediate_size=704,
           rope_base=10000, norm_eps=1e-5)


def _state_dict(cfg, seed=0):
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


def _mk(thresholds=None):
    from teal_b200.model import ModelArgs, SparseDecoder
    sd = _state_dict(CFG)
    args = ModelArgs(**{k: v for k, v in CFG.items()})
    dec = SparseDecoder(args, max_seq=64, state_dict=sd)
    if thresholds is not None:
        dec.set_thresholds(thresholds)
    orc = O.OracleDecoder(CFG, sd, thresholds if thresholds is not None else [None] * CFG["n_layer"], max_seq=64)
    return dec, orc


THR = [dict(q=0.55, k=0.4, v=0.7, o=0.02, gate=0.6, up=0.5, down=0.01),
       dict(q=0.6, k=0.6, v=0.6, o=0.03, gate=0.5, up=0.5, down=0.02)]


@pytest.mark.parametrize("mode,thr", [("dense_own", None), ("sparse", THR), ("dense_cublas", None)])
def test_decode_logits_match_oracle(mode, thr):
    dec, orc = _mk(thr)
    toks = [3, 17, 42, 8, 77, 300, 5, 11]
    dec.reset(toks[0])
    for i, t in enumerate(toks):
        dec.token.fill_(t)  # teacher-forced token stream
        dec._enqueue_step(mode)
        torch.cuda.synchronize()
        with torch.no_grad():
            ref = orc.forward(torch.tensor([[t]]), torch.tensor([i]))
        got = dec.logits.float().cpu().reshape(-1)
        want = ref.float().reshape(-1)
        # two fp16 pipelines with different summation orders: compare at the logit scale
        scale = float(want.abs().max())
        # sparse mode: a lane within 1 fp16 ulp of a threshold may be kept by one pipeline and dropped by the
        # other (their inputs differ by rounding), which moves a logit by ~|w x|; per-op survivor parity on
        # identical inputs is bit-exact (tests/test_gemv_gpu.py) -- here we bound the end-to-end drift.
        tol = (6e-2 if mode == "sparse" else 2e-2) * scale + 2e-2
        assert float((got - want).abs().max()) <= tol, (mode, i)
        assert int(dec.token.item()) == int(torch.argmax(dec.logits.reshape(-1)))  # fused arg-max == torch.argmax
        assert int(dec.pos.item()) == i + 1


def test_graph_replay_equals_eager_and_is_deterministic_in_tokens():
    dec, _ = _mk(THR)
    dec.reset(5)
    eager = []
    for _ in range(12):
        dec._enqueue_step("sparse")
        torch.cuda.synchronize()
        eager.append(int(dec.token.item()))
    out = dec.generate(5, 12, mode="sparse").cpu().tolist()
    assert out == eager
    out2 = dec.generate(5, 12, mode="sparse").cpu().tolist()
    assert out2 == eager


def test_attention_kernel_vs_sdpa():
    from teal_b200 import _lib
    lib = _lib.load()
    torch.manual_seed(0)
    n_head, n_kv, hd, S = 8, 2, 128, 48
    freqs = __import__("teal_b200.model", fromlist=["x"]).precompute_freqs_cis(S, hd, 10000).cuda().contiguous()
    kc = torch.zeros(n_kv, S, hd, dtype=torch.float16, device="cuda")
    vc = torch.zeros_like(kc)
    out = torch.zeros(n_head * hd, dtype=torch.float16, device="cuda")
    pos = torch.zeros(1, dtype=torch.int32, device="cuda")
    kref = torch.zeros(1, n_kv, S, hd, dtype=torch.float16)
    vref = torch.zeros_like(kref)
    for p in range(20):
        qkv = torch.randn((n_head + 2 * n_kv) * hd).half()
        pos.fill_(p)
        qg = qkv.cuda()
        rc = lib.teal_attn_decode_f16(qg.data_ptr(), freqs.data_ptr(), kc.data_ptr(), vc.data_ptr(), pos.data_ptr(),
                                      out.data_ptr(), n_head, n_kv, hd, S, torch.cuda.current_stream().cuda_stream)
        assert rc == 0
        q, k, v = qkv.split([n_head * hd, n_kv * hd, n_kv * hd])
        fc = freqs.cpu()[p:p + 1]
        q = O.apply_rotary_emb(q.view(1, 1, n_head, hd), fc).transpose(1, 2)
        k = O.apply_rotary_emb(k.view(1, 1, n_kv, hd), fc).transpose(1, 2)
        kref[:, :, p] = k[:, :, 0]
        vref[:, :, p] = v.view(1, n_kv, hd)
        kk = kref.repeat_interleave(n_head // n_kv, dim=1)[:, :, : p + 1].float()
        vv = vref.repeat_interleave(n_head // n_kv, dim=1)[:, :, : p + 1].float()
        y = torch.nn.functional.scaled_dot_product_attention(q.float(), kk, vv).transpose(1, 2).reshape(-1)
        assert torch.allclose(out.float().cpu(), y, rtol=2e-2, atol=2e-3), p
    assert torch.equal(kc.cpu(), kref[0]) and torch.equal(vc.cpu(), vref[0])  # cache append is bit-exact


@pytest.mark.parametrize("hd,n_head,n_kv", [(128, 8, 2), (64, 4, 2)])
def test_attention_over_several_sweeps_of_the_cache(hd, n_head, n_kv):
    """Positions on both sides of every sweep boundary of attn_decode_kernel (a sweep = 128 cache rows; the first sweep is
    requested before the PDL wait, later ones inside the loop), both input forms (fp16 q|k|v and the fp32 sums of a deferred
    QKV launch), against RoPE + SDPA of the oracle over a cache of random rows; the appended row bit-exact.  The CPU twin
    on the emulation is tests/test_emu_kernels.py::test_attention_over_several_sweeps_of_the_cache."""
    from teal_b200 import _lib
    lib = _lib.load()
    S = 320
    g = torch.Generator().manual_seed(hd)
    freqs_t = O.precompute_freqs_cis(S, hd, 10000)
    freqs = freqs_t.cuda().contiguous()
    nq = (n_head + 2 * n_kv) * hd
    sp = torch.cuda.current_stream().cuda_stream
    for P in (0, 1, 127, 128, 129, 200, 255, 256, 257, 319):
        kc_t, vc_t = torch.randn(n_kv, S, hd, generator=g).half(), torch.randn(n_kv, S, hd, generator=g).half()
        kc_t[:, P:], vc_t[:, P:] = 0, 0
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
        pos = torch.full((1,), P, dtype=torch.int32, device="cuda")
        outs = []
        for form in ("f16", "acc"):
            kc, vc = kc_t.cuda().contiguous(), vc_t.cuda().contiguous()
            out = torch.zeros(n_head * hd, dtype=torch.float16, device="cuda")
            if form == "f16":
                src = qkv_t.cuda()
                rc = lib.teal_attn_decode_f16(src.data_ptr(), freqs.data_ptr(), kc.data_ptr(), vc.data_ptr(), pos.data_ptr(), out.data_ptr(),
                                              n_head, n_kv, hd, S, sp)
            else:
                src = qkv_t.float().cuda()
                rc = lib.teal_attn_decode_acc_f16(src.data_ptr(), freqs.data_ptr(), kc.data_ptr(), vc.data_ptr(), pos.data_ptr(), out.data_ptr(),
                                                  n_head, n_kv, hd, S, sp)
            assert rc == 0, lib.teal_last_error()
            torch.cuda.synchronize()
            got = out.float().cpu()
            assert torch.allclose(got, want, rtol=2e-2, atol=2e-3), (P, form, float((got - want).abs().max()))
            assert torch.equal(kc.cpu(), kref) and torch.equal(vc.cpu(), vref)
            outs.append(out.cpu())
        assert torch.equal(outs[0], outs[1]), P


def test_dense_rowmajor_gemv_and_argmax():
    from teal_b200 import _lib
    lib = _lib.load()
    torch.manual_seed(1)
    for Z, V in [(4096, 32000), (256, 512), (8192, 1000)]:
        x = torch.randn(Z).half()
        nw = (1 + 0.1 * torch.randn(Z)).half()
        W = (torch.randn(V, Z) * 0.02).half()
        y = torch.zeros(V, dtype=torch.float16, device="cuda")
        am = torch.zeros(1, dtype=torch.int64, device="cuda")
        xg, ng, Wg = x.cuda(), nw.cuda(), W.cuda()  # keep the device copies alive across the async launch
        rc = lib.teal_dense_gemv_f16(xg.data_ptr(), ng.data_ptr(), 1e-5, Wg.data_ptr(), y.data_ptr(),
                                     am.data_ptr(), Z, V, Z, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        assert rc == 0
        xn = O.rms_norm(x.view(1, 1, Z), nw, 1e-5).view(Z)
        want = W.float() @ xn.float()
        assert torch.allclose(y.float().cpu(), want, rtol=4e-3, atol=4e-3 * float(want.abs().max()))
        idx = 0xFFFFFFFF - (int(am.item()) & 0xFFFFFFFF)
        assert idx == int(torch.argmax(y))


def test_calibration_hits_target_sparsity():
    dec, _ = _mk(None)
    thr = dec.calibrate(0.5, n_tokens=12)
    assert len(thr) == CFG["n_layer"] and set(thr[0]) == set(O.PROJS)
    probe = torch.zeros(2, dtype=torch.int64, device="cuda")
    dec.lib.teal_set_bytes_probe(probe.data_ptr())
    dec.reset(9)
    for _ in range(6):
        dec._enqueue_step("sparse")
    torch.cuda.synchronize()
    dec.lib.teal_set_bytes_probe(None)
    kept = float(probe[0]) / float(probe[1])
    assert 0.35 < kept < 0.65, kept


def test_tp_route_equals_fused_route_on_one_gpu():
    """The TP code path (plain GEMV -> all-reduce -> residual_add kernel) against the fused residual epilogue with a
    single rank: the same fp16 rounding sequence for the residual stream; only the sum of squares is accumulated in a
    different order (per-tile partials vs one total), so the RMSNorm scale may differ in its last fp32 bit and a
    normalised lane in its last fp16 bit -- logits agree to rounding, tokens exactly."""
    a, _ = _mk(THR)
    b, _ = _mk(THR)
    b.force_tp_path = True
    a.reset(7); b.reset(7)
    for i in range(6):
        a._enqueue_step("sparse"); b._enqueue_step("sparse")
        torch.cuda.synchronize()
        d = float((a.logits.float() - b.logits.float()).abs().max())
        assert d <= 2e-2 * float(a.logits.float().abs().max()) + 1e-3, (i, d)
        b.token.copy_(a.token)  # (a near-tie in the arg-max must not send the two decoders down different paths)


@pytest.mark.parametrize("geometry", ["toy", "llama-3-8b"])
def test_deferred_epilogue_equals_finaliser_route(geometry):
    """The decoder's default route (QKV / gate+up leave fp32 split-K sums; attention / the down projection round them on
    load; the o projection clears the buffer) against the route where every launch finalises its own tiles.  The
    roundings are the same single fp32 -> fp16 conversions (bit-identical on the emulation with a fixed RED order,
    tests/test_emu_kernels.py); on the GPU the order of the fp32 REDs differs from run to run on BOTH routes, so the
    comparison is to rounding, step by step, eager and from the captured graph -- and the deferred buffer must come back
    cleared after every step (its q|k|v part) so that nothing accumulates over a long generation."""
    from teal_b200.model import ModelArgs, SparseDecoder
    if geometry == "toy":
        a, _ = _mk(THR)
        b, _ = _mk(THR)
    else:
        cfg = REAL_LAYERS[geometry]
        a = SparseDecoder(ModelArgs(**cfg), max_seq=64, seed=11)
        b = SparseDecoder(ModelArgs(**cfg), max_seq=64, seed=11)
        thr = a.calibrate(0.5, n_tokens=4)
        a.set_thresholds(thr); b.set_thresholds(thr)
    assert a.deferred_epilogue
    b.deferred_epilogue = False
    a.reset(7); b.reset(7)
    for i in range(10):
        a._enqueue_step("sparse"); b._enqueue_step("sparse")
        torch.cuda.synchronize()
        scale = float(b.logits.float().abs().max())
        d = float((a.logits.float() - b.logits.float()).abs().max())
        assert d <= (1e-1 if geometry != "toy" else 2e-2) * scale + 1e-3, (i, d, scale)   # (calibrated thresholds: lanes may flip)
        assert not a.defer_acc[: a.nqkv_l].any(), "q|k|v sums not cleared"
        b.token.copy_(a.token)
    # the captured graph replays the same launches (PDL chain: the clear runs while the next kernels already wait)
    ta = a.generate(5, 40, mode="sparse").cpu()
    assert int(ta.min()) >= 0 and int(ta.max()) < a.args.vocab_size
    assert torch.isfinite(a.logits.float()).all() and not a.defer_acc[: a.nqkv_l].any()
    a.reset(7); b.reset(7)                 # and after the replays the two routes still agree step by step
    for i in range(4):
        a._enqueue_step("dense_own"); b._enqueue_step("dense_own")
        torch.cuda.synchronize()
        d = float((a.logits.float() - b.logits.float()).abs().max())
        assert d <= 2e-2 * float(b.logits.float().abs().max()) + 1e-3, (i, d)
        b.token.copy_(a.token)


def test_gemv_fused_tp_exchange_route_with_deferred_epilogue_on_one_gpu():
    """The default tensor-parallel route (o / down GEMV with the exchange in its tile finalisers) at world = 1, with local
    receive buffers standing in for the peer-mapped ones: the push / poll protocol degenerates to one source, everything
    else is the multi-GPU launch sequence -- TP kernel twin, the down projection reading the deferred gate / up sums, the
    o projection clearing the deferred buffer -- against the single-GPU route, eager and from a captured graph.
    (tests/test_tp_gpu.py runs the real 2-rank exchange when the box has two GPUs.)"""
    import ctypes
    a, _ = _mk(THR)
    b, _ = _mk(THR)
    dim = b.args.dim
    bufs = [torch.zeros(dim, dtype=torch.int64, device="cuda") for _ in (0, 1)]          # world * N / 2 eight-byte words, x 2 B spare
    b.tp_recv = [(ctypes.c_void_p * 1)(t.data_ptr()) for t in bufs]
    b.tp_tile_epochs = [torch.zeros(64, dtype=torch.int32, device="cuda") for _ in (0, 1)]
    b.tp_fused, b.tp_gemv, b.force_tp_path = True, True, True
    assert b.deferred_epilogue
    a.reset(7); b.reset(7)
    for i in range(8):
        a._enqueue_step("sparse"); b._enqueue_step("sparse")
        torch.cuda.synchronize()
        d = float((a.logits.float() - b.logits.float()).abs().max())
        assert d <= 2e-2 * float(a.logits.float().abs().max()) + 1e-3, (i, d)
        assert not b.defer_acc[: b.nqkv_l].any()
        b.token.copy_(a.token)
    assert int(b.tp_tile_epochs[0].max()) == 8 * b.args.n_layer                            # one exchange per layer per step on each set
    tb = b.generate(5, 16, mode="sparse").cpu()
    assert int(tb.min()) >= 0 and int(tb.max()) < b.args.vocab_size and torch.isfinite(b.logits.float()).all()


def test_hf_sparse_linear_decode_uses_kernel_and_matches_reference_semantics():
    """HF surface (teal/mlp.py:50-56): SparseLinear at q_len == 1 on CUDA == mask-then-matmul of the reference."""
    from teal_b200.hf import SparseLinear, SparsifyFn
    torch.manual_seed(3)
    Z, N = 512, 1024
    lin = torch.nn.Linear(Z, N, bias=False, dtype=torch.float16)
    lin.weight.data = (torch.randn(N, Z) * 0.05).half()
    fn = SparsifyFn(init_threshold=0.6)
    x = torch.randn(1, 1, Z).half()
    want = O.sparse_linear_f32(x, lin.weight.data, 0.6).reshape(-1)
    sl = SparseLinear(lin, fn).cuda()
    y = sl(x.cuda())
    err = (y.reshape(-1).float().cpu() - want).abs()
    assert (err <= want.abs() * 2.0 ** -10 + 1e-3).all()
    xp = torch.randn(1, 6, Z).half()  # prefill: half-sequence rule, dense matmul
    yp = sl(xp.cuda()).float().cpu()
    wp = torch.nn.functional.linear(O.sparsify_forward(xp, 0.6).float(), lin.weight.data.float().cpu())
    assert torch.allclose(yp, wp, rtol=1e-2, atol=1e-2)


def test_gptfast_call_sites_drop_in():
    """The reference's call sites (gemv1/gemv2/thresh_*/sparsity_bin attrs, apply_monkeypatch, torch.ops.teal.*)
    on a gpt-fast-shaped nn.Module: prefill (dense branch of the ops) + decode steps vs the oracle, then the same
    decode steps replayed from a captured CUDA graph."""
    from teal_b200.gptfast_compat import Transformer, install_sparsity
    from teal_b200.model import ModelArgs
    sd = _state_dict(CFG)
    m = Transformer(ModelArgs(**CFG))
    m.load_state_dict(sd)
    m = m.half().cuda()
    install_sparsity(m, THR)
    m.setup_caches(1, 32)
    orc = O.OracleDecoder(CFG, sd, THR, max_seq=32)
    assert m.layers[0].attention.gemv1 is torch.ops.teal.sparse_qkv_gemv
    assert m.layers[0].feed_forward.w1.weight.stride() == (1, CFG["intermediate_size"])
    with torch.no_grad():
        prompt = torch.tensor([[3, 17, 42, 8]])
        got = m(prompt.cuda(), torch.arange(4, device="cuda"))
        want = orc.forward(prompt, torch.arange(4))
        assert torch.allclose(got.float().cpu(), want.float(), rtol=3e-2, atol=3e-2)
        for i, t in enumerate([77, 300, 5]):
            got = m(torch.tensor([[t]], device="cuda"), torch.tensor([4 + i], device="cuda"))
            want = orc.forward(torch.tensor([[t]]), torch.tensor([4 + i]))
            scale = float(want.abs().max())
            assert float((got.float().cpu() - want.float()).abs().max()) <= 6e-2 * scale + 2e-2
        # capture one decode step of the nn.Module (the ops are graph-safe: no sync, no host allocation)
        tok = torch.tensor([[9]], device="cuda")
        pos = torch.tensor([7], device="cuda")
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            m(tok, pos)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=s):
                out = m(tok, pos)
            g.replay()
        torch.cuda.synchronize()
        want = orc.forward(torch.tensor([[9]]), torch.tensor([7]))
        assert float((out.float().cpu() - want.float()).abs().max()) <= 6e-2 * float(want.abs().max()) + 2e-2


REAL_LAYERS = {
    "llama-3-8b": dict(block_size=8192, vocab_size=1024, n_layer=1, n_head=32, dim=4096, n_local_heads=8, intermediate_size=14336,
                       rope_base=500000, norm_eps=1e-5),
    "llama-2-70b": dict(block_size=2048, vocab_size=1024, n_layer=1, n_head=64, dim=8192, n_local_heads=8, intermediate_size=28672,
                        rope_base=10000, norm_eps=1e-5),
}


@pytest.mark.parametrize("name", sorted(REAL_LAYERS))
def test_one_layer_at_real_geometry_vs_oracle(name):
    """One decoder layer at Llama-3-8B / Llama-2-70B geometry (head_dim 128, GQA 4 / 8, 4096 -> 14336 / 8192 -> 28672;
    small vocabulary) on the GPU against OracleDecoder: the per-op tests cover these shapes one projection at a time,
    this runs the five launches of a real-size layer back to back (sum-of-squares hand-off, residual adds, attention)."""
    from teal_b200.model import ModelArgs, SparseDecoder
    cfg = REAL_LAYERS[name]
    g = torch.Generator().manual_seed(5)
    sd = {k: v for k, v in _state_dict(cfg, seed=3).items()}
    for k in sd:   # smaller weights: the inner dimensions are 16-100x those of the toy model
        if k.endswith(("wqkv.weight", "wo.weight", "w1.weight", "w2.weight", "w3.weight")):
            sd[k] = (sd[k].float() * 0.3).half()
    dec = SparseDecoder(ModelArgs(**cfg), max_seq=32, state_dict=sd)
    thr = dec.calibrate(0.5, n_tokens=4)                      # thresholds from the model's own activations (dense steps)
    orc = O.OracleDecoder(cfg, sd, thr, max_seq=32)
    toks = [3, 17, 42]
    dec.reset(toks[0])
    for i, t in enumerate(toks):
        dec.token.fill_(t)
        dec._enqueue_step("sparse")
        torch.cuda.synchronize()
        with torch.no_grad():
            want = orc.forward(torch.tensor([[t]]), torch.tensor([i])).float().reshape(-1)
        got = dec.logits.float().cpu().reshape(-1)
        scale = float(want.abs().max())
        assert float((got - want).abs().max()) <= 6e-2 * scale + 2e-2, (name, i, float((got - want).abs().max()), scale)


def test_two_decoders_captured_from_two_threads():
    """The launch context (PDL switch, kernel selection, probes) is thread-local: two decoders built, captured and run
    from two host threads at the same time produce the tokens of the same decoders run one after the other."""
    import threading
    want = _mk(THR)[0].generate(5, 12, mode="sparse").cpu().tolist()
    results, errors = {}, []

    def worker(i):
        try:
            dec, _ = _mk(THR)
            for _ in range(3):   # capture (toggles teal_set_pdl in this thread) and replay, repeatedly
                dec.graph = None
                results[i] = dec.generate(5, 12, mode="sparse").cpu().tolist()
        except Exception:  # noqa: BLE001
            import traceback
            errors.append(traceback.format_exc())

    threads = [threading.Thread(target=worker, args=(i,)) for i in range(2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=300)
    assert not errors, errors
    assert results[0] == want and results[1] == want


def test_prefill_then_sparse_decode_matches_reference_semantics():
    """generate.py:68-92,149-217: the prompt goes through the DENSE branch of the ops (S > 1: kernels/sparse_gemv.py:271,298),
    then every new token through the sparse kernels.  SparseDecoder.prefill feeds the prompt position by position with
    nothing dropped; OracleDecoder does the reference's S = 4 dense prefill in one call."""
    dec, orc = _mk(THR)
    prompt = [3, 17, 42, 8]
    dec.prefill(prompt)
    with torch.no_grad():
        want = orc.forward(torch.tensor([prompt]), torch.arange(len(prompt)))[:, -1].float().reshape(-1)
    got = dec.logits.float().cpu().reshape(-1)
    scale = float(want.abs().max())
    assert float((got - want).abs().max()) <= 2e-2 * scale + 2e-2
    assert int(dec.pos.item()) == len(prompt) and int(dec.token.item()) == int(torch.argmax(dec.logits.reshape(-1)))
    # the next (sparse) decode step continues from the prefilled cache on both sides
    t = int(dec.token.item())
    dec._enqueue_step("sparse")
    torch.cuda.synchronize()
    with torch.no_grad():
        want = orc.forward(torch.tensor([[t]]), torch.tensor([len(prompt)])).float().reshape(-1)
    got = dec.logits.float().cpu().reshape(-1)
    assert float((got - want).abs().max()) <= 6e-2 * float(want.abs().max()) + 2e-2
    # public API: prompt in, n_new tokens out, the first one being the prefill's sample
    out = dec.generate(prompt, 6, mode="sparse").cpu().tolist()
    assert len(out) == 6 and out[0] == t


def test_gptfast_checkpoint_file_round_trip(tmp_path):
    """a gpt-fast `model.pth` whose attention weights are still split (wq / wk / wv, as
    scripts/convert_hf_checkpoint.py writes them) -> load_checkpoint (mmap, `Attention.load_hook` fusion, model.py:209-214)
    -> SparseDecoder -> logits of the oracle on the same weights."""
    from teal_b200.model import ModelArgs, SparseDecoder, load_checkpoint
    sd = _state_dict(CFG)
    hd = CFG["dim"] // CFG["n_head"]
    kv = CFG["n_local_heads"] * hd
    split = {}
    for k, v in sd.items():
        if k.endswith("attention.wqkv.weight"):
            p = k[: -len("wqkv.weight")]
            split[p + "wq.weight"], split[p + "wk.weight"], split[p + "wv.weight"] = v.split([CFG["dim"], kv, kv], dim=0)
        else:
            split[k] = v
    path = tmp_path / "model.pth"
    torch.save({k: v.clone() for k, v in split.items()}, path)
    dec = SparseDecoder(ModelArgs(**CFG), max_seq=64, state_dict=load_checkpoint(str(path)))
    dec.set_thresholds(THR)
    orc = O.OracleDecoder(CFG, sd, THR, max_seq=64)
    dec.reset(3)
    for i, t in enumerate([3, 17, 42]):
        dec.token.fill_(t)
        dec._enqueue_step("sparse")
        torch.cuda.synchronize()
        with torch.no_grad():
            want = orc.forward(torch.tensor([[t]]), torch.tensor([i])).float().reshape(-1)
        got = dec.logits.float().cpu().reshape(-1)
        assert float((got - want).abs().max()) <= 6e-2 * float(want.abs().max()) + 2e-2
