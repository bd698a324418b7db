"""Opt-in kernels through the SAME parity tests as the production kernel: row-chunk generations 2 / 3
(rowchunk_gemv2_kernel: shared-memory survivor list, rolling load ring, FHFMA, optional speculative L2 prefetch, warp-wise
completion) and the flag-chained decode step (TEAL_CHAIN_*).

It was written after round 1's GPU budget was spent: compiled for sm_100a and SASS-checked, never run on hardware.
Nothing selects it by default (teal_set_tuning(TEAL_TUNE_GENERATION, 2) / config ids 42-49 are opt-in), and these
tests are opt-in too (TEAL_GEN2_GPU_TESTS=1) until they have passed on a B200 once.
"""
import os

import pytest

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("TEAL_GEN2_GPU_TESTS") != "1",
                                 reason="generation-2 kernel not yet run on hardware (set TEAL_GEN2_GPU_TESTS=1)")]

import test_decoder_gpu as D  # noqa: E402
import test_gemv_gpu as T  # noqa: E402

GEN2_CFGS = [42, 43, 44, 45, 46, 47, 48, 49, 50, 51, 52, 53]   # 30 + variant 12..23 (50+: warp-wise completion)
SPEC = [dict(generation=2, spec_rows=0), dict(generation=2, spec_rows=4, spec_mode=1), dict(generation=2, spec_rows=8, spec_mode=2),
        dict(generation=2, spec_rows=32, spec_mode=1), dict(generation=3, spec_rows=0), dict(generation=3, spec_rows=8, spec_mode=1)]


@pytest.fixture(params=SPEC, ids=lambda s: "gen%d_spec%d_mode%d" % (s["generation"], s["spec_rows"], s.get("spec_mode", 1)))
def gen2(request):
    ops = T._ops()
    ops.set_tuning(**request.param)
    yield request.param
    ops.set_tuning()
    ops.set_config(0)


@pytest.mark.parametrize("cfg", GEN2_CFGS)
def test_golden_cases_every_gen2_variant(golden_dir, cfg):
    T.test_golden_cases_all_configs(golden_dir, cfg)


@pytest.mark.parametrize("cfg", GEN2_CFGS)
@pytest.mark.parametrize("sparsity", [0.0, 0.4, 0.9, 1.0])
def test_sparsity_sweep_every_gen2_variant(sparsity, cfg):
    T.test_sparsity_sweep_config1(sparsity, cfg)


def test_survivors_and_ties(gen2, golden_dir):
    for Z, thr in [(8, 0.5), (200, 0.3), (4096, 0.6745), (4096, 0.0), (4096, 100.0), (65536, 1.0)]:
        T.test_survivor_set_bit_exact(Z, thr)
    T.test_survivor_tie_rule(golden_dir)


@pytest.mark.parametrize("Z,N", T.SHAPES)
def test_full_size_vs_oracle(gen2, Z, N):
    T.test_full_size_vs_oracle(Z, N, 0.5)


def test_fused_entry_points(gen2):
    for Z, N, kv in [(4096, 12288, 4096), (4096, 6144, 1024), (8192, 10240, 1024), (8192, 1280, 128)]:
        for uniform in (True, False):
            T.test_qkv_three_thresholds(Z, N, kv, uniform)
    for Z, I in [(4096, 11008), (4096, 14336), (8192, 3584), (128, 352)]:
        for same in (True, False):
            T.test_gate_up_fused(Z, I, same)
    for Z, N, kv in [(4096, 6144, 1024), (8192, 10240, 1024), (256, 384, 64)]:
        for uniform in (True, False):
            T.test_fused_rmsnorm_qkv(Z, N, kv, uniform)
    for Z, I in [(4096, 14336), (4096, 11008), (128, 352)]:
        T.test_fused_rmsnorm_gate_up_and_residual_down(Z, I)
    T.test_nan_propagates_like_reference()
    T.test_linearity_property_full_size()
    T.test_cuda_graph_capture()


def test_decode_step(gen2):
    for mode, thr in [("dense_own", None), ("sparse", D.THR)]:
        D.test_decode_logits_match_oracle(mode, thr)
    D.test_graph_replay_equals_eager_and_is_deterministic_in_tokens()


@pytest.mark.parametrize("norms", [False, True])
def test_flag_chained_blocks_in_the_decode_step(norms):
    """TEAL_CHAIN_MLP / TEAL_CHAIN_ATTN: consumers start on the tiles / heads they need (no griddepcontrol.wait); logits == oracle,
    graph replay == eager, tokens deterministic."""
    dec, orc = D._mk(D.THR)
    dec.chain_mlp = dec.chain_attn = True
    dec.chain_norms = norms
    import torch
    toks = [3, 17, 42, 8, 77, 300, 5, 11]
    dec.capture("sparse")
    dec.reset(toks[0])
    for i, t in enumerate(toks):
        dec.token.fill_(t)
        dec.graph.replay()
        torch.cuda.synchronize()
        with torch.no_grad():
            want = orc.forward(torch.tensor([[t]]), torch.tensor([i])).float().reshape(-1)
        got = dec.logits.float().cpu().reshape(-1)
        assert float((got - want).abs().max()) <= 6e-2 * float(want.abs().max()) + 2e-2, i
    a = dec.generate(1, 24, mode="sparse").cpu().tolist()
    b = dec.generate(1, 24, mode="sparse").cpu().tolist()
    assert a == b
