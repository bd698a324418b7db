"""teal_b200.model.SparseDecoder's own Python control flow on the CPU: the class is instantiated on CPU tensors with the
CUDA plumbing (streams, availability check, workspace) stubbed and the library replaced by the CPU emulation of the kernels
(tests/emu), so `_enqueue_step` -- the exact sequence of C-ABI calls the captured graph contains, including the sampling and
TP-route branches -- runs for real and is checked against OracleDecoder.  CUDA-graph capture itself is not
exercised here (that is what the -m gpu tests do)."""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import teal_oracle as O

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu"))
import harness as H  # noqa: E402
from test_emu_kernels import CFG, THR, TOKS, _state_dict  # noqa: E402


class _Stream:
    cuda_stream = 0

    def synchronize(self):
        pass


@pytest.fixture
def cpu_decoder(monkeypatch):
    from teal_b200 import _lib, model, ops
    emu = H.lib()
    H.set_config(0)
    emu.teal_emu_set_concurrent(0)
    monkeypatch.setattr(_lib, "load", lambda: emu)
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(torch.cuda, "Stream", lambda device=None: _Stream())
    monkeypatch.setattr(torch.cuda, "current_stream", lambda device=None: _Stream())
    monkeypatch.setattr(torch.cuda, "synchronize", lambda device=None: None)
    import contextlib
    monkeypatch.setattr(torch.cuda, "stream", lambda s: contextlib.nullcontext())
    ws = torch.zeros(emu.teal_workspace_bytes(1 << 17), dtype=torch.uint8)
    monkeypatch.setattr(ops, "workspace", lambda dev, sp: ws)   # (the decoder allocates its own with torch.zeros: a CPU tensor here)

    def make(thresholds, **env):
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        dec = model.SparseDecoder(model.ModelArgs(**CFG), device="cpu", max_seq=64, state_dict=_state_dict(CFG))
        dec.set_thresholds(thresholds)
        return dec

    yield make
    H.set_config(0)


def _check(dec, mode, thresholds):
    orc = O.OracleDecoder(CFG, _state_dict(CFG), thresholds, max_seq=64)
    dec.reset(TOKS[0])
    for i, t in enumerate(TOKS):
        dec.token.fill_(t)
        dec._enqueue_step(mode)
        with torch.no_grad():
            want = orc.forward(torch.tensor([[t]]), torch.tensor([i])).float().reshape(-1)
        got = dec.logits.float().reshape(-1)
        tol = (6e-2 if mode == "sparse" else 2e-2) * float(want.abs().max()) + 2e-2
        assert float((got - want).abs().max()) <= tol, (mode, i)
        assert int(dec.pos) == i + 1 and int(dec.token) == int(torch.argmax(got))


@pytest.mark.parametrize("mode", ["sparse", "dense_own"])
def test_default_step_sequence(cpu_decoder, mode):
    dec = cpu_decoder(THR)
    _check(dec, mode, THR if mode == "sparse" else [None] * CFG["n_layer"])
    assert dec.kernels_per_step == 1 + 5 * CFG["n_layer"] + 2


def test_tp_route_at_world_one(cpu_decoder):
    """force_tp_path: plain row-parallel GEMV -> (all-reduce, a no-op at world 1) -> residual add, the route TP ranks take
    when the peer-memory exchange is unavailable -- two more launches per layer, same logits."""
    dec = cpu_decoder(THR)
    dec.force_tp_path = True
    _check(dec, "sparse", THR)
    assert dec.kernels_per_step == 1 + 7 * CFG["n_layer"] + 2


def test_gemv_fused_exchange_route_at_world_one(cpu_decoder):
    """The default TP route (exchange fused into the o / down tile finalisers) with one rank and local receive buffers: the
    multi-GPU launch sequence -- TP kernel twin fed by the deferred gate / up sums, the o projection clearing the deferred
    buffer -- with five launches per layer and the single-GPU logits (tests/test_decoder_gpu.py runs the same on a B200)."""
    import ctypes
    dec = cpu_decoder(THR)
    bufs = [torch.zeros(dec.args.dim, dtype=torch.int64) for _ in (0, 1)]
    dec.tp_recv = [(ctypes.c_void_p * 1)(t.data_ptr()) for t in bufs]
    dec.tp_tile_epochs = [torch.zeros(64, dtype=torch.int32) for _ in (0, 1)]
    dec.tp_fused, dec.tp_gemv, dec.force_tp_path = True, True, True
    assert dec.deferred_epilogue
    _check(dec, "sparse", THR)
    assert dec.kernels_per_step == 1 + 5 * CFG["n_layer"] + 2
    assert int(dec.tp_tile_epochs[0].max()) == len(TOKS) * CFG["n_layer"] and not dec.defer_acc[: dec.nqkv_l].any()


def test_two_rank_decode_step_on_the_emulation(cpu_decoder, monkeypatch):
    """SparseDecoder at world = 2 (gpt-fast/tp.py sharding of ONE state dict, exchange fused into the o / down GEMV
    finalisers, deferred epilogue on the column-parallel launches), both ranks driven from two host threads against the
    emulation with shared host buffers in place of the NVLink peer-mapped ones: the replicas' logits are bit-identical
    and equal the single-rank decoder's to the TP rounding bound.  (The hardware run of the same is tests/test_tp_gpu.py.)"""
    import ctypes
    import threading
    from teal_b200 import model
    world, dim = 2, CFG["dim"]
    recv = [[torch.zeros(world * dim // 2, dtype=torch.int64) for _ in range(world)] for _ in (0, 1)]   # [parity][rank]: {2 x fp16, epoch} words

    def wire(self):
        self.tp_recv = [(ctypes.c_void_p * world)(*[recv[par][q].data_ptr() for q in range(world)]) for par in (0, 1)]
        self.tp_tile_epochs = [torch.zeros(64, dtype=torch.int32) for _ in (0, 1)]
        self.tp_fused, self.tp_gemv = True, True

    monkeypatch.setattr(model.SparseDecoder, "_setup_peer_exchange", wire)
    H.lib().teal_emu_set_concurrent(1)     # finalisers poll for the peer's words: every CTA of a grid must be alive
    single = cpu_decoder(THR)
    sd = _state_dict(CFG)
    logits, errors = {0: [], 1: []}, []

    def rank_main(r):
        try:
            dec = model.SparseDecoder(model.ModelArgs(**CFG), device="cpu", max_seq=64, rank=r, world=world, state_dict=sd)
            dec.set_thresholds(THR)
            assert dec.deferred_epilogue and dec.nqkv_l == single.nqkv_l // world and dec.inter_l == CFG["intermediate_size"] // world
            dec.reset(TOKS[0])
            for t in TOKS:
                dec.token.fill_(t)
                dec._enqueue_step("sparse")
                logits[r].append(dec.logits.clone().reshape(-1))
            assert dec.kernels_per_step == 1 + 5 * CFG["n_layer"] + 2 and not dec.defer_acc[: dec.nqkv_l].any()
        except Exception:  # noqa: BLE001
            import traceback
            errors.append(traceback.format_exc())

    threads = [threading.Thread(target=rank_main, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=300)
    H.lib().teal_emu_set_concurrent(0)
    assert not errors and not any(t.is_alive() for t in threads), errors
    single.reset(TOKS[0])
    for i, t in enumerate(TOKS):
        single.token.fill_(t)
        single._enqueue_step("sparse")
        want = single.logits.float().reshape(-1)
        assert torch.equal(logits[0][i], logits[1][i]), i                                     # replicas bit-identical
        d = float((logits[0][i].float() - want).abs().max())
        assert d <= 2e-2 * float(want.abs().max()) + 2e-2, (i, d)                              # the bound of tools/tp_check.py


def test_position_is_bounded(cpu_decoder):
    """The device position advances without the host looking: generate() / decode_one_token() refuse to run past the KV
    cache, the RoPE table has one row per cache position (model.py:573)."""
    dec = cpu_decoder(THR)
    assert dec.freqs.shape[0] == dec.max_seq
    dec.steps_enqueued = dec.max_seq - 1
    dec._check_room(1)
    with pytest.raises(ValueError):
        dec._check_room(2)
    dec.reset(1)
    assert dec.steps_enqueued == 0
    with pytest.raises(ValueError):
        dec._check_room(dec.max_seq + 1)


def test_sampling_branch(cpu_decoder):
    dec = cpu_decoder(THR)
    dec.set_sampling(0.8, 20, seed=5)
    dec.reset(3)
    for i in range(3):
        dec._enqueue_step("sparse")
        want, _ = O.sample_topk(dec.logits.reshape(-1).clone(), 0.8, 20, O.sample_noise(5, CFG["vocab_size"], i))
        assert int(dec.token) == want
    assert dec.kernels_per_step == 1 + 5 * CFG["n_layer"] + 3
    dec.set_sampling(0.0)
    assert dec.sampling is None


def test_calibration_runs_through_the_taps(cpu_decoder):
    dec = cpu_decoder([None] * CFG["n_layer"])
    thr = dec.calibrate(0.5, n_tokens=3)
    assert len(thr) == CFG["n_layer"] and all(v > 0 for t in thr for v in t.values())


def test_prefill_is_the_dense_branch(cpu_decoder):
    """SparseDecoder.prefill: the prompt position by position with nothing dropped == the reference's S > 1 dense prefill
    (kernels/sparse_gemv.py:271,298), then the sparse decode continues from the filled KV cache."""
    dec = cpu_decoder(THR)
    prompt = [3, 17, 42, 8]
    dec.prefill(prompt)
    orc = O.OracleDecoder(CFG, _state_dict(CFG), THR, max_seq=64)
    with torch.no_grad():
        want = orc.forward(torch.tensor([prompt]), torch.arange(len(prompt)))[:, -1].float().reshape(-1)
    got = dec.logits.float().reshape(-1)
    assert float((got - want).abs().max()) <= 2e-2 * float(want.abs().max()) + 2e-2
    assert int(dec.pos) == len(prompt) and dec.steps_enqueued == len(prompt)
    t = int(dec.token)
    dec._enqueue_step("sparse")
    with torch.no_grad():
        want = orc.forward(torch.tensor([[t]]), torch.tensor([len(prompt)])).float().reshape(-1)
    assert float((dec.logits.float().reshape(-1) - want).abs().max()) <= 6e-2 * float(want.abs().max()) + 2e-2
