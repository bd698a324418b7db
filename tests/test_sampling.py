"""Top-k / temperature sampling (gpt-fast/generate.py:49-66): the oracle restatement against golden vectors produced by
the reference's own `sample()` (tests/golden/sampling.npz, oracle/gen_golden.py), and the CUDA sampling kernel -- run on
the CPU emulation of tests/emu through the C ABI -- against both.  The GPU run of the same cases is in
tests/test_decoder_gpu.py-style gated tests (tests/test_sampling_gpu.py)."""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import teal_oracle as O

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu"))
import harness as H  # noqa: E402


def _cases(golden_dir):
    g = np.load(os.path.join(golden_dir, "sampling.npz"))
    names = sorted({k.rsplit("__", 1)[0] for k in g.files})
    for n in names:
        k = int(g[n + "__k"])
        yield n, torch.from_numpy(g[n + "__logits"]), torch.from_numpy(g[n + "__q"]), torch.from_numpy(g[n + "__probs"]), \
            int(g[n + "__token"]), float(g[n + "__T"]), (None if k < 0 else k)


def test_oracle_matches_the_reference_sample(golden_dir):
    n = 0
    for name, logits, q, probs, token, T, k in _cases(golden_dir):
        tok, p = O.sample_topk(logits, T, k, q)
        assert np.array_equal(p.numpy().view(np.uint16), probs.numpy().view(np.uint16)), name   # probs bit-exact
        assert tok == token, name
        n += 1
    assert n == 21


def _emu_sample(logits, T, k, q=None, seed=0, pos=0, want_probs=True):
    V = logits.numel()
    la = H.aligned((V,), np.float16, logits.numpy())
    qa = None if q is None else H.aligned((V,), np.float16, q.numpy())
    cell = H.aligned((1,), np.uint64, 0)
    posa = H.aligned((1,), np.int32, pos)
    probs = H.aligned((V,), np.float16, 0) if want_probs else None
    H.check(H.lib().teal_sample_topk_f16(H.ptr(la), V, T, 0 if k is None else k, H.ptr(qa), seed, H.ptr(posa), H.ptr(cell), H.ptr(probs), None),
            "sample")
    return int(0xFFFFFFFF - (int(cell[0]) & 0xFFFFFFFF)), (None if probs is None else torch.from_numpy(probs.copy()))


def test_kernel_matches_the_reference_sample(golden_dir):
    """Same token as the reference on the reference's own noise; probs within 1 fp16 ulp (the fp32 sum of exponentials
    is accumulated in a different order), bit-exact zeros outside the top-k set (same survivors of the filter)."""
    for name, logits, q, probs, token, T, k in _cases(golden_dir):
        if T == 0.0:   # the reference divides by 1e-5, overflows fp16 and returns all-NaN probs (the oracle reproduces that
            continue   # bit for bit above); the library treats temperature 0 as greedy arg-max instead (SparseDecoder)
        tok, p = _emu_sample(logits, T, k, q)
        assert np.array_equal((p == 0).numpy(), (probs == 0).numpy()), name
        d = (p.float() - probs.float()).abs()
        assert (d <= probs.float() * 2.0 ** -10 + 1e-7).all(), (name, float(d.max()))
        if tok != token:  # only legitimate when the two best ratios are within one fp16 ulp of each other
            r = (probs.float() / q.float())
            top2 = torch.topk(r, 2).values
            assert float(top2[0] - top2[1]) <= float(top2[0]) * 2.0 ** -9, (name, tok, token)


def test_kernel_own_noise_is_the_oracles_philox_stream():
    g = torch.Generator().manual_seed(3)
    for V, k, T, seed, pos in [(512, 50, 0.8, 0x1234ABCD5678, 0), (777, 200, 1.0, 7, 41), (96, None, 0.7, 2 ** 40 + 3, 5)]:
        logits = (torch.randn(V, generator=g) * 2).half()
        q = O.sample_noise(seed, V, pos)
        want, _ = O.sample_topk(logits, T, k, q)
        got, _ = _emu_sample(logits, T, k, None, seed=seed, pos=pos)
        assert got == want, (V, k, got, want)


def test_philox_known_answer_and_exponential_distribution():
    # Random123 known-answer vectors for Philox4x32-10
    assert O.philox4x32_10(0, 0, 0, 0, 0, 0) == (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)
    assert O.philox4x32_10(0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF) == (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)
    q = O.sample_noise(99, 20000, 3).float().numpy()
    assert q.min() > 0
    assert abs(q.mean() - 1.0) < 0.03 and abs(np.median(q) - np.log(2)) < 0.03 and abs((q > 2).mean() - np.exp(-2)) < 0.01


def test_sampling_frequencies_follow_the_filtered_softmax():
    """argmax(p / Exp(1)) picks i with probability p_i: 3000 draws of the oracle with the library's noise stream."""
    logits = torch.tensor([2.0, 1.0, 0.5, 0.0, -1.0, -3.0, 1.5, 0.2]).half()
    p = O.logits_to_probs(logits, 0.8, 4).float().numpy()
    counts = np.zeros(8)
    for pos in range(3000):
        tok, _ = O.sample_topk(logits, 0.8, 4, O.sample_noise(11, 8, pos))
        counts[tok] += 1
    assert counts[p == 0].sum() == 0
    chi2 = (((counts - 3000 * p) ** 2) / np.maximum(3000 * p, 1e-9))[p > 0].sum()
    assert chi2 < 16.3   # 3 degrees of freedom, p = 0.001
