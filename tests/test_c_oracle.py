"""The plain-C restatement (oracle/teal_oracle_c.c) against the reference's golden outputs and the torch oracle."""
import os

import numpy as np
import torch

from oracle import c_oracle as C
from oracle import teal_oracle as O


def test_half_conversions_bit_exact():
    lib = C.load()
    allh = np.arange(65536, dtype=np.uint16)
    f = allh.view(np.float16).astype(np.float32)
    for h in list(range(0, 65536, 257)) + [0, 1, 0x3ff, 0x400, 0x7bff, 0x7c00, 0x8000, 0xfbff]:
        got = lib.teal_ref_h2f(h)
        want = f[h]
        assert (np.isnan(got) and np.isnan(want)) or got == want, h
    rng = np.random.RandomState(0)
    vals = np.concatenate([rng.standard_normal(2000) * 10 ** rng.uniform(-8, 5, 2000), [0.0, 65504.0, 65519.9, 65520.0, 6e-8, 2.98e-8, 3.0e-8]])
    for v in vals:
        assert lib.teal_ref_f2h(float(v)) == np.float16(v).view(np.uint16), v
    for t in (0.02037687972187996, 0.02038193, 0.6745, 1e-9, 7e4):
        assert lib.teal_ref_round_threshold(t) == O.round_threshold_fp16(t)


def test_c_oracle_matches_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "sparsify_linear.npz"))
    for name in sorted({k.split("__")[0] for k in g.files if k.endswith("__W")}):
        x, W, thr = g[f"{name}__x"], g[f"{name}__W"], float(g[f"{name}__thr"])
        idx = C.survivors(x, thr)
        assert np.array_equal(idx, np.flatnonzero(g[f"{name}__xm"].reshape(-1) != 0) if thr > 0 else O.survivor_indices(torch.from_numpy(x), thr))
        assert np.array_equal(idx, O.survivor_indices(torch.from_numpy(x), thr))
        y16, y32 = C.sparse_gemv(x, W, thr)
        assert np.allclose(y32, g[f"{name}__y32"].reshape(-1), rtol=1e-5, atol=1e-6), name
        ref16 = g[f"{name}__y16"].reshape(-1).astype(np.float32)
        assert np.all(np.abs(y16.astype(np.float32) - ref16) <= np.abs(ref16) * 2.0 ** -10 + 1e-4), name
    tie = C.survivors(g["tie__x"], float(g["tie__thr"]))
    assert np.array_equal(tie, np.flatnonzero(g["tie__xm"].reshape(-1) != 0))
