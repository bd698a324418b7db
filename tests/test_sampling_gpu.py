"""The sampling kernel on the GPU against the reference's own `sample()` (golden vectors) and inside the decode step
(first hardware run: profiles/r02_call1_gpu_tests_sampling.log)."""
import os

import numpy as np
import pytest
import torch

from oracle import teal_oracle as O

pytestmark = pytest.mark.gpu


def _sample(logits, T, k, q=None, seed=0, pos=0):
    from teal_b200 import _lib
    lib = _lib.load()
    V = logits.numel()
    lg = logits.cuda()
    qd = None if q is None else q.cuda()
    cell = torch.zeros(1, dtype=torch.int64, device="cuda")
    posd = torch.tensor([pos], dtype=torch.int32, device="cuda")
    probs = torch.zeros(V, dtype=torch.float16, device="cuda")
    _lib.check(lib.teal_sample_topk_f16(lg.data_ptr(), V, T, 0 if k is None else k, None if qd is None else qd.data_ptr(), seed,
                                        posd.data_ptr(), cell.data_ptr(), probs.data_ptr(), torch.cuda.current_stream().cuda_stream), "sample")
    torch.cuda.synchronize()
    return int(0xFFFFFFFF - (int(cell.item()) & 0xFFFFFFFF)), probs.cpu()


def test_same_token_as_the_reference_on_its_own_noise(golden_dir):
    g = np.load(os.path.join(golden_dir, "sampling.npz"))
    for n in sorted({k.rsplit("__", 1)[0] for k in g.files}):
        T, k = float(g[n + "__T"]), int(g[n + "__k"])
        if T == 0.0:
            continue
        logits, q, probs = (torch.from_numpy(g[n + s]) for s in ("__logits", "__q", "__probs"))
        tok, p = _sample(logits, T, None if k < 0 else k, q)
        assert np.array_equal((p == 0).numpy(), (probs == 0).numpy()), n
        assert ((p.float() - probs.float()).abs() <= probs.float() * 2.0 ** -10 + 1e-7).all(), n
        if tok != int(g[n + "__token"]):
            top2 = torch.topk(probs.float() / q.float(), 2).values
            assert float(top2[0] - top2[1]) <= float(top2[0]) * 2.0 ** -9, n


def test_full_vocabulary_and_own_noise():
    g = torch.Generator().manual_seed(0)
    logits = (torch.randn(128256, generator=g) * 3).half()
    q = O.sample_noise(5, 128256, 17)
    want, _ = O.sample_topk(logits, 0.8, 200, q)
    got, p = _sample(logits, 0.8, 200, None, seed=5, pos=17)
    assert got == want and int((p > 0).sum()) >= 200


def test_decoder_samples_inside_the_captured_step():
    from test_decoder_gpu import THR, _mk
    dec, _ = _mk(THR)
    dec.set_sampling(0.8, 50, seed=3)
    a = dec.generate(1, 12, mode="sparse").cpu().tolist()
    b = dec.generate(1, 12, mode="sparse").cpu().tolist()
    assert a == b                                        # same seed, same positions -> same tokens
    dec.set_sampling(0.8, 50, seed=4)
    c = dec.generate(1, 12, mode="sparse").cpu().tolist()
    dec.set_sampling(0.0)
    d = dec.generate(1, 12, mode="sparse").cpu().tolist()
    assert c != a or d != a                              # another seed / greedy takes another path
