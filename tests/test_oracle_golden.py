"""Pins oracle/teal_oracle.py against outputs of the reference's own code
(tests/golden/*.npz, produced by oracle/gen_golden.py)."""
import os

import numpy as np
import torch

from oracle import teal_oracle as O


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


def test_icdf_known_answers(golden_dir):
    g = _load(golden_dir, "icdf_kat.npz")
    qs = g["qs"]
    tags = sorted({k.split("__")[0] for k in g.files if "__" in k})
    assert len(tags) == 3
    for tag in tags:
        for ht in ("h1", "h2"):
            d = O.Distribution(g[f"{tag}__{ht}_centers"], g[f"{tag}__{ht}_counts"])
            got = np.array([float(d.icdf(float(q))) for q in qs])
            assert np.array_equal(got, g[f"{tag}__{ht}_icdf"]), (tag, ht)
            thr = [O.sparsity_to_threshold_hf(d, float(s)) for s in g[f"{tag}__{ht}_hf_s"]]
            assert np.array_equal(np.array(thr), g[f"{tag}__{ht}_hf_thr"])
            assert thr[0] == 0.0  # utils/utils.py:29 special case
            # gpt-fast path has no special case (generate.py:279)
            assert O.sparsity_to_threshold_gptfast(d, 0.0) == float(d.icdf(0.5))


def test_survey_kat_llama2_7b_layer0(golden_dir):
    # SURVEY.md section 8c known answers, recomputed from the committed histogram
    g = _load(golden_dir, "icdf_kat.npz")
    d = O.Distribution(g["Llama_2_7B_L0_mlp__h1_centers"], g["Llama_2_7B_L0_mlp__h1_counts"])
    assert abs(O.sparsity_to_threshold_gptfast(d, 0.5) - 0.02037687972187996) < 1e-12
    d2 = O.Distribution(g["Llama_2_7B_L0_mlp__h2_centers"], g["Llama_2_7B_L0_mlp__h2_counts"])
    assert abs(O.sparsity_to_threshold_gptfast(d2, 0.5) - 0.00060872809) < 1e-9


def test_sparsify_and_linear(golden_dir):
    g = _load(golden_dir, "sparsify_linear.npz")
    names = sorted({k.split("__")[0] for k in g.files})
    n_lin = 0
    for name in names:
        if f"{name}__W" not in g.files:
            continue
        n_lin += 1
        x = torch.from_numpy(g[f"{name}__x"])
        W = torch.from_numpy(g[f"{name}__W"])
        thr = float(g[f"{name}__thr"])
        xm = O.sparsify_forward(x, thr)
        assert np.array_equal(xm.numpy().view(np.uint16), g[f"{name}__xm"].view(np.uint16)), name
        # effective threshold is the fp16-rounded one
        t16 = O.round_threshold_fp16(thr)
        keep = np.abs(g[f"{name}__x"].astype(np.float32)) > np.float32(t16)
        assert np.array_equal(keep.reshape(-1), O.survivor_mask(x, thr).numpy().reshape(-1))
        y = O.sparse_linear(x, W, thr)
        # same code path (torch CPU fp16 matmul): identical on the same build, else 1 ulp
        ref = g[f"{name}__y16"].astype(np.float32)
        assert np.allclose(y.float().numpy(), ref, rtol=2e-3, atol=2e-3), name
        y32 = O.sparse_linear_f32(x, W, thr).numpy()
        assert np.allclose(y32, g[f"{name}__y32"], rtol=1e-5, atol=1e-5), name
    assert n_lin == 6
    # all-dropped case really is all zeros; threshold 0 keeps every non-zero lane
    assert not g["all_dropped_z128_n64__xm"].any()
    assert O.survivor_indices(torch.from_numpy(g["thr0_z128_n64__x"]), 0.0).size == \
        np.count_nonzero(g["thr0_z128_n64__x"])


def test_tie_and_prefill(golden_dir):
    g = _load(golden_dir, "sparsify_linear.npz")
    x = torch.from_numpy(g["tie__x"])
    thr = float(g["tie__thr"])
    xm = O.sparsify_forward(x, thr)
    assert np.array_equal(xm.numpy().view(np.uint16), g["tie__xm"].view(np.uint16))
    # |x| == fp16(thr) is dropped although fp16(thr) > thr would keep it in an fp32 compare
    assert xm[0, 0, 0] == 0 and xm[0, 0, 1] == 0 and xm[0, 0, 2] != 0
    xp = torch.from_numpy(g["prefill__x"])
    t = float(g["prefill__thr"])
    assert np.array_equal(O.sparsify_forward(xp, t, True).numpy(), g["prefill__xm"])
    assert np.array_equal(O.sparsify_forward(xp, t, False).numpy(), g["prefill_off__xm"])
    # first half untouched, second half sparsified (utils/utils.py:36-43)
    half = xp.shape[1] // 2
    assert np.array_equal(g["prefill__xm"][:, :-half], g["prefill__x"][:, :-half])


def test_nan_poisons_output():
    x = torch.tensor([[[1.0, float("nan"), 0.1, 2.0]]], dtype=torch.float16)
    W = torch.ones(3, 4, dtype=torch.float16)
    y = O.sparse_linear(x, W, 0.5)
    assert torch.isnan(y).all()


def test_hf_mlp_forward(golden_dir):
    g = _load(golden_dir, "hf_mlp.npz")
    x = torch.from_numpy(g["x"])
    tg, tu, td = (float(t) for t in g["thr"])
    y = O.ffn_forward(x, torch.from_numpy(g["w_gate"]), torch.from_numpy(g["w_up"]),
                      torch.from_numpy(g["w_down"]), tg, tu, td)
    assert np.allclose(y.float().numpy(), g["y"].astype(np.float32), rtol=2e-3, atol=2e-3)


def test_greedy_lookup(golden_dir):
    g = _load(golden_dir, "greedy_lookup.npz")
    cols = [str(c) for c in g["columns"]]
    for layer in (0, 1):
        tab = g[f"layer{layer}_table"]
        rows = [dict(zip(cols, r)) for r in tab]
        for target in (0.25, 0.4, 0.5, 0.9):
            sel = O.greedy_row_lookup(rows, target)
            want = g[f"target_{target}"][layer]
            assert np.array_equal(np.array([sel[p] for p in O.PROJS]), want), (layer, target)


def test_gptfast_dense_decoder(golden_dir):
    g = _load(golden_dir, "gptfast_dense.npz")
    cfg = {str(k): (float(v) if str(k) == "rope_base" else int(v))
           for k, v in zip(g["cfg_keys"], g["cfg_vals"])}
    w = {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("w__")}
    dec = O.OracleDecoder(cfg, w, [None] * cfg["n_layer"], max_seq=16)
    with torch.no_grad():
        out = dec.forward(torch.from_numpy(g["prompt"]), torch.arange(5))
        assert np.allclose(out.float().numpy(), g["prefill_logits"].astype(np.float32), rtol=5e-3, atol=5e-3)
        for j, t in enumerate(g["next"]):
            out = dec.forward(torch.tensor([[int(t)]]), torch.tensor([5 + j]))
            assert np.allclose(out.float().numpy(), g[f"decode_logits_{j}"].astype(np.float32),
                               rtol=5e-3, atol=5e-3)


def test_tp_sharding_matches_unsharded():
    torch.manual_seed(1)
    Z, I = 64, 128
    x = torch.randn(1, 1, Z).half()
    w1, w3 = (torch.randn(I, Z) * 0.1).half(), (torch.randn(I, Z) * 0.1).half()
    w2 = (torch.randn(Z, I) * 0.1).half()
    full = O.ffn_forward(x, w1, w3, w2, 0.5, 0.5, 0.01)
    for world in (2, 4):
        tp = O.tp_ffn_forward(x, w1, w3, w2, 0.5, 0.5, 0.01, world)
        assert torch.allclose(tp.float(), full.float(), rtol=1e-2, atol=2e-3)
    # qkv sharding: per-rank [q_r; k_r; v_r] (tp.py:73-78)
    wqkv = torch.arange(16 * 2).reshape(16, 2).float()
    s = O.tp_shard_qkv(wqkv, [8, 4, 4], rank=1, world=2)
    assert s[:, 0].tolist() == [8., 10., 12., 14., 20., 22., 28., 30.]
