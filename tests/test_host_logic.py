"""Host-side logic of the product (no GPU): threshold lookup, greedy lookup, op schemas, config table."""
import os

import numpy as np
import torch

from teal_b200.distribution import Distribution, greedy_row, threshold_for, PROJ_HIST
from teal_b200.model import ModelArgs, build_histogram, precompute_freqs_cis, transformer_configs
from oracle import teal_oracle as O


def test_product_distribution_matches_reference_kats(golden_dir):
    g = np.load(os.path.join(golden_dir, "icdf_kat.npz"))
    for tag in sorted({k.split("__")[0] for k in g.files if "__" in k}):
        for ht in ("h1", "h2"):
            d = Distribution(centers=g[f"{tag}__{ht}_centers"], counts=g[f"{tag}__{ht}_counts"])
            got = np.array([d.icdf(float(q)) for q in g["qs"]])
            assert np.array_equal(got, g[f"{tag}__{ht}_icdf"]), (tag, ht)
            thr = [threshold_for(d, float(s), hf_zero_rule=True) for s in g[f"{tag}__{ht}_hf_s"]]
            assert np.array_equal(np.array(thr), g[f"{tag}__{ht}_hf_thr"])


def test_product_greedy_lookup_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "greedy_lookup.npz"))
    cols = [str(c) for c in g["columns"]]
    for layer in (0, 1):
        rows = [dict(zip(cols, r)) for r in g[f"layer{layer}_table"]]
        for target in (0.25, 0.4, 0.5, 0.9):
            sel = greedy_row(rows, target)
            assert np.array_equal(np.array([sel[p] for p in O.PROJS]), g[f"target_{target}"][layer])


def test_op_schemas_match_reference(golden_dir):
    from teal_b200.ops import SparseGEMV, SparseQKVGEMV
    want = dict(l.strip().split("(", 1) for l in open(os.path.join(golden_dir, "op_schemas.txt")))
    assert SparseGEMV.schematize() == "(" + want["sparse_gemv"]
    assert SparseQKVGEMV.schematize() == "(" + want["sparse_qkv_gemv"]
    k = SparseGEMV.initialize("sparse_gemv", "cuda")
    assert k.name == "sparse_gemv" and k.target == "cuda" and k.schema == SparseGEMV.SCHEMA


def test_model_configs_match_gpt_fast():
    a = ModelArgs.from_name("llama-3-8b")
    assert (a.n_layer, a.n_head, a.n_local_heads, a.dim, a.intermediate_size, a.vocab_size, a.head_dim) == \
        (32, 32, 8, 4096, 14336, 128256, 128)
    a7 = ModelArgs.from_name("7B")
    assert a7.intermediate_size == 11008 and a7.n_local_heads == 32 and a7.vocab_size == 32000  # model.py:45-49
    a70 = ModelArgs.from_name("70B")
    assert (a70.dim, a70.intermediate_size, a70.n_layer, a70.n_local_heads) == (8192, 28672, 80, 8)
    assert ModelArgs.from_name("Mistral-7B").intermediate_size == 14336
    assert set(PROJ_HIST) == set(O.PROJS)
    assert torch.equal(precompute_freqs_cis(16, 64, 10000), O.precompute_freqs_cis(16, 64, 10000))


def test_histogram_round_trip_gives_target_sparsity():
    torch.manual_seed(0)
    v = torch.randn(200000)
    c, n = build_histogram(v)
    d = Distribution(centers=c, counts=n)
    for s in (0.3, 0.5, 0.7):
        t = threshold_for(d, s)
        assert abs(float((v.abs() <= t).float().mean()) - s) < 0.01
