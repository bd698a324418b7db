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


def test_hf_sparsify_fn_matches_reference_golden(golden_dir):
    """teal_b200.hf.SparsifyFn (CPU path = the reference's mask semantics) against the reference's own outputs."""
    from teal_b200.hf import SparsifyFn
    g = np.load(os.path.join(golden_dir, "sparsify_linear.npz"))
    fn = SparsifyFn(init_threshold=float(g["tie__thr"]))
    assert np.array_equal(fn(torch.from_numpy(g["tie__x"])).numpy().view(np.uint16), g["tie__xm"].view(np.uint16))
    fn = SparsifyFn(init_threshold=float(g["prefill__thr"]))
    xp = torch.from_numpy(g["prefill__x"])
    assert np.array_equal(fn(xp).numpy(), g["prefill__xm"])
    fn.apply_prefill = False
    assert np.array_equal(fn(xp).numpy(), g["prefill_off__xm"])
    k = np.load(os.path.join(golden_dir, "icdf_kat.npz"))
    d = Distribution(centers=k["Llama_2_7B_L0_mlp__h1_centers"], counts=k["Llama_2_7B_L0_mlp__h1_counts"])
    fn = SparsifyFn(d)
    for s, want in zip(k["Llama_2_7B_L0_mlp__h1_hf_s"], k["Llama_2_7B_L0_mlp__h1_hf_thr"]):
        fn.set_threshold(float(s))
        assert fn.threshold == float(want)


def test_hf_mlp_patch_cpu_path_matches_reference(golden_dir):
    from teal_b200.hf import patch_mlp
    g = np.load(os.path.join(golden_dir, "hf_mlp.npz"))

    class MLP(torch.nn.Module):
        def __init__(self):
            super().__init__()
            Z, I = g["w_gate"].shape[1], g["w_gate"].shape[0]
            self.gate_proj = torch.nn.Linear(Z, I, bias=False, dtype=torch.float16)
            self.up_proj = torch.nn.Linear(Z, I, bias=False, dtype=torch.float16)
            self.down_proj = torch.nn.Linear(I, Z, bias=False, dtype=torch.float16)
            self.act_fn = torch.nn.SiLU()
    m = MLP()
    m.gate_proj.weight.data = torch.from_numpy(g["w_gate"])
    m.up_proj.weight.data = torch.from_numpy(g["w_up"])
    m.down_proj.weight.data = torch.from_numpy(g["w_down"])
    patch_mlp(m, {"h1": None, "h2": None})
    for k, t in zip(("gate", "up", "down"), g["thr"]):
        m.sparse_fns[k].threshold = float(t)
    with torch.no_grad():
        y = m(torch.from_numpy(g["x"]))
    assert np.allclose(y.float().numpy(), g["y"].astype(np.float32), rtol=2e-3, atol=2e-3)


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under teal_b200/ may import or execute it."""
    import re
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "teal_b200")
    for dirpath, _, files in os.walk(root):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "teal_oracle" not in src, f


def test_product_has_no_cpu_execution_path():
    """The CPU emulation of the kernels (tests/emu) is test infrastructure like the oracle: the Python package never refers to
    it, the shipped library exports none of its symbols and was built without TEAL_EMU (no emulation twin of any inline-PTX
    helper is in it -- the twins exist only behind `#ifdef TEAL_EMU` in the .cu sources)."""
    import ctypes
    import re
    import subprocess
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "teal_b200")
    for f in os.listdir(root):
        if f.endswith(".py"):
            src = open(os.path.join(root, f)).read()
            assert not re.search(r"tests[/.]emu|libteal_emu|TEAL_EMU|harness", src), f
    from teal_b200 import build
    assert not any("TEAL_EMU" in flag for flag in build.NVCC_FLAGS)
    lib = build.build()
    syms = subprocess.run(["nm", "-D", "--defined-only", lib], capture_output=True, text=True).stdout
    assert "teal_emu_" not in syms and "emu" not in " ".join(s.split()[-1] for s in syms.splitlines() if "teal" in s)
    assert not hasattr(ctypes.CDLL(lib), "teal_emu_set_concurrent")


def test_hf_self_attn_patch_and_setters():
    from teal_b200.hf import SparseLinear, patch_mlp, patch_self_attn, set_sparsities

    class D:  # stand-in Distribution: icdf(q) = q
        def icdf(self, q):
            return q

    class Attn(torch.nn.Module):
        def __init__(self):
            super().__init__()
            for p in "qkvo":
                setattr(self, f"{p}_proj", torch.nn.Linear(16, 16, bias=False, dtype=torch.float16))

    class MLP(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.gate_proj = torch.nn.Linear(16, 32, bias=False, dtype=torch.float16)
            self.up_proj = torch.nn.Linear(16, 32, bias=False, dtype=torch.float16)
            self.down_proj = torch.nn.Linear(32, 16, bias=False, dtype=torch.float16)
            self.act_fn = torch.nn.SiLU()

    class Layer(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.self_attn, self.mlp = Attn(), MLP()

    layers = [Layer(), Layer()]
    for l in layers:
        patch_self_attn(l.self_attn, {"h1": D(), "h2": D()})
        patch_mlp(l.mlp, {"h1": D(), "h2": D()})
    assert isinstance(layers[0].self_attn.q_proj, SparseLinear) and set(layers[0].self_attn.sparse_fns) == {"q", "k", "v", "o"}
    sp = {p: [0.0, 0.5] for p in O.PROJS}
    set_sparsities(layers, sp)
    assert layers[0].mlp.sparse_fns["down"].threshold == 0.0          # utils/utils.py:29
    assert layers[1].self_attn.sparse_fns["o"].threshold == 0.75       # icdf(0.5 + 0.5/2)
    x = torch.randn(1, 3, 16).half()
    y = layers[1].self_attn.q_proj(x)                                   # CPU: reference mask-then-matmul semantics
    want = torch.nn.functional.linear(O.sparsify_forward(x, 0.75), layers[1].self_attn.q_proj.linear.weight)
    assert torch.equal(y, want)


def test_checkpoint_helpers(tmp_path):
    from teal_b200.model import fuse_qkv_state_dict, load_checkpoint
    from teal_b200.distribution import resolve_histogram_dir
    sd = {"layers.0.attention.wq.weight": torch.ones(8, 4), "layers.0.attention.wk.weight": 2 * torch.ones(2, 4),
          "layers.0.attention.wv.weight": 3 * torch.ones(2, 4), "norm.weight": torch.ones(4)}
    f = fuse_qkv_state_dict(sd)
    assert "layers.0.attention.wq.weight" not in f and f["layers.0.attention.wqkv.weight"].shape == (12, 4)
    assert f["layers.0.attention.wqkv.weight"][8:10].eq(2).all() and f["layers.0.attention.wqkv.weight"][10:].eq(3).all()
    p = tmp_path / "model.pth"
    torch.save(sd, p)
    assert set(load_checkpoint(str(p))) == {"layers.0.attention.wqkv.weight", "norm.weight"}
    (tmp_path / "m" / "histograms" / "Untitled" / "layer-0").mkdir(parents=True)
    (tmp_path / "m2" / "histograms" / "layer-0").mkdir(parents=True)
    assert resolve_histogram_dir(str(tmp_path / "m")).endswith("Untitled")
    assert resolve_histogram_dir(str(tmp_path / "m2")).endswith("histograms")


def test_hf_model_level_setters_on_shipped_shaped_fixtures(tmp_path):
    """teal/model.py:95-155: build_sparse_layers / set_uniform_sparsity / set_mlp_sparsity / set_self_attn_sparsity /
    load_greedy_sparsities / reset_sparsities / set_apply_prefill on a two-layer HF-shaped model, with histograms.pt and
    results.csv files in the reference's on-disk layout; SparseLinear keeps ONE copy of the weight."""
    from teal_b200.hf import SparseLinear, SparseModelSetters
    from teal_b200.distribution import Distribution

    class Attn(torch.nn.Module):
        def __init__(self):
            super().__init__()
            for p in "qkvo":
                setattr(self, f"{p}_proj", torch.nn.Linear(16, 16, bias=False, dtype=torch.float16))

    class MLP(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.gate_proj = torch.nn.Linear(16, 32, bias=False, dtype=torch.float16)
            self.up_proj = torch.nn.Linear(16, 32, bias=False, dtype=torch.float16)
            self.down_proj = torch.nn.Linear(32, 16, bias=False, dtype=torch.float16)
            self.act_fn = torch.nn.SiLU()

    class Layer(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.self_attn, self.mlp = Attn(), MLP()

    class Inner(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.layers = torch.nn.ModuleList([Layer(), Layer()])

    class LM(SparseModelSetters, torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.model = Inner()

    g = torch.Generator().manual_seed(0)
    for i in range(2):
        for sub in ("mlp", "self_attn"):
            d = tmp_path / "hist" / f"layer-{i}" / sub
            d.mkdir(parents=True)
            centers = torch.linspace(-2, 2, 101)
            hist = {"h1": torch.rand(101, generator=g) + 0.1, "h1_centers": centers, "h2": torch.rand(101, generator=g) + 0.1,
                    "h2_centers": centers * 0.1}
            torch.save(hist, d / "histograms.pt")
        lk = tmp_path / "lookup" / f"layer-{i}"
        lk.mkdir(parents=True)
        (lk / "results.csv").write_text("Effective Sparsity,Activation Error,Baseline Error,q,k,v,o,gate,up,down\n"
                                        "0.3,0,0,0.3,0.3,0.3,0.3,0.3,0.3,0.3\n0.5,0,0,0.6,0.5,0.2,0.55,0.45,0.4,0.6\n")
    m = LM().configure_sparsity(str(tmp_path / "hist"), uniform_sparsity=0.5, apply_prefill=False)
    d_h1 = Distribution(str(tmp_path / "hist" / "layer-1" / "mlp"), "h1")
    assert m.model.layers[1].mlp.sparse_fns["gate"].threshold == d_h1.icdf(0.75)
    assert m.model.layers[0].self_attn.sparse_fns["o"].apply_prefill is False
    assert isinstance(m.model.layers[0].self_attn.q_proj, SparseLinear)
    assert not any(p.requires_grad for p in m.parameters())
    m.set_mlp_sparsity(0.4)
    assert m.model.layers[0].mlp.sparse_fns["down"].sparsity_level == 0.4 and m.model.layers[0].self_attn.sparse_fns["q"].sparsity_level == 0.5
    m.load_greedy_sparsities(str(tmp_path / "lookup"), 0.5)
    d_a1 = Distribution(str(tmp_path / "hist" / "layer-0" / "self_attn"), "h1")
    assert m.model.layers[0].self_attn.sparse_fns["v"].threshold == d_a1.icdf(0.5 + 0.2 / 2)
    m.reset_sparsities()
    assert all(l.mlp.sparse_fns[p].threshold == 0.0 for l in m.model.layers for p in ("gate", "up", "down"))
    # one copy of the weight: SparseLinear has no parameter or buffer of its own besides the wrapped Linear's
    sl = m.model.layers[0].mlp.sparse_linears["gate"]
    assert [n for n, _ in sl.named_parameters()] == ["linear.weight"] and not list(sl.buffers())
