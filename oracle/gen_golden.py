"""Generate tests/golden/*.npz by running the REFERENCE's own Python code on CPU.

Run once in the build container (where /root/reference exists):

    python oracle/gen_golden.py

The fixtures are committed; tests never need /root/reference.  Everything stored here is an
output of reference code (file:line given per block) on seeded inputs, or a small piece of
the reference's shipped calibration data (two histograms, two greedy-lookup tables) needed as
the *input* of those known-answer tests.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn

REF = os.environ.get("TEAL_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def _import_reference():
    sys.path.insert(0, REF)
    sys.path.insert(0, os.path.join(REF, "gpt-fast"))
    from utils.utils import SparsifyFn, get_layer_greedy_sparsities        # noqa
    from utils.utils import Distribution as UDistribution                   # noqa
    import distribution as gdist                                           # gpt-fast/distribution.py
    import model as gmodel                                                  # gpt-fast/model.py
    from teal.mlp import _mlp_forward                                       # noqa
    return SparsifyFn, get_layer_greedy_sparsities, UDistribution, gdist, gmodel, _mlp_forward


def seeded(shape, seed, scale=1.0, dist="normal"):
    rng = np.random.RandomState(seed)
    if dist == "normal":
        a = rng.standard_normal(size=shape)
    else:  # laplace-like (h2-type activations, notebooks/distributions.ipynb)
        a = rng.laplace(size=shape)
    return (a * scale).astype(np.float32)


def gen_sampling():
    """7. sampling (gpt-fast/generate.py:49-66): the reference's `sample()` on fp16 logits, the Exponential(1) noise
    it drew (same torch seed replayed), the token it returned and its probs."""
    sys.path.insert(0, REF)
    sys.path.insert(0, os.path.join(REF, "gpt-fast"))
    import generate as G
    cases = {}
    specs = [("v4096_k200_t08", 4096, 200, 0.8, 3.0), ("v512_k50_t10", 512, 50, 1.0, 2.0), ("v1000_k5_t05", 1000, 5, 0.5, 1.0),
             ("v96_none_t08", 96, None, 0.8, 2.0), ("v300_k1000_t13", 300, 1000, 1.3, 4.0), ("v2048_k200_ties", 2048, 200, 0.8, 0.05),
             ("v640_k64_t0", 640, 64, 0.0, 2.0)]
    for si, (name, V, k, T, scale) in enumerate(specs):
        for rep in range(3):
            logits = torch.from_numpy(seeded((1, 1, V), 900 + 10 * si + rep, scale)).to(torch.float16)
            if "ties" in name:   # coarse logits: many exact ties around the pivot, +/-0.0 among them
                logits = (logits * 8).round() / 8
                logits[0, 0, 5] = -0.0
            torch.manual_seed(1234 + rep)
            idx, probs = G.sample(logits, temperature=T, top_k=k)
            torch.manual_seed(1234 + rep)
            q = torch.empty_like(probs).exponential_(1)
            key = f"{name}__{rep}"
            cases[key + "__logits"] = logits.reshape(-1).numpy()
            cases[key + "__q"] = q.numpy()
            cases[key + "__probs"] = probs.numpy()
            cases[key + "__token"] = np.array(int(idx))
            cases[key + "__T"] = np.array(T, dtype=np.float64)
            cases[key + "__k"] = np.array(-1 if k is None else k)
    np.savez_compressed(os.path.join(OUT, "sampling.npz"), **cases)
    print("wrote sampling.npz:", len(specs) * 3, "cases")


def gen_histogram():
    """8. calibration histograms (utils/utils.py:125-185): the reference's ActivationModule fed with synthetic prefill
    activations.  The reference hard-codes `.to('cuda')` for its sort; this container has no GPU, so for the duration
    of the call `Tensor.to('cuda')` is redirected to the CPU (the arithmetic -- sort, order statistics, linspace,
    torch.histogram -- is the reference's own, unchanged)."""
    sys.path.insert(0, REF)
    from utils.utils import ActivationModule
    real_to = torch.Tensor.to

    def to_cpu_when_cuda(self, *a, **k):
        if a and a[0] == "cuda":
            return self
        return real_to(self, *a, **k)

    blob = {}
    for name, dist, scale, ntok, width in [("gauss", "normal", 0.03, 48, 256), ("laplace", "laplace", 0.5, 33, 704)]:
        am = ActivationModule("/tmp/unused")
        chunks = [torch.from_numpy(seeded((1, ntok, width), 700 + j, scale, dist)).half() for j in range(3)]
        for c in chunks:
            am.grab_activations(c, "h1")
        am.grab_activations(chunks[0][:, :1], "h1")   # a decode-shaped call: ignored by the reference
        torch.Tensor.to = to_cpu_when_cuda
        torch.cuda.empty_cache = lambda: None
        try:
            h = am.find_histogram(num_bins=2000 if name == "gauss" else 10000)
        finally:
            torch.Tensor.to = real_to
        blob[f"{name}__acts"] = torch.cat([c.squeeze(0) for c in chunks]).numpy()
        blob[f"{name}__counts"] = h["h1"].numpy()
        blob[f"{name}__centers"] = h["h1_centers"].numpy()
        blob[f"{name}__bins"] = np.array(2000 if name == "gauss" else 10000)
    np.savez_compressed(os.path.join(OUT, "histogram.npz"), **blob)
    print("wrote histogram.npz")


def main():
    os.makedirs(OUT, exist_ok=True)
    if "--only" in sys.argv:
        return {"sampling": gen_sampling, "histogram": gen_histogram}[sys.argv[sys.argv.index("--only") + 1]]()
    SparsifyFn, ref_greedy, UDistribution, gdist, gmodel, _mlp_forward = _import_reference()
    torch.manual_seed(0)

    # ---------------------------------------------------------------- 1. icdf known answers
    # gpt-fast/distribution.py:48-66 and utils/utils.py:105-123 (the two copies must agree)
    hist_cases = [("Llama-2-7B", 0, "mlp"), ("Llama-2-7B", 0, "self_attn"), ("Mistral-7B", 15, "mlp")]
    qs = [0.0, 0.001, 0.3, 0.5, 0.5 + 0.05, 0.5 + 0.2, 0.75, 0.8, 0.95, 0.999, 1.0, 1.05, 3.7]
    blob = {}
    for model, layer, sub in hist_cases:
        path = os.path.join(REF, "models", model, "histograms", f"layer-{layer}", sub)
        tag = f"{model}_L{layer}_{sub}".replace("-", "_").replace(".", "_")
        h = torch.load(os.path.join(path, "histograms.pt"))
        for ht in ("h1", "h2"):
            blob[f"{tag}__{ht}_counts"] = h[ht].numpy()
            blob[f"{tag}__{ht}_centers"] = h[f"{ht}_centers"].numpy()
            dg = gdist.Distribution(path, ht)
            du = UDistribution(path, ht)
            vals_g = np.array([float(dg.icdf(q)) for q in qs], dtype=np.float64)
            vals_u = np.array([float(du.icdf(q)) for q in qs], dtype=np.float64)
            assert np.array_equal(vals_g, vals_u), "reference's two Distribution copies disagree"
            blob[f"{tag}__{ht}_icdf"] = vals_g
            # SparsifyFn.set_threshold (utils/utils.py:28-30) incl. the s == 0 special case
            fn = SparsifyFn(du)
            ss = [0.0, 0.25, 0.4, 0.5, 0.6, 0.9]
            thr = []
            for s in ss:
                fn.set_threshold(s)
                thr.append(float(fn.threshold))
            blob[f"{tag}__{ht}_hf_thr"] = np.array(thr, dtype=np.float64)
            blob[f"{tag}__{ht}_hf_s"] = np.array(ss, dtype=np.float64)
    blob["qs"] = np.array(qs, dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, "icdf_kat.npz"), **blob)

    # ---------------------------------------------------------------- 2. threshold-then-matmul
    # SparsifyFn.forward (utils/utils.py:32-52) followed by nn.Linear, as teal/mlp.py:50-56 does.
    du = UDistribution(os.path.join(REF, "models", "Llama-2-7B", "histograms", "layer-0", "mlp"), "h1")
    cases = {}
    specs = [  # name, Z, N, x scale, x dist, threshold (None => s=0.5 from the histogram)
        ("gauss_z256_n384", 256, 384, 1.0, "normal", 0.6745),
        ("laplace_z320_n192", 320, 192, 0.05, "laplace", 0.03),
        ("hist_thr_z512_n256", 512, 256, 0.0302, "normal", None),
        ("thr0_z128_n64", 128, 64, 1.0, "normal", 0.0),
        ("all_dropped_z128_n64", 128, 64, 1.0, "normal", 100.0),
        ("ragged_z200_n72", 200, 72, 1.0, "normal", 0.5),
    ]
    for i, (name, Z, N, xs, xd, thr) in enumerate(specs):
        x = torch.from_numpy(seeded((1, 1, Z), 100 + i, xs, xd)).to(torch.float16)
        W = torch.from_numpy(seeded((N, Z), 200 + i, 0.02)).to(torch.float16)
        fn = SparsifyFn(du)
        if thr is None:
            fn.set_threshold(0.5)
        else:
            fn.threshold = thr
        lin = nn.Linear(Z, N, bias=False, dtype=torch.float16)
        lin.weight.data = W
        with torch.no_grad():
            xm = fn(x)                      # reference mask*x
            y16 = lin(xm)                   # reference fp16 CPU matmul
            y32 = nn.functional.linear(xm.float(), W.float())
        cases[f"{name}__x"] = x.numpy()
        cases[f"{name}__W"] = W.numpy()
        cases[f"{name}__thr"] = np.float64(fn.threshold)
        cases[f"{name}__xm"] = xm.numpy()
        cases[f"{name}__y16"] = y16.numpy()
        cases[f"{name}__y32"] = y32.numpy()
    # tie case: |x| exactly equal to fp16(threshold) must be DROPPED (strict >, fp16-rounded thr)
    t = 0.02038193
    t16 = float(torch.tensor(t, dtype=torch.float16))
    x = torch.tensor([[[t16, -t16, np.nextafter(np.float16(t16), np.float16(1)), 0.0, -0.0, 1.0, -1.0, 6e-8]]],
                     dtype=torch.float16)
    fn = SparsifyFn(du)
    fn.threshold = t
    cases["tie__x"] = x.numpy()
    cases["tie__thr"] = np.float64(t)
    cases["tie__xm"] = fn(x).numpy()
    # prefill half-sequence rule utils/utils.py:36-43
    xp = torch.from_numpy(seeded((1, 7, 64), 300)).to(torch.float16)
    fn.threshold = 0.7
    cases["prefill__x"] = xp.numpy()
    cases["prefill__thr"] = np.float64(0.7)
    cases["prefill__xm"] = fn(xp).numpy()
    fn.apply_prefill = False
    cases["prefill_off__xm"] = fn(xp).numpy()
    np.savez_compressed(os.path.join(OUT, "sparsify_linear.npz"), **cases)

    # ---------------------------------------------------------------- 3. HF MLP forward (teal/mlp.py:37-57)
    class FakeMLP(nn.Module):
        def __init__(self, Z, I):
            super().__init__()
            self.gate_proj = nn.Linear(Z, I, bias=False, dtype=torch.float16)
            self.up_proj = nn.Linear(Z, I, bias=False, dtype=torch.float16)
            self.down_proj = nn.Linear(I, Z, bias=False, dtype=torch.float16)
            self.act_fn = nn.SiLU()
            self.grabbing_mode = False
    Z, I = 128, 352
    mlp = FakeMLP(Z, I)
    mlp.gate_proj.weight.data = torch.from_numpy(seeded((I, Z), 401, 0.08)).to(torch.float16)
    mlp.up_proj.weight.data = torch.from_numpy(seeded((I, Z), 402, 0.08)).to(torch.float16)
    mlp.down_proj.weight.data = torch.from_numpy(seeded((Z, I), 403, 0.08)).to(torch.float16)
    fns = {k: SparsifyFn(du) for k in ("gate", "up", "down")}
    fns["gate"].threshold, fns["up"].threshold, fns["down"].threshold = 0.6, 0.4, 0.05
    mlp.sparse_fns = nn.ModuleDict(fns)
    x = torch.from_numpy(seeded((1, 1, Z), 404)).to(torch.float16)
    with torch.no_grad():
        y = _mlp_forward(mlp, x)
    np.savez_compressed(os.path.join(OUT, "hf_mlp.npz"), x=x.numpy(),
                        w_gate=mlp.gate_proj.weight.data.numpy(), w_up=mlp.up_proj.weight.data.numpy(),
                        w_down=mlp.down_proj.weight.data.numpy(), thr=np.array([0.6, 0.4, 0.05]),
                        y=y.numpy())

    # ---------------------------------------------------------------- 4. greedy lookup (utils/utils.py:243-258)
    g = {}
    import pandas as pd
    for model in ("Mistral-7B",):
        rdir = os.path.join(REF, "models", model, "lookup")
        for layer in (0, 1):
            df = pd.read_csv(os.path.join(rdir, f"layer-{layer}", "results.csv"))
            g[f"layer{layer}_table"] = df.to_numpy(dtype=np.float64)
            g["columns"] = np.array(list(df.columns))
        for target in (0.25, 0.4, 0.5, 0.9):
            res = ref_greedy([target, target], rdir)
            g[f"target_{target}"] = np.array([[res[p][l] for p in ["q", "k", "v", "o", "gate", "up", "down"]]
                                              for l in (0, 1)], dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, "greedy_lookup.npz"), **g)

    # ---------------------------------------------------------------- 5. gpt-fast dense model forward
    # gpt-fast/model.py:99-143 (prefill of 5 tokens then 3 decode steps, dense old_forward path)
    cfg = dict(block_size=64, vocab_size=96, n_layer=2, n_head=4, dim=64, n_local_heads=2,
               intermediate_size=160, rope_base=10000)
    args = gmodel.ModelArgs(**cfg)
    m = gmodel.Transformer(args)
    sd = {}
    for i, (k, v) in enumerate(m.state_dict().items()):
        if "norm" in k:
            w = 1.0 + 0.1 * seeded(tuple(v.shape), 500 + i)
        else:
            w = seeded(tuple(v.shape), 500 + i, 0.08)
        sd[k] = torch.from_numpy(w).to(torch.float16)
    m.load_state_dict(sd)
    m = m.to(torch.float16).eval()
    with torch.no_grad():
        m.setup_caches(max_batch_size=1, max_seq_length=16)
        toks = torch.tensor([[3, 17, 42, 8, 77]])
        outs = [m(toks, torch.arange(5)).numpy()]
        nxt = [11, 60, 5]
        for j, t in enumerate(nxt):
            outs.append(m(torch.tensor([[t]]), torch.tensor([5 + j])).numpy())
    blob = {f"w__{k}": v.numpy() for k, v in sd.items()}
    blob["prefill_logits"] = outs[0]
    for j in range(3):
        blob[f"decode_logits_{j}"] = outs[1 + j]
    blob["prompt"] = toks.numpy()
    blob["next"] = np.array(nxt)
    blob["cfg_keys"] = np.array(list(cfg.keys()))
    blob["cfg_vals"] = np.array([float(v) for v in cfg.values()])
    np.savez_compressed(os.path.join(OUT, "gptfast_dense.npz"), **blob)

    # ---------------------------------------------------------------- 6. op schemas (compile_wrapper.py:151-162)
    from kernels.sparse_gemv import SparseGEMV, SparseQKVGEMV
    with open(os.path.join(OUT, "op_schemas.txt"), "w") as f:
        f.write("sparse_gemv" + SparseGEMV.schematize() + "\n")
        f.write("sparse_qkv_gemv" + SparseQKVGEMV.schematize() + "\n")
    gen_sampling()
    gen_histogram()
    print("wrote", sorted(os.listdir(OUT)))


if __name__ == "__main__":
    main()
