"""Generate tests/golden/greedyopt.npz by running the REFERENCE's greedy optimiser (teal/greedyopt.py:101-159,
`process_layer`, imported from /root/reference) on the small layer of oracle/greedy_fixture.py built with the reference's
own SparsifyFn and Distribution (utils/utils.py:9-55,72-123).  Run once in the build container:

    python oracle/gen_golden_greedy.py

The reference hard-codes `device = "cuda"`; this container has no GPU, so `.to("cuda")` is redirected to the CPU for the
duration of the call (as oracle/gen_golden.py does for the histogram builder) -- the arithmetic is the reference's.
Stored: the histograms (inputs), the layer input, the CSV the reference wrote, the allocation it returned, and the
reference's `f` / `calculate_activation_error` on seeded inputs."""
import os
import sys
import tempfile

import numpy as np
import torch

REF = os.environ.get("TEAL_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "..", "tests", "golden")
sys.path.insert(0, os.path.join(HERE, ".."))
sys.path.insert(0, REF)

from oracle import greedy_fixture as GF  # noqa: E402
from teal_b200.calibration import find_histogram  # noqa: E402  (bit-identical to the reference's builder: tests/golden/histogram.npz)


def main():
    import teal.greedyopt as G                       # the reference module
    from utils.utils import Distribution, SparsifyFn  # the reference classes
    blob = {}
    with tempfile.TemporaryDirectory() as tmp:
        # histograms of the dense layer's own site activations, in the reference's file layout
        dense = GF.Layer({p: torch.nn.Identity() for p in GF.PROJS}, seed=0)
        x = GF.inputs(seed=1)
        sites = GF.site_activations(dense, x)
        distr = {}
        for block in ("self_attn", "mlp"):
            d = os.path.join(tmp, "histograms", "layer-0", block)
            os.makedirs(d)
            hist = {}
            for ht in ("h1", "h2"):
                counts, centers = find_histogram(sites[block][ht], num_bins=400)
                hist[ht], hist[f"{ht}_centers"] = counts, centers
                blob[f"{block}__{ht}_counts"], blob[f"{block}__{ht}_centers"] = counts.numpy(), centers.numpy()
            torch.save(hist, os.path.join(d, "histograms.pt"))
            distr[block] = {ht: Distribution(d, ht) for ht in ("h1", "h2")}
        site = {"q": ("self_attn", "h1"), "k": ("self_attn", "h1"), "v": ("self_attn", "h1"), "o": ("self_attn", "h2"),
                "gate": ("mlp", "h1"), "up": ("mlp", "h1"), "down": ("mlp", "h2")}
        fns = {p: SparsifyFn(distr[site[p][0]][site[p][1]], init_sparsity=0.0) for p in GF.PROJS}
        layer = GF.Layer(fns, seed=0)
        os.makedirs(os.path.join(tmp, "activations"))
        torch.save(x, os.path.join(tmp, "activations", "act_0.pt"))

        real_to, real_mod_to = torch.Tensor.to, torch.nn.Module.to
        torch.Tensor.to = lambda self, *a, **k: self if (a and a[0] == "cuda") else real_to(self, *a, **k)
        torch.nn.Module.to = lambda self, *a, **k: self if (a and a[0] == "cuda") else real_mod_to(self, *a, **k)
        try:
            for tag, model_type, target, step in (("l3_8b", "Llama-3-8B", 0.5, 0.05), ("l2_7b", "Llama-2-7B", 0.35, 0.1)):
                with torch.no_grad():
                    final = G.process_layer(layer, model_type, 0, target, step, 0.25, tmp)
                blob[f"{tag}__csv"] = np.array(open(os.path.join(tmp, "lookup", "layer-0", "results.csv")).read())
                blob[f"{tag}__final"] = np.array([final[p] for p in GF.PROJS], dtype=np.float64)
                blob[f"{tag}__args"] = np.array([target, step, 0.25], dtype=np.float64)
        finally:
            torch.Tensor.to, torch.nn.Module.to = real_to, real_mod_to
    blob["x"] = x.numpy()
    # the two pure helpers on seeded inputs
    g = torch.Generator().manual_seed(5)
    a, b = torch.randn(2, 16, 8, generator=g), torch.randn(2, 16, 8, generator=g)
    blob["err_a"], blob["err_b"] = a.numpy(), b.numpy()
    blob["err_vals"] = np.array([float(G.calculate_activation_error(a, b, lf)) for lf in (0.25, 0.5, 1.0)], dtype=np.float64)
    sp = {"q": 0.3, "k": 0.9, "v": 0.1, "o": 0.45, "gate": 0.6, "up": 0.55, "down": 0.2}
    blob["f_vals"] = np.array([G.f(sp, G.weight_dict[m]) for m in sorted(G.weight_dict)], dtype=np.float64)
    blob["f_models"] = np.array(sorted(G.weight_dict))
    blob["weights"] = np.array([[G.weight_dict[m][p] for p in GF.PROJS] for m in sorted(G.weight_dict)], dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, "greedyopt.npz"), **blob)
    print("wrote greedyopt.npz;", blob["l3_8b__csv"].item().count("\n") - 1, "+", blob["l2_7b__csv"].item().count("\n") - 1, "greedy steps")


if __name__ == "__main__":
    main()
