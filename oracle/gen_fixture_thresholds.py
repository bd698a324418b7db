"""Thresholds of the reference's SHIPPED calibration fixtures, for the GPU box (which has no /root/reference):
models/<model>/histograms[/Untitled]/layer-i/{mlp,self_attn}/histograms.pt -> Distribution.icdf(0.5 + s/2) per projection
(gpt-fast/generate.py:266-287), through teal_b200.distribution (bit-identical to the reference's Distribution, see
tests/test_host_logic.py).  Writes tests/golden/fixture_thresholds.json; tools/fixture_run.py feeds them to the decoder."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from teal_b200.distribution import PROJS, layer_thresholds, resolve_histogram_dir  # noqa: E402

REF = os.environ.get("TEAL_REFERENCE", "/root/reference")
MODELS = {"llama-3-8b": ("Llama-3-8B", 32), "llama-2-7b": ("Llama-2-7B", 32), "mistral-7b": ("Mistral-7B", 32)}


def main():
    out = {}
    for key, (name, n_layer) in MODELS.items():
        hist = resolve_histogram_dir(os.path.join(REF, "models", name))
        out[key] = {"source": os.path.relpath(hist, REF), "sparsity": {}}
        for s in (0.4, 0.5):
            out[key]["sparsity"][str(s)] = [layer_thresholds(hist, i, {p: s for p in PROJS}) for i in range(n_layer)]
    path = os.path.join(ROOT, "tests", "golden", "fixture_thresholds.json")
    with open(path, "w") as f:
        json.dump(out, f)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
