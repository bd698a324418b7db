"""Decode with the thresholds of the reference's SHIPPED calibration fixtures (gpt-fast/generate.py:266-287 over
models/Llama-3-8B/histograms/Untitled, precomputed into tests/golden/fixture_thresholds.json by
oracle/gen_fixture_thresholds.py) on the random-init model of the same shape.  The histograms describe the REAL model's
activations, so on random weights the kept fraction is not the nominal 50 %: this tool reports the measured
kept fraction per projection kind and the resulting tokens/s next to the synthetic-calibrated run.

    python tools/fixture_run.py [--workload llama-3-8b] [--sparsity 0.5] [--out gpurun_out/fixture_run.jsonl]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from teal_b200.model import ModelArgs, SparseDecoder  # noqa: E402

NAMES = {"llama-3-8b": "llama-3-8b", "llama-2-7b": "7B", "mistral-7b": "Mistral-7B"}


def tok_s(dec, steps=96):
    dec.capture("sparse")
    dec.reset(1)
    with torch.cuda.stream(dec.stream):
        for _ in range(4):
            dec.graph.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(dec.stream)
        for _ in range(steps):
            dec.graph.replay()
        e1.record(dec.stream)
        dec.stream.synchronize()
    return 1e3 * steps / e0.elapsed_time(e1)


def kept(dec):
    """Measured kept fraction of one step: gathered / dense weight bytes (the kernels' own probe)."""
    probe = torch.zeros(2, dtype=torch.int64, device=dec.dev)
    dec.lib.teal_set_bytes_probe(probe.data_ptr())
    dec.reset(1)
    for _ in range(4):
        dec._enqueue_step("sparse")
    torch.cuda.synchronize()
    dec.lib.teal_set_bytes_probe(None)
    g, d = (int(v) for v in probe.tolist())
    return g / max(1, d)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="llama-3-8b", choices=sorted(NAMES))
    ap.add_argument("--sparsity", default="0.5")
    ap.add_argument("--out", default="gpurun_out/fixture_run.jsonl")
    args = ap.parse_args()
    fx = json.load(open(os.path.join(ROOT, "tests", "golden", "fixture_thresholds.json")))[args.workload]
    thr = fx["sparsity"][args.sparsity]
    dec = SparseDecoder(ModelArgs.from_name(NAMES[args.workload]), max_seq=128, out_cap=128)
    dec.set_thresholds(thr)
    rec = {"workload": args.workload, "nominal_sparsity": float(args.sparsity), "thresholds_from": fx["source"],
           "kept_fraction_with_shipped_thresholds": round(kept(dec), 4), "tok_s_with_shipped_thresholds": round(tok_s(dec), 1)}
    dec.calibrate(float(args.sparsity), n_tokens=8)
    rec["kept_fraction_synthetic_calibration"] = round(kept(dec), 4)
    rec["tok_s_synthetic_calibration"] = round(tok_s(dec), 1)
    rec["note"] = ("shipped histograms describe the real checkpoint's activations; on random-init weights of the same shape the "
                   "kept fraction differs from nominal (reported, not corrected)")
    print(json.dumps(rec))
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "a") as f:
        f.write(json.dumps(rec) + "\n")


if __name__ == "__main__":
    main()
