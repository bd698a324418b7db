"""Where does a decode step's time go?  In-kernel timeline of the production sparse-GEMV kernel inside the real graph
replay (run on the GPU box).

Captures one decode step with the TIMED twin of the kernel (teal_set_debug_timeline: globaltimer stamps per
CTA, one slot per launch), replays it, and prints per projection kind -- averaged over the layers --
  * kernel span (first CTA start -> last CTA end), head (first CTA start -> first rows consumed anywhere),
  * the gap to the previous GEMV kernel (its last CTA end -> this kernel's first survivor set / first rows),
  * CTA lifetime distribution and its split: wait for producer | x + survivors | first DRAM rows | rest of the gather |
    reduce + RED | fence + arrival | (finaliser),
  * resident-CTA occupancy over time (how long HBM has < half / < quarter of the slots streaming: the tail),
  * waves: CTA start times in units of the median lifetime.
One JSON line per kind (+ one per launch with --per-launch).  Timed kernels carry a few extra stores: treat absolute
numbers as +-3 %, the structure is the point.

    python tools/gemv_timeline.py [--workload llama-3-8b] [--sparsity 0.5] [--out gpurun_out/gemv_timeline.jsonl]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from teal_b200.model import ModelArgs, SparseDecoder  # noqa: E402

KINDS = ("qkv", "o", "gate_up", "down")
SLOT_CTAS = 4096


def analyse(t):
    """t: [n_cta, 16] int64 stamps of one launch (rows with t[:,0] == 0 are unused)."""
    t = t[t[:, 0] > 0]
    n = len(t)
    start, waited, surv, first, gather, red, arrive, fin = (t[:, i].astype(np.float64) for i in range(8))
    end = np.where(fin > 0, fin, arrive)
    first = np.where(first > 0, first, gather)   # thread 0's warp had no survivor in its first lane group
    t0 = start.min()
    life = end - start
    q = lambda a: [round(float(np.percentile(a, p)) / 1e3, 2) for p in (10, 50, 90)]
    # occupancy over time
    grid = np.linspace(t0, end.max(), 400)
    active = ((start[None, :] <= grid[:, None]) & (end[None, :] > grid[:, None])).sum(1)
    streaming = ((first[None, :] <= grid[:, None]) & (gather[None, :] > grid[:, None])).sum(1)
    dt = (grid[1] - grid[0]) / 1e3
    peak = max(1, active.max())
    return {
        "ctas": n, "span_us": round((end.max() - t0) / 1e3, 2),
        "head_us": round((first.min() - t0) / 1e3, 2),                       # launch -> first weight rows consumed anywhere
        "first_wait_return_us": round((waited.min() - t0) / 1e3, 2),         # how long the first CTAs sat in griddepcontrol.wait
        "last_start_us": round((start.max() - t0) / 1e3, 2),                 # when the last CTA got a slot
        "tail_after_last_start_us": round((end.max() - start.max()) / 1e3, 2),
        "life_us_p10_50_90": q(life),
        "split_us_p50": {"wait_producer": q(waited - start)[1], "x_survivors": q(surv - waited)[1], "first_rows": q(first - surv)[1],
                         "rest_of_gather": q(gather - first)[1], "reduce_red": q(red - gather)[1], "fence_arrive": q(arrive - red)[1],
                         "finalise(last only)": q((fin - arrive)[fin > 0])[1] if (fin > 0).any() else None},
        "peak_resident_ctas": int(peak),
        "us_below_half_streaming": round(float((streaming < 0.5 * streaming.max()).sum() * dt), 2),
        "us_below_quarter_streaming": round(float((streaming < 0.25 * streaming.max()).sum() * dt), 2),
        "sms_used": int(len(np.unique(t[:, 8]))),
        "_t0": float(t0), "_end": float(end.max()), "_first_surv": float(surv.min()), "_first_rows": float(first.min()),
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="llama-3-8b")
    ap.add_argument("--sparsity", type=float, default=0.5)
    ap.add_argument("--mode", default="sparse")
    ap.add_argument("--per-launch", action="store_true")
    ap.add_argument("--out", default="gpurun_out/gemv_timeline.jsonl")
    args = ap.parse_args()
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    a = ModelArgs.from_name(args.workload)
    dec = SparseDecoder(a, max_seq=128, out_cap=128)
    dec.calibrate(args.sparsity, n_tokens=4)
    n_launch = 4 * a.n_layer
    # capture() runs one eager warm-up step before it captures: slots 0 .. n-1 go to the warm-up, n .. 2n-1 to the graph
    buf = torch.zeros(2 * n_launch * SLOT_CTAS * 16, dtype=torch.int64, device=dec.dev)
    dec.lib.teal_set_debug_timeline(buf.data_ptr(), 2 * n_launch)
    try:
        dec.capture(args.mode)
    finally:
        dec.lib.teal_set_debug_timeline(None, 0)
    dec.reset(1)
    for _ in range(6):
        buf.zero_()
        dec.graph.replay()
    torch.cuda.synchronize()
    stamps = buf.cpu().numpy().reshape(2 * n_launch, SLOT_CTAS, 16)[n_launch:]
    out = open(args.out, "a")
    per = [analyse(stamps[i]) for i in range(n_launch)]
    for i, r in enumerate(per):
        r["kind"], r["layer"] = KINDS[i % 4], i // 4
        if i > 0:
            r["gap_prev_end_to_first_survivors_us"] = round((r["_first_surv"] - per[i - 1]["_end"]) / 1e3, 2)
            r["gap_prev_end_to_first_rows_us"] = round((r["_first_rows"] - per[i - 1]["_end"]) / 1e3, 2)
            r["start_before_prev_end_us"] = round((per[i - 1]["_end"] - r["_t0"]) / 1e3, 2)   # PDL overlap
    for kind in KINDS:
        rows = [r for r in per if r["kind"] == kind and r["layer"] >= 2]
        agg = {"kind": kind, "layers_averaged": len(rows)}
        for key in ("ctas", "span_us", "head_us", "first_wait_return_us", "last_start_us", "tail_after_last_start_us", "peak_resident_ctas",
                    "us_below_half_streaming", "us_below_quarter_streaming", "gap_prev_end_to_first_survivors_us",
                    "gap_prev_end_to_first_rows_us", "start_before_prev_end_us"):
            agg[key] = round(float(np.mean([r[key] for r in rows if key in r])), 2)
        agg["life_us_p10_50_90"] = [round(float(np.mean([r["life_us_p10_50_90"][j] for r in rows])), 2) for j in range(3)]
        agg["split_us_p50"] = {k: (round(float(np.mean([r["split_us_p50"][k] for r in rows if r["split_us_p50"][k] is not None])), 2)
                                   if any(r["split_us_p50"][k] is not None for r in rows) else None) for k in rows[0]["split_us_p50"]}
        line = json.dumps(agg)
        print(line, flush=True)
        out.write(line + "\n")
    if args.per_launch:
        for r in per:
            out.write(json.dumps({k: v for k, v in r.items() if not k.startswith("_")}) + "\n")
    layer_span = np.mean([per[4 * l + 3]["_end"] - per[4 * l]["_t0"] for l in range(2, a.n_layer)]) / 1e3
    print(json.dumps({"layer_us_first_qkv_cta_to_last_down_cta": round(float(layer_span), 2)}))


if __name__ == "__main__":
    main()
