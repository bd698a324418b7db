"""Per-phase timeline of ONE token through the step program (teal_step_desc::debug_times): for every phase the time
its first ticket was taken, its first item saw its dependencies, its last item finished, and the summed busy time of
its items.  Prints one JSON line per phase plus per-kind averages over the layers.

    python tools/program_timeline.py [--workload llama-3-8b] [--out gpurun_out/x.jsonl] [knob=value ...]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

import torch  # noqa: E402

from teal_b200.model import ModelArgs, SparseDecoder  # noqa: E402

NAMES = ["qkv", "attn", "o", "gate_up", "down"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="llama-3-8b")
    ap.add_argument("--mode", default="sparse")
    ap.add_argument("--out", default=None)
    ap.add_argument("knobs", nargs="*")
    args = ap.parse_args()
    knobs = {k: int(v) for k, v in (kv.split("=") for kv in args.knobs)}
    margs = ModelArgs.from_name(args.workload)
    dec = SparseDecoder(margs, max_seq=64, out_cap=64)
    dec.calibrate(0.5, n_tokens=4)
    nph = 1 + 5 * margs.n_layer + 2
    dbg = torch.zeros(nph, 4, dtype=torch.int64, device=dec.dev)
    prog = dec.build_program(args.mode, debug_times=dbg.data_ptr(), **knobs)
    # items per phase from the plan (binary search over tickets)
    n = prog.n_items
    starts = []
    t = 0
    while t < n:
        ph = prog.item(t).phase
        starts.append(t)
        lo, hi = t, n
        while hi - lo > 1:
            mid = (lo + hi) // 2
            if prog.item(mid).phase == ph:
                lo = mid
            else:
                hi = mid
        t = hi
    sizes = [b - a for a, b in zip(starts, starts[1:] + [n])]
    assert len(sizes) == nph, (len(sizes), nph)
    dec.reset(1)
    for _ in range(6):
        dec._enqueue_program()
    torch.cuda.synchronize()
    init = torch.tensor([-1, -1, 0, 0], dtype=torch.int64, device=dec.dev).repeat(nph, 1)
    dbg.copy_(init)
    torch.cuda.synchronize()
    dec._enqueue_program()
    torch.cuda.synchronize()
    d = dbg.cpu().tolist()
    t0 = min(r[0] for r in d)
    out = open(args.out, "a") if args.out else None
    rows = []
    for i, r in enumerate(d):
        kind = "embed" if i == 0 else ("lm_head" if i == nph - 2 else ("advance" if i == nph - 1 else NAMES[(i - 1) % 5]))
        rec = {"phase": i, "kind": kind, "layer": (i - 1) // 5 if 0 < i < nph - 2 else -1, "items": sizes[i],
               "first_ticket_us": (r[0] - t0) / 1e3, "first_ready_us": (r[1] - t0) / 1e3, "last_end_us": (r[2] - t0) / 1e3,
               "busy_us_per_item": r[3] / 1e3 / sizes[i]}
        rows.append(rec)
        if out:
            out.write(json.dumps(rec) + "\n")
    # per-kind averages over layers 4..n-1 (steady state)
    summ = {"knobs": knobs, "mode": args.mode, "token_us": rows[-1]["last_end_us"]}
    for k, name in enumerate(NAMES):
        sel = [r for r in rows if r["kind"] == name and r["layer"] >= 4]
        prev = [rows[r["phase"] - 1] for r in sel]
        summ[name] = {
            "span_us": round(sum(r["last_end_us"] - r["first_ready_us"] for r in sel) / len(sel), 2),
            "gap_after_prev_end_us": round(sum(r["first_ready_us"] - p["last_end_us"] for r, p in zip(sel, prev)) / len(sel), 2),
            "ticket_to_ready_us": round(sum(r["first_ready_us"] - r["first_ticket_us"] for r in sel) / len(sel), 2),
            "busy_us_per_item": round(sum(r["busy_us_per_item"] for r in sel) / len(sel), 2), "items": sel[0]["items"]}
    h = rows[-2]
    summ["lm_head"] = {"span_us": round(h["last_end_us"] - h["first_ready_us"], 2),
                       "gap_after_prev_end_us": round(h["first_ready_us"] - rows[-3]["last_end_us"], 2),
                       "busy_us_per_item": round(h["busy_us_per_item"], 2), "items": h["items"]}
    line = json.dumps(summ)
    print(line, flush=True)
    for r in rows[1 + 5 * 10: 1 + 5 * 11]:
        print(json.dumps(r), flush=True)
    if out:
        out.write(line + "\n")


if __name__ == "__main__":
    main()
