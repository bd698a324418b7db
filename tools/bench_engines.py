"""A/B of the two decode engines on one GPU: launch-per-projection CUDA graph vs the one-kernel step program,
plus a sweep of the step program's tuning knobs.  Prints one JSON line per configuration (CUDA events on the replay
stream, same thresholds, same positions 0..warm+steps for every configuration).

    python tools/bench_engines.py [--workload llama-3-8b] [--steps 96] [--quick] [--out gpurun_out/x.jsonl]
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

import torch  # noqa: E402

from teal_b200.model import ModelArgs, SparseDecoder  # noqa: E402

SWEEP2 = [dict(opts=1), dict(opts=1, poll_ns=200), dict(opts=1, poll_ns=1000), dict(poll_ns=-1), dict(opts=1, poll_ns=-1),
          dict(opts=1, wide_melems=1), dict(opts=1, wide_melems=1, ctas_per_sm=6), dict(opts=1, wide_melems=1, rows_large=64),
          dict(opts=1, wide_melems=1, rows_large=64, ctas_per_sm=6), dict(opts=1, wide_melems=1, poll_ns=200, ctas_per_sm=6),
          dict(opts=1, wide_melems=1, rows_large=64, group_gate_up=4), dict(opts=1, wide_melems=1, head_rows=64)]
# round-2 candidates (variants written after round 1's GPU budget was spent)
SWEEP3 = [dict(warps=2), dict(warps=2, rows_large=64), dict(warps=2, rows_large=64, wide_melems=1), dict(warps=2, opts=2),
          dict(warps=2, opts=2, stagger_ns=100), dict(warps=2, opts=2, stagger_ns=300), dict(opts=2), dict(opts=2, stagger_ns=100),
          dict(opts=2, stagger_ns=300), dict(warps=2, rows_qkv=32, rows_o=32), dict(warps=2, rows_gate_up=128, rows_down=128),
          dict(warps=2, opts=3, head_rows=64), dict(warps=2, group_gate_up=4), dict(warps=2, group_gate_up=16),
          dict(warps=2, ctas_per_sm=12), dict(warps=2, ctas_per_sm=10)]
# multi-tile items (x loaded once per K chunk, next tile's first rows prefetched into L2 before the reduction)
SWEEP4 = [dict(tiles_per_item=2), dict(tiles_per_item=4), dict(tiles_per_item=2, opts=4), dict(tiles_per_item=4, opts=4),
          dict(tiles_per_item=2, warps=2, rows_small=64, rows_large=64), dict(tiles_per_item=4, warps=2, rows_small=64, rows_large=64),
          dict(tiles_per_item=4, wide_melems=1), dict(tiles_per_item=2, wide_melems=1, opts=4),
          dict(tiles_per_item=4, group_gate_up=16, group_qkv=8)]
GRAPH_TUNES = [dict(v1_opts=1), dict(v1_opts=2), dict(v1_opts=4), dict(v1_opts=5), dict(v1_opts=7), dict(rows_wide=256), dict(rows_wide=96), dict(rows_narrow=128), dict(rows_narrow=32), dict(rows_wide=256, rows_narrow=128),
               dict(generation=2), dict(generation=2, fhfma=0), dict(generation=2, fhfma=0, rows_wide=256), dict(generation=2, rows_wide=256, rows_narrow=128),
               dict(generation=2, spec_rows=2), dict(generation=2, spec_rows=4), dict(generation=2, spec_rows=8),
               dict(generation=2, spec_rows=8, spec_mode=2), dict(generation=2, spec_rows=16, spec_mode=2), dict(generation=3),
               dict(generation=3, spec_rows=4), dict(generation=3, spec_rows=8, spec_mode=2)]
SWEEP = [dict(ctas_per_sm=6), dict(ctas_per_sm=5), dict(ctas_per_sm=4), dict(rows_small=128), dict(rows_small=32),
         dict(rows_large=256), dict(rows_large=64), dict(group_gate_up=2), dict(group_gate_up=16), dict(group_gate_up=56),
         dict(group_qkv=1), dict(group_qkv=24), dict(wide_melems=1), dict(wide_melems=1000), dict(head_rows=16),
         dict(head_rows=64), dict(head_rows=128), dict(rows_small=128, rows_large=256), dict(rows_small=32, rows_large=64)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="llama-3-8b")
    ap.add_argument("--steps", type=int, default=96)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--sparsity", type=float, default=0.5)
    ap.add_argument("--quick", action="store_true", help="no knob sweep")
    ap.add_argument("--sweep", type=int, default=1, help="1: structural knobs, 2: polling / caching knobs, 3: 2-warp CTAs / scale-once / stagger, 4: multi-tile items")
    ap.add_argument("--check", action="store_true", help="teacher-forced logits of the two engines side by side first")
    ap.add_argument("--out", default=None)
    ap.add_argument("--graph-tunes", action="store_true",
                    help="also time the launch-per-projection graph under every opt-in dispatch setting (row-chunk generations 2 / 3, "
                         "speculative L2 prefetch rows) -- the round-2 A/B of kernels written without GPU access")
    args = ap.parse_args()
    t0 = time.time()
    margs = ModelArgs.from_name(args.workload)
    dec = SparseDecoder(margs, max_seq=256, out_cap=256)
    dec.calibrate(args.sparsity, n_tokens=4)
    out = open(args.out, "a") if args.out else None

    def emit(rec):
        rec["t"] = round(time.time() - t0, 1)
        line = json.dumps(rec)
        print(line, flush=True)
        if out:
            out.write(line + "\n")
            out.flush()

    def timed(mode, engine, **tuning):
        dec.capture(mode, engine=engine, **tuning)
        dec.reset(1)
        with torch.cuda.stream(dec.stream):
            for _ in range(args.warmup):
                dec.graph.replay()
            dec.stream.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(dec.stream)
            for _ in range(args.steps):
                dec.graph.replay()
            e1.record(dec.stream)
            dec.stream.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        toks = dec.out_tokens[: args.warmup + args.steps].cpu().tolist()
        return ms, toks

    emit({"workload": args.workload, "sparsity": args.sparsity, "steps": args.steps, "setup_s": round(time.time() - t0, 1)})
    if args.check:  # same tokens, same positions through both engines (eager launches), logits compared
        toks = [1, 17, 4242, 99]
        for mode in ("sparse", "dense_own"):
            dec.build_program(mode)
            logs = {}
            for engine in ("graph", "program"):
                dec.reset(toks[0])
                logs[engine] = []
                for t in toks:
                    dec.token.fill_(t)
                    if engine == "program":
                        dec._enqueue_program()
                    else:
                        dec._enqueue_step(mode)
                    torch.cuda.synchronize()
                    logs[engine].append((dec.logits.float().reshape(-1).clone(), int(dec.token.item()), dec.x.float().reshape(-1).clone()))
            for i, (a, b) in enumerate(zip(logs["graph"], logs["program"])):
                d = (a[0] - b[0]).abs()
                top2 = torch.topk(a[0], 2).values
                emit({"check": mode, "pos": i, "logit_absmax": round(float(a[0].abs().max()), 4), "max_abs_diff": float(d.max()),
                      "mean_abs_diff": float(d.mean()), "n_diff": int((d > 0).sum()), "top2_gap": float(top2[0] - top2[1]),
                      "argmax": [a[1], b[1]], "x_max_abs_diff": float((a[2] - b[2]).abs().max()),
                      "x_absmax": float(a[2].abs().max())})
    base_tok = None
    for mode, engine, tuning in [("sparse", "graph", {}), ("sparse", "program", {}), ("dense_own", "graph", {}),
                                 ("dense_own", "program", {})] + ([] if args.quick else [("sparse", "program", s) for s in {1: SWEEP, 2: SWEEP2, 3: SWEEP3, 4: SWEEP4}[args.sweep]]):
        try:
            ms, toks = timed(mode, engine, **tuning)
            if mode == "sparse" and engine == "graph":
                base_tok = toks
            agree = None
            if mode == "sparse" and base_tok is not None:
                agree = sum(int(a == b) for a, b in zip(toks, base_tok))
            emit({"mode": mode, "engine": engine, "tuning": tuning, "ms_per_step": round(ms, 4), "tok_s": round(1e3 / ms, 1),
                  "kernels_per_step": dec.kernels_per_step, "tokens_equal_to_graph_engine": agree})
        except Exception as e:  # noqa: BLE001 - keep the remaining configurations
            emit({"mode": mode, "engine": engine, "tuning": tuning, "error": repr(e)[:300]})


    if args.graph_tunes:
        import teal_b200.ops as ops
        for tune in GRAPH_TUNES:
            for mode in ("sparse", "dense_own"):
                try:
                    ops.set_tuning(**tune)
                    ms, toks = timed(mode, "graph")
                    agree = sum(int(a == b) for a, b in zip(toks, base_tok)) if (mode == "sparse" and base_tok) else None
                    emit({"mode": mode, "engine": "graph", "dispatch_tuning": tune, "ms_per_step": round(ms, 4), "tok_s": round(1e3 / ms, 1),
                          "tokens_equal_to_graph_engine": agree})
                except Exception as e:  # noqa: BLE001
                    emit({"mode": mode, "engine": "graph", "dispatch_tuning": tune, "error": repr(e)[:300]})
                finally:
                    ops.set_tuning()


if __name__ == "__main__":
    main()
