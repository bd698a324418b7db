"""Smallest possible hardware A/B of kernel variants / K chunks inside the real decode step (one short GPU call):
Llama-3-8B-shaped decode at 50 %, tokens/s with every sparse GEMV launch forced to a teal_set_config id, plus a
dense-mode logits check against the default dispatch.  Prints one JSON line per setting as it goes.

    python tools/quick_ab.py --configs 1,17,33,65,3,35 [--steps 96]
"""
import argparse
import json
import os
import sys
import time

T0 = time.time()
import torch  # noqa: E402

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import teal_b200.ops as ops  # noqa: E402
from teal_b200.model import ModelArgs, SparseDecoder  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=96)
    ap.add_argument("--workload", default="llama-3-8b")
    ap.add_argument("--out", default="gpurun_out/quick_ab.jsonl")
    ap.add_argument("--configs", default="", help="comma-separated teal_set_config ids: (variant + 1) + 16 * (lanes per warp / 32)")
    ap.add_argument("--repeats", type=int, default=1)
    ap.add_argument("--ab-deferred", type=int, default=0, help="N alternating rounds: finaliser route / deferred epilogue (SparseDecoder.deferred_epilogue)")
    args = ap.parse_args()
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    out = open(args.out, "a")

    def emit(rec):
        rec["t"] = round(time.time() - T0, 2)
        line = json.dumps(rec)
        print(line, flush=True)
        out.write(line + "\n")
        out.flush()

    dec = SparseDecoder(ModelArgs.from_name(args.workload), max_seq=args.steps + 32, out_cap=args.steps + 32)
    dec.calibrate(0.5, n_tokens=2)
    emit({"setup": args.workload})

    def run(cfg):
        ops.set_config(cfg)
        rec = {"config": cfg}
        try:
            dec.reset(1)
            with torch.cuda.stream(dec.stream):
                dec._enqueue_step("dense_own")
            dec.stream.synchronize()
            rec["dense_logits"] = dec.logits.float().reshape(-1).clone()
            dec.capture("sparse")
            best = []
            for _ in range(args.repeats):
                dec.reset(1)
                with torch.cuda.stream(dec.stream):
                    for _ in range(4):
                        dec.graph.replay()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(dec.stream)
                    for _ in range(args.steps):
                        dec.graph.replay()
                    e1.record(dec.stream)
                    dec.stream.synchronize()
                best.append(round(1e3 * args.steps / e0.elapsed_time(e1), 1))
            rec["tok_s"] = max(best)
            rec["tok_s_all"] = best
        except Exception as e:  # noqa: BLE001
            rec["error"] = repr(e)[:200]
        finally:
            ops.set_config(0)
        return rec

    if args.ab_deferred:
        ref = None
        for r in range(args.ab_deferred):
            for deferred in (False, True):
                dec.deferred_epilogue = deferred
                rec = run(0)
                lg = rec.pop("dense_logits", None)
                ref = lg if ref is None else ref
                rec.update(deferred_epilogue=deferred, round=r, kernels_per_step=dec.kernels_per_step,
                           dense_logits_max_abs_diff_vs_first=round(float((lg - ref).abs().max()), 4))
                emit(rec)
        dec.deferred_epilogue = True
        return
    base = run(0)
    base_logits = base.pop("dense_logits", None)
    emit(base)
    for cfg in ([int(c) for c in args.configs.split(",")] if args.configs else []):
        rec = run(cfg)
        lg = rec.pop("dense_logits", None)
        if lg is not None and base_logits is not None:
            rec["dense_logits_max_abs_diff_vs_default"] = round(float((lg - base_logits).abs().max()), 4)
        emit(rec)


if __name__ == "__main__":
    main()
