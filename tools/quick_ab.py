"""Smallest possible hardware A/B of the opt-in row-chunk generations against the validated default (one short GPU call):
Llama-3-8B-shaped decode at 50 %, tokens/s per dispatch setting + a dense-mode logits check against the default kernels.
Prints one JSON line per setting as it goes (partial output survives a timeout).

    python tools/quick_ab.py [--deadline 11] [--steps 40]
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

TUNES = [dict(generation=5), dict(generation=5, fhfma=0), dict(v1_opts=9), dict(generation=2)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--deadline", type=float, default=11.0, help="do not start another setting after this many seconds")
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--workload", default="llama-3-8b")
    ap.add_argument("--out", default="gpurun_out/r01_quick_ab.jsonl")
    ap.add_argument("--configs", default="", help="comma-separated teal_set_config ids: tokens/s with every GEMV launch forced to that kernel variant")
    ap.add_argument("--only", default="all", choices=["all", "tunes", "chain", "configs"],
                    help="the chained variants spin on flags: a protocol bug on hardware traps and poisons the CUDA context, so "
                         "run them in their own process (tunes first, chain second)")
    args = ap.parse_args()
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    out = open(args.out, "a")

    def emit(rec):
        rec["t"] = round(time.time() - T0, 2)
        line = json.dumps(rec)
        print(line, flush=True)
        out.write(line + "\n")
        out.flush()

    dec = SparseDecoder(ModelArgs.from_name(args.workload), max_seq=128, out_cap=128)
    dec.calibrate(0.5, n_tokens=2)
    emit({"setup": args.workload})

    def run(tune):
        ops.set_tuning(**tune) if tune else ops.set_tuning()
        rec = {"tuning": tune}
        try:
            dec.reset(1)
            with torch.cuda.stream(dec.stream):
                dec._enqueue_step("dense_own")
            dec.stream.synchronize()
            rec["dense_logits"] = dec.logits.float().reshape(-1).clone()
            dec.capture("sparse")
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
            rec["tok_s"] = round(1e3 * args.steps / e0.elapsed_time(e1), 1)
        except Exception as e:  # noqa: BLE001
            rec["error"] = repr(e)[:200]
        finally:
            ops.set_tuning()
        return rec

    def run_chain(mlp, attn, norms=False):
        dec.chain_mlp, dec.chain_attn, dec.chain_norms = mlp, attn, norms
        try:
            return dict(run({}), tuning={"chain_mlp": int(mlp), "chain_attn": int(attn), "chain_norms": int(norms)})
        finally:
            dec.chain_mlp = dec.chain_attn = dec.chain_norms = False

    base = run({})
    base_logits = base.pop("dense_logits", None)
    emit(base)
    for mlp, attn, norms in (() if args.only in ("tunes", "configs") else ((True, False, False), (False, True, False), (True, True, False), (True, True, True))):
        rec = run_chain(mlp, attn, norms)
        lg = rec.pop("dense_logits", None)
        if lg is not None and base_logits is not None:
            rec["dense_logits_max_abs_diff_vs_default"] = round(float((lg - base_logits).abs().max()), 4)
        emit(rec)
    for cfg in ([int(c) for c in args.configs.split(",")] if args.configs else []):
        ops.set_config(cfg)
        try:
            rec = run({})
        finally:
            ops.set_config(0)
        rec["config"] = cfg
        lg = rec.pop("dense_logits", None)
        if lg is not None and base_logits is not None:
            rec["dense_logits_max_abs_diff_vs_default"] = round(float((lg - base_logits).abs().max()), 4)
        emit(rec)
    for tune in ([] if args.only in ("chain", "configs") else TUNES):
        if time.time() - T0 > args.deadline:
            emit({"skipped": tune})
            continue
        rec = run(tune)
        lg = rec.pop("dense_logits", None)
        if lg is not None and base_logits is not None:
            rec["dense_logits_max_abs_diff_vs_default"] = round(float((lg - base_logits).abs().max()), 4)
            rec["dense_logits_absmax"] = round(float(base_logits.abs().max()), 3)
        emit(rec)


if __name__ == "__main__":
    main()
