"""Engine-vs-engine numerics at real shapes: the launch-per-projection engine run twice (its own run-to-run noise: fp32
RED order) and against the step program, for 1-, 2- and full-depth models.  A random-init model with thresholding is a
chaotic map (a lane within one ulp of its threshold flips), so full-depth differences only mean something next to the
graph engine's own run-to-run difference; the shallow models localise real bugs.

    python tools/check_engines.py [--workload llama-3-8b] [--out gpurun_out/x.jsonl]
"""
import argparse
import dataclasses
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

import torch  # noqa: E402

from teal_b200.model import ModelArgs, SparseDecoder  # noqa: E402


def run(dec, engine, mode, toks):
    dec.reset(toks[0])
    res = []
    for t in toks:
        dec.token.fill_(t)
        if engine == "program":
            dec._enqueue_program()
        else:
            dec._enqueue_step(mode)
        torch.cuda.synchronize()
        res.append((dec.logits.float().reshape(-1).clone(), dec.x.float().reshape(-1).clone(), int(dec.token.item())))
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="llama-3-8b")
    ap.add_argument("--depths", default="1,2,32")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    out = open(args.out, "a") if args.out else None
    toks = [1, 17, 4242]
    for depth in [int(d) for d in args.depths.split(",")]:
        margs = dataclasses.replace(ModelArgs.from_name(args.workload), n_layer=depth)
        dec = SparseDecoder(margs, max_seq=64, out_cap=64)
        dec.calibrate(0.5, n_tokens=4)
        for mode in ("sparse", "dense_own"):
            dec.build_program(mode)
            a = run(dec, "graph", mode, toks)
            b = run(dec, "graph", mode, toks)
            c = run(dec, "program", mode, toks)
            d = run(dec, "program", mode, toks)
            for i in range(len(toks)):
                rec = {"depth": depth, "mode": mode, "pos": i, "x_absmax": round(float(a[i][1].abs().max()), 3),
                       "logit_absmax": round(float(a[i][0].abs().max()), 3)}
                for name, (p, q) in {"graph_vs_graph": (a, b), "graph_vs_program": (a, c), "program_vs_program": (c, d)}.items():
                    rec[name] = {"x": round(float((p[i][1] - q[i][1]).abs().max()), 5),
                                 "logits": round(float((p[i][0] - q[i][0]).abs().max()), 5),
                                 "argmax_equal": p[i][2] == q[i][2]}
                line = json.dumps(rec)
                print(line, flush=True)
                if out:
                    out.write(line + "\n")
                    out.flush()
        del dec
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
