#!/usr/bin/env python
"""bench.py -- decode tok/s, batch 1, Llama-3-8B-shaped random-init model at 50 % activation sparsity.

One "step" = one decoded token through the whole hot path (128 fused sparse-GEMV launches +
attention + lm_head + arg-max), replayed from one captured CUDA graph.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload llama-3-8b] [--sparsity 0.5]
    python bench.py --impl reference ...      # the reference's own CPU path (SparsifyFn.apply + F.linear), whole tokens

Prints ONE JSON line (see README / DESIGN.md section 7 for the fields).  For N > 1 launch with
torchrun; the model is tensor-parallel (gpt-fast/tp.py topology) => "scaling": "strong".
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {"llama-3-8b": "llama-3-8b", "llama-2-7b": "7B", "mistral-7b": "Mistral-7B", "llama-2-70b": "70B",
             "llama-3-70b": "llama-3-70b", "tiny": "tiny"}


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = float(r[1])
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# --------------------------------------------------------------------------------------------- CPU arm
def _reference_sparsify():
    """The reference's own SparsifyFn (utils/utils.py:9-55) from oracle/_ref (fetched by oracle/fetch_ref.py, git-ignored,
    travels with gpurun) -> kind "reference"; else the oracle port of the same two lines -> kind "port"."""
    ref = os.path.join(ROOT, "oracle", "_ref")
    if os.path.exists(os.path.join(ref, "utils", "utils.py")):
        try:
            sys.path.insert(0, ref)
            from utils.utils import SparsifyFn  # the reference's class, unmodified

            def make(thr):
                fn = SparsifyFn(None, init_threshold=thr)
                fn.threshold = thr  # what set_threshold() leaves (utils.py:28-30)
                return fn.apply     # x.abs().gt(threshold) * x (utils.py:51-52)
            return make, "reference"
        except Exception:
            pass
        finally:
            if ref in sys.path:
                sys.path.remove(ref)
    from oracle import teal_oracle as O
    return (lambda thr: (lambda x: O.sparsify_apply(x, thr))), "port"


def cpu_reference(args_model, sparsity, steps, warmup=1, budget_s=None, threads=None):
    """The reference's CPU path for this workload, timed for real: every step is ONE whole token -- n_layer layer passes of
    SparsifyFn.apply + F.linear per projection (utils/utils.py:51-52: it masks, it does not skip rows) over a rotating pool of
    4 distinct layer weight sets (1.7 GB fp16 for the 8B shape: nothing stays in the CPU caches) + the dense lm_head.
    `steps` x ms_per_step is the wall time of the timed region.  With budget_s the number of steps is cut so that the timed
    region stays near the budget (cpu_baseline leg).  All host threads (torchrun sets OMP_NUM_THREADS=1: undone here)."""
    import torch
    import torch.nn.functional as F
    torch.set_num_threads(threads or os.cpu_count() or 1)
    make, kind = _reference_sparsify()
    a = args_model
    kv = a.n_local_heads * a.head_dim
    g = torch.Generator().manual_seed(0)
    shapes = {"qkv": (a.dim + 2 * kv, a.dim), "o": (a.dim, a.dim), "gate": (a.intermediate_size, a.dim),
              "up": (a.intermediate_size, a.dim), "down": (a.dim, a.intermediate_size)}
    npool = 4
    pool = [{k: (torch.randn(n, z, generator=g) * 0.02).half() for k, (n, z) in shapes.items()} for _ in range(npool)]
    lm = (torch.randn(a.vocab_size, a.dim, generator=g) * 0.02).half()
    x0 = torch.randn(1, 1, a.dim, generator=g).half()

    def quantile_thr(v):
        return float(v.abs().float().flatten().kthvalue(max(1, int(v.numel() * sparsity))).values)

    def norm(v):
        vf = v.float()
        return (vf * torch.rsqrt(vf.pow(2).mean(-1, keepdim=True) + a.norm_eps)).half()

    fns = {}

    def sp(site, v):  # thresholds are fixed on first use (calibration), like the reference's set_threshold
        if site not in fns:
            fns[site] = make(quantile_thr(v))
        return fns[site](v)

    def token():
        x = x0
        for li in range(a.n_layer):
            W, s = pool[li % npool], li % npool
            qkv = F.linear(sp((s, "qkv"), norm(x)), W["qkv"])
            att = qkv[..., : a.dim]  # (attention itself is not part of the threshold+matmul path; context is short)
            x = x + F.linear(sp((s, "o"), att), W["o"])
            xn = norm(x)
            h = F.silu(F.linear(sp((s, "gate"), xn), W["gate"])) * F.linear(sp((s, "up"), xn), W["up"])
            x = x + F.linear(sp((s, "down"), h), W["down"])
        return F.linear(norm(x), lm)

    for _ in range(max(1, warmup)):
        t0 = time.perf_counter()
        token()
        one = time.perf_counter() - t0
    if budget_s is not None:
        steps = max(1, min(steps, int(budget_s / max(one, 1e-3))))
    t0 = time.perf_counter()
    for _ in range(steps):
        token()
    sec = (time.perf_counter() - t0) / steps
    return {"value": 1.0 / sec, "unit": "tok/s", "cores": torch.get_num_threads(), "kind": kind,
            "sample": f"{steps} whole tokens, each {a.n_layer} layer passes (qkv,o,gate,up,down: SparsifyFn.apply + F.linear, fp16) "
                      f"over a rotating pool of {npool} layer weight sets + lm_head; wall time / steps; the reference CPU path "
                      f"masks, then runs the dense matmul"}, sec, steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=128)
    ap.add_argument("--warmup", type=int, default=16)
    ap.add_argument("--impl", default="teal_b200", choices=["teal_b200", "reference"])
    ap.add_argument("--workload", default="llama-3-8b", choices=sorted(WORKLOADS))
    ap.add_argument("--sparsity", type=float, default=0.5)
    ap.add_argument("--profile", default="uniform", choices=["uniform", "greedy-means"],
                    help="greedy-means: per-projection sparsities = the means of models/Mistral-7B/lookup at 0.5 "
                         "(SURVEY.md section 8d, config 5): non-uniform thresholds, 3-threshold QKV, gate != up")
    ap.add_argument("--no-dense", action="store_true", help="skip the dense baselines")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extra", action="store_true", help="skip the secondary legs (reference-default sampling, 70B TP leg)")
    ap.add_argument("--config", type=int, default=0, help="teal_set_config id (kernel variant / K chunk; 0 = by launch size)")
    ap.add_argument("--repeats", type=int, default=3, help="timed repetitions of the K steps (value = median, spread reported)")
    ap.add_argument("--top-k", type=int, default=0, help="with --temperature: sample like the reference's generate.py default "
                    "(--temperature 0.8 --top-k 200) instead of greedy arg-max in the HEADLINE number")
    ap.add_argument("--temperature", type=float, default=0.0)
    ap.add_argument("--no-70b", action="store_true", help="skip the Llama-2-70B leg (tp_70b object; north_star's TP target, 137 GB / N per GPU)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    from teal_b200.model import ModelArgs
    margs = ModelArgs.from_name(WORKLOADS[args.workload])
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    metric = f"decode tok/s batch=1 {args.workload} @{int(args.sparsity * 100)}% sparsity"
    config = {"workload": f"{args.workload} full decode step (random-init fp16, synthetic-calibrated thresholds)",
              "sparsity": args.sparsity, "context": "positions 0..steps (short context, generate.py:171)",
              "sampling": "greedy", "parallelism": f"tp{world}" if world > 1 else "single",
              "l2": "weights (>= 8 GB per GPU) far exceed the 126 MB L2; every step streams them from HBM",
              "engine": "graph"}

    # ------------------------------------------------------------------ reference arm: the reference's CPU path
    if args.impl == "reference":
        if rank != 0:
            return
        base, sec, steps = cpu_reference(margs, args.sparsity, args.steps, warmup=min(args.warmup, 2))
        line = {"impl": "reference", "metric": metric, "value": base["value"], "unit": "tok/s", "n_gpus": args.gpus,
                "steps": steps, "warmup": min(args.warmup, 2), "ms_per_step": sec * 1e3, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f16", "data": "synthetic", "config": config,
                "cpu_baseline": base,
                "e2e": {"value": base["value"], "unit": "tok/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return

    # ------------------------------------------------------------------ our arm
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the decode path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    pg = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        import datetime
        dist.init_process_group("nccl", rank=rank, world_size=world, timeout=datetime.timedelta(seconds=180))
        pg = dist.group.WORLD
    from teal_b200 import _lib, ops
    from teal_b200.model import SparseDecoder
    ops.set_config(args.config)
    if args.config:
        config["kernel_config"] = args.config
    K, Wm = args.steps, args.warmup
    max_seq = max(256, K + Wm + 16)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def build(margs_, sparsity):
        dec_ = SparseDecoder(margs_, device=f"cuda:{local_rank}", max_seq=max_seq, rank=rank, world=world, process_group=pg,
                             out_cap=max_seq)
        if args.profile == "greedy-means":
            sp = {"q": 0.589, "k": 0.469, "v": 0.300, "o": 0.566, "gate": 0.474, "up": 0.417, "down": 0.582}
            config["sparsity_profile"] = sp
            thr_ = dec_.calibrate(sp, n_tokens=8)
        else:
            thr_ = dec_.calibrate(sparsity, n_tokens=8)
        return dec_, thr_

    def timed(dec_, mode, steps, warmup, e2e=False):
        """K graph replays, CUDA events on the replay stream (max over ranks); e2e: the public-API loop with the token
        H2D from pinned memory and the sampled token D2H + sync every step."""
        if dec_.graph is None or dec_.graph_mode != mode:
            dec_.capture(mode)
        dec_.reset(1)
        host_in = torch.ones(1, dtype=torch.int32).pin_memory()
        host_out = torch.zeros(1, dtype=torch.int32).pin_memory()
        with torch.cuda.stream(dec_.stream):
            for _ in range(warmup):
                dec_.graph.replay()
            dec_.stream.synchronize()
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t_wall = time.perf_counter()
            e0.record(dec_.stream)
            for _ in range(steps):
                if e2e:
                    dec_.token.copy_(host_in, non_blocking=True)
                    dec_.graph.replay()
                    host_out.copy_(dec_.token, non_blocking=True)
                    dec_.stream.synchronize()
                    host_in.copy_(host_out)
                else:
                    dec_.graph.replay()
            e1.record(dec_.stream)
            dec_.stream.synchronize()
            wall = time.perf_counter() - t_wall
            barrier()
        ms = e0.elapsed_time(e1)
        if e2e:
            ms = max(ms, wall * 1e3)
        if world > 1:
            t = torch.tensor([ms], device=dec_.dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms / steps

    dec, thr = build(margs, args.sparsity)
    lib = dec.lib
    tp_exchange = None  # (reported next to `config`, not inside it: both arms must carry the same `config`)
    if world > 1:
        tp_exchange = ("fused into the row-parallel GEMV finalisers (push protocol over NVLink peer memory)" if getattr(dec, "tp_gemv", False)
                       else "one-shot peer-memory kernel" if dec.tp_fused else "nccl all-reduce")
    if args.temperature > 0:
        dec.set_sampling(args.temperature, args.top_k, seed=0)
        config["sampling"] = f"top_k={args.top_k} temperature={args.temperature} (generate.py:49-66 on device)"

    probe = torch.zeros(2, dtype=torch.int64, device=dec.dev)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ncu_window = os.environ.get("TEAL_NCU_WINDOW") == "1"  # ncu --profile-from-start off: profile only this region
    if ncu_window:
        torch.cuda.profiler.start()
    reps = [timed(dec, "sparse", K, Wm) for _ in range(max(1, args.repeats))]
    if ncu_window:
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
    clocks = sampler.stop() if rank == 0 else None
    ms_step = sorted(reps)[len(reps) // 2]
    ms_e2e = timed(dec, "sparse", K, Wm, e2e=True)
    kernels_per_step = dec.kernels_per_step
    launches = kernels_per_step * K

    # the reference's default sampling (generate.py:353-354: temperature 0.8, top_k 200) as a second reported number
    sampled = None
    if not args.no_extra and args.temperature <= 0:
        try:
            dec.set_sampling(0.8, 200, seed=0)
            sampled = {"tok_s": 1e3 / timed(dec, "sparse", max(16, K // 2), 4), "sampling": "top_k=200 temperature=0.8 (generate.py:49-66,353-354 on device)"}
        except Exception as e:  # noqa: BLE001
            sampled = {"error": repr(e)[:200]}
        dec.set_sampling(0.0)

    # measured kept fraction / algorithmic bytes of one step (eager, outside the timed region)
    lib.teal_set_bytes_probe(probe.data_ptr())
    dec.reset(1)
    for _ in range(4):
        dec._enqueue_step("sparse")
    torch.cuda.synchronize()
    lib.teal_set_bytes_probe(None)
    gathered, dense_eq = (int(v) / 4 for v in probe.tolist())
    layer_bytes, head_bytes = dec.weight_bytes()

    dense = {}
    if not args.no_dense:
        for mode in ("dense_own", "dense_cublas"):
            try:
                dense[mode] = 1e3 / timed(dec, mode, max(16, K // 2), 4)
            except Exception as e:  # keep the headline even if a baseline variant fails
                dense[mode] = None
                if rank == 0:
                    print(f"[bench] {mode} failed: {e}", file=sys.stderr)

    # ------------------------------------------------------------------ roofline of the dominant kernel
    # fused RMSNorm + gate/up gather-GEMV + SiLU*mul, timed alone with CUDA events on the launching stream,
    # rotating over all layers' weights (n_layer x ~117 MB >> L2) with the decoder's real activations.
    peak, peak_kind = measured_peak()
    a = margs
    rf = None
    # every rank replays 4 sparse steps (they contain the TP exchanges) to capture each layer's real MLP input
    dec.set_thresholds(thr)
    dec.reset(1)
    with torch.cuda.stream(dec.stream):
        for _ in range(3):
            dec._enqueue_step("sparse")
        dec.ffn_inputs = []
        dec._enqueue_step("sparse")
    dec.stream.synchronize()
    ffn_in, dec.ffn_inputs = dec.ffn_inputs, None
    barrier()
    if rank == 0:
        ws = dec.ws
        probe.zero_()
        t_l = [t for t in thr]
        nrep = max(1, 96 // a.n_layer)
        with torch.cuda.stream(dec.stream):
            gu_acc = dec.defer_acc.data_ptr() + 4 * dec.nqkv_l

            def gate_up_all():
                # the launch the decode step makes: the deferred-epilogue form (fp32 sums stay in L2 for the down projection,
                # SparseDecoder.deferred_epilogue).  Back-to-back launches keep adding into the same sums -- timing only; the
                # buffer is cleared below.
                for li, Wl in enumerate(dec.layers):
                    if dec.deferred_epilogue:
                        _lib.check(lib.teal_norm_sparse_gate_up_acc_f16(ffn_in[li][0].data_ptr(), Wl["ffn_norm"].data_ptr(), a.norm_eps,
                                                                        ffn_in[li][1].data_ptr(), Wl["w1"].data_ptr(), Wl["w3"].data_ptr(), gu_acc,
                                                                        a.dim, dec.inter_l, Wl["w1"].stride(1), Wl["w3"].stride(1),
                                                                        t_l[li]["gate"], t_l[li]["up"], dec.stream.cuda_stream), "gate_up probe")
                        continue
                    lib.teal_norm_sparse_gate_up_f16(ffn_in[li][0].data_ptr(), Wl["ffn_norm"].data_ptr(), a.norm_eps,
                                                     ffn_in[li][1].data_ptr(), Wl["w1"].data_ptr(), Wl["w3"].data_ptr(), dec.h.data_ptr(), None,
                                                     a.dim, dec.inter_l, Wl["w1"].stride(1), Wl["w3"].stride(1),
                                                     t_l[li]["gate"], t_l[li]["up"], ws.data_ptr(), ws.numel(),
                                                     dec.stream.cuda_stream)
            gate_up_all()
            g = torch.cuda.CUDAGraph()
            lib.teal_set_pdl(1)  # launched as in the decode step's graph: programmatic dependent launch, each launch still
            try:                 # waits (griddepcontrol.wait) for the previous grid to complete before it reads anything
                with torch.cuda.graph(g, stream=dec.stream):
                    gate_up_all()
            finally:
                lib.teal_set_pdl(0)
            lib.teal_set_bytes_probe(probe.data_ptr())
            gate_up_all()
            dec.stream.synchronize()
            lib.teal_set_bytes_probe(None)
            for _ in range(2):
                g.replay()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(dec.stream)
            for _ in range(nrep):
                g.replay()
            e1.record(dec.stream)
            dec.stream.synchronize()
            dec.defer_acc.zero_()
            dec.stream.synchronize()
        n_l = nrep * a.n_layer
        us = e0.elapsed_time(e1) * 1e3 / n_l
        alg = int(probe[0]) / a.n_layer + a.dim * 2 + (0 if dec.deferred_epilogue else dec.inter_l * 2)  # gathered rows + x (+ h) per launch
        # dram__bytes_read + dram__bytes_write of this kernel from the committed ncu --set full capture (per launch):
        # a STATIC record of that capture, only reported while the capture is of the kernel that ran here
        kname = "sparse_gemv_kernel<1,8,2>"
        traffic, traffic_src = None, None
        try:
            with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
                rec = json.load(f).get(args.workload, {})
            if world == 1 and rec.get("kernel") == kname and not args.config:
                traffic, traffic_src = rec.get("traffic_bytes"), f"static: {rec.get('source')}"
        except Exception:
            pass
        rf = {"bound": "hbm", "kernel": f"{kname} (fused RMSNorm + gate/up gather-GEMV" + (", fp32 sums left in L2 for the down projection: deferred epilogue)" if dec.deferred_epilogue else " + SiLU*mul)"), "achieved": alg / us / 1e3,
              "peak": peak, "unit": "GB/s", "frac": alg / us / 1e3 / peak, "traffic": traffic, "traffic_source": traffic_src,
              "peak_kind": f"of {peak_kind}", "us_per_launch": us, "alg_bytes_per_launch": alg,
              "kept_fraction": float(probe[0]) / max(1.0, float(probe[1])),
              "timing": "CUDA events on the launching stream, graph of n_layer PDL launches (as in the decode step), each layer on its own weights and its own real decode input (L2 defeated)"}

    # ------------------------------------------------------------------ Llama-2-70B leg (north_star's TP target)
    # The same step on the 70B shapes at this world size (137 GB / N of weights per GPU), reported next to the headline
    # metric, never instead of it: {tok_s, ms_per_step, dense_tok_s}.  Skipped when the GPU cannot hold the shard.
    tp70 = None
    if not args.no_70b and not args.no_extra and args.workload != "llama-2-70b":
        import gc
        headline_kernels = kernels_per_step
        try:
            g = gate_up_all = ffn_in = None  # (the roofline graph keeps the 8B weights alive)
        except Exception:
            pass
        dec.graph = None
        del dec
        gc.collect()
        torch.cuda.empty_cache()
        m70 = ModelArgs.from_name("70B")
        need = 2 * (m70.n_layer * (m70.dim * (m70.dim + 2 * m70.n_local_heads * m70.head_dim) + m70.dim * m70.dim
                                   + 3 * m70.dim * m70.intermediate_size) / world + 2 * m70.vocab_size * m70.dim) + (4 << 30)
        free = torch.tensor([torch.cuda.mem_get_info()[0]], device=f"cuda:{local_rank}", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(free, op=dist.ReduceOp.MIN)
        if float(free) < need:
            tp70 = {"skipped": f"needs {need / 2**30:.0f} GiB per GPU, {float(free) / 2**30:.0f} GiB free"}
        else:
            try:
                d70, _ = build(m70, args.sparsity)
                k70 = max(16, K // 4)
                ms70 = timed(d70, "sparse", k70, 3)
                dn70 = None if args.no_dense else 1e3 / timed(d70, "dense_own", k70, 3)
                tp70 = {"workload": "llama-2-70b full decode step", "tok_s": 1e3 / ms70, "ms_per_step": ms70, "dense_tok_s": dn70,
                        "speedup_vs_dense": (1e3 / ms70 / dn70) if dn70 else None, "steps": k70, "n_gpus": world}
                del d70
            except Exception as e:  # noqa: BLE001 -- the headline line must survive this leg
                tp70 = {"error": repr(e)[:200]}
        kernels_per_step = headline_kernels

    cpu = None
    if rank == 0 and not args.no_cpu:
        try:
            cpu, _, _ = cpu_reference(margs, args.sparsity, 16, warmup=1, budget_s=12.0)
        except Exception as e:
            cpu = {"value": None, "unit": "tok/s", "cores": None, "kind": "port", "sample": f"failed: {e}"}

    if rank == 0:
        tok_s = 1e3 / ms_step
        best_dense = max([v for v in dense.values() if v] or [0.0])
        step_bytes = gathered + head_bytes  # algorithmic weight bytes of one token on this rank
        line = {"metric": metric, "value": tok_s, "unit": "tok/s", "n_gpus": world, "steps": K, "warmup": Wm,
                "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f16", "data": "synthetic", "config": config,
                "e2e": {"value": 1e3 / ms_e2e, "unit": "tok/s", "h2d_bytes_per_step": 4, "d2h_bytes_per_step": 4},
                "gpu_launches": launches, "clocks": clocks, "roofline": rf, "cpu_baseline": cpu,
                "value_spread": {"tok_s_min": 1e3 / max(reps), "tok_s_max": 1e3 / min(reps), "repeats": len(reps)},
                "dense_tok_s": dense, "speedup_vs_dense": (tok_s / best_dense) if best_dense else None,
                "sampled": sampled, "tp_70b": tp70, "tp_exchange": tp_exchange,
                "kept_fraction": gathered / max(1.0, dense_eq),
                "step_weight_bytes": step_bytes, "step_hbm_frac": step_bytes / (ms_step * 1e-3) / 1e9 / peak,
                "kernels_per_step": kernels_per_step}
        print(json.dumps(line))
    if world > 1:
        # the captured graphs still reference the NCCL communicator: a clean destroy_process_group() can hang,
        # so every rank synchronises, flushes and leaves without tearing the communicator down
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
