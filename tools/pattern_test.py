"""Does the random-row gather pattern itself cost DRAM efficiency?  Same kernel, same bytes, different survivor patterns."""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import teal_b200.ops as ops
from tools.bench_gemv import time_graph

for (Z, N) in [(4096, 14336), (4096, 4096)]:
    ncopies = max(2, min(48, (768 << 20) // (Z * N * 2)))
    Ws = [(torch.randn(Z, N, device="cuda", dtype=torch.float16) * 0.02).T for _ in range(ncopies)]
    pats = {}
    g = torch.Generator(device="cuda").manual_seed(1)
    xr = torch.randn(1, 1, Z, device="cuda", dtype=torch.float16, generator=g)
    pats["random50"] = (xr, float(xr.abs().float().flatten().kthvalue(Z // 2).values))
    xb = torch.zeros(1, 1, Z, device="cuda", dtype=torch.float16); xb[..., : Z // 2] = 1.0
    pats["firsthalf"] = (xb, 0.5)
    xe = torch.zeros(1, 1, Z, device="cuda", dtype=torch.float16); xe[..., ::2] = 1.0
    pats["every2nd"] = (xe, 0.5)
    xq = torch.zeros(1, 1, Z, device="cuda", dtype=torch.float16); xq.view(-1, 8)[::2] = 1.0
    pats["blocks8"] = (xq, 0.5)
    pats["dense"] = (xr, -1.0)
    for cfg in (0, 18):
        ops.set_config(cfg)
        for name, (x, thr) in pats.items():
            nnz = Z if thr < 0 else int((x.abs() > thr).sum())
            fns = [(lambda i=i: ops.splitk_sparse_gemv(x, Ws[i % ncopies], thr, 0)) for i in range(48)]
            avg, best = time_graph(fns)
            alg = nnz * N * 2
            print(f"{Z}x{N} cfg {cfg} {name:10s} nnz {nnz:5d}: {avg:7.2f} us  {alg/avg/1e3:7.0f} GB/s", flush=True)
    ops.set_config(0)
    del Ws
