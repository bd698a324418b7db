"""Per-node time of trivial kernels in a captured graph chain (launch floor), with and without PDL, and the
effect of PDL on the sparse GEMV chain."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import teal_b200.ops as ops
from teal_b200 import _lib
from tools.bench_gemv import time_graph
lib = _lib.load()
am = torch.zeros(1, dtype=torch.int64, device="cuda"); tok = torch.zeros(1, dtype=torch.int32, device="cuda")
pos = torch.zeros(1, dtype=torch.int32, device="cuda"); out = torch.zeros(4096, dtype=torch.int32, device="cuda")
def adv():
    lib.teal_advance(am.data_ptr(), tok.data_ptr(), pos.data_ptr(), out.data_ptr(), 0, torch.cuda.current_stream().cuda_stream)
for pdl in (0, 1):
    lib.teal_set_pdl(pdl)
    avg, _ = time_graph([adv] * 64)
    print(f"trivial 1-CTA kernel chain, pdl={pdl}: {avg:.2f} us per node")
Z, N = 4096, 4096
Ws = [(torch.randn(Z, N, device="cuda", dtype=torch.float16) * 0.02).T for _ in range(24)]
x = torch.randn(1, 1, Z, device="cuda", dtype=torch.float16)
thr = float(x.abs().float().flatten().kthvalue(Z // 2).values)
for cfg in (30, 32, 14):
    ops.set_config(cfg)
    for pdl in (0, 1):
        lib.teal_set_pdl(pdl)
        avg, _ = time_graph([(lambda i=i: ops.splitk_sparse_gemv(x, Ws[i % 24], thr, 0)) for i in range(48)])
        print(f"4096x4096 cfg {cfg} pdl={pdl}: {avg:.2f} us")
    # all-dropped: pure fixed cost of the kernel (no weight traffic)
    lib.teal_set_pdl(0)
    avg, _ = time_graph([(lambda i=i: ops.splitk_sparse_gemv(x, Ws[i % 24], 100.0, 0)) for i in range(48)])
    print(f"4096x4096 cfg {cfg} all-dropped (no gather): {avg:.2f} us")
lib.teal_set_pdl(0); ops.set_config(0)
