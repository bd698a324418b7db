"""GPU comparator: the reference's own Triton kernels (oracle/_ref/kernels/sparse_gemv.py, fetched by
oracle/fetch_ref.py, unmodified) vs libteal_b200 on the same B200, same inputs, same timing method.
Also cross-checks parity between the two GPU implementations."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.path.join(ROOT, "oracle", "_ref")
if not os.path.isdir(os.path.join(REF, "kernels")):
    print(json.dumps({"ref_triton": "unavailable (oracle/_ref not fetched)"})); sys.exit(0)
sys.path.insert(0, REF)
import teal_b200.ops as ops
from tools.bench_gemv import time_graph, PEAK
try:
    from kernels.sparse_gemv import splitk_sparse_gemv as ref_gemv, qkv_gemv as ref_qkv
except Exception as e:
    print(json.dumps({"ref_triton": f"import failed: {e}"})); sys.exit(0)

out = open(os.path.join(ROOT, "gpurun_out", "ref_triton.jsonl"), "a")
for shp in sys.argv[1].split(",") if len(sys.argv) > 1 else ["4096x14336", "14336x4096", "4096x4096", "4096x6144"]:
    Z, N = (int(v) for v in shp.split("x"))
    per = Z * N * 2
    ncopies = max(2, min(32, (768 << 20) // per))
    Ws = [(torch.randn(Z, N, device="cuda", dtype=torch.float16) * 0.02).T for _ in range(ncopies)]
    x = torch.randn(1, 1, Z, device="cuda", dtype=torch.float16)
    thr = float(x.abs().float().flatten().kthvalue(Z // 2).values)
    nnz = int((x.abs() > torch.tensor(thr, dtype=torch.float16, device="cuda")).sum())
    alg = nnz * N * 2 + Z * 2 + N * 2
    try:
        for i in range(3):
            yr = ref_gemv(x, Ws[0], thr, 10)  # autotune warm-up
        ym = ops.splitk_sparse_gemv(x, Ws[0], thr, 0)
        diff = float((yr.float() - ym.float()).abs().max())
        fr = [(lambda i=i: ref_gemv(x, Ws[i % ncopies], thr, 10)) for i in range(32)]
        ref_us, _ = time_graph(fr)
    except Exception as e:
        print(json.dumps({"shape": shp, "ref_triton": f"failed: {type(e).__name__}: {e}"[:300]})); continue
    fm = [(lambda i=i: ops.splitk_sparse_gemv(x, Ws[i % ncopies], thr, 0)) for i in range(32)]
    my_us, _ = time_graph(fm)
    rec = dict(shape=shp, nnz=nnz, ref_triton_us=round(ref_us, 2), teal_b200_us=round(my_us, 2), speedup=round(ref_us / my_us, 2),
               ref_gbs=round(alg / ref_us / 1e3), ours_gbs=round(alg / my_us / 1e3), ours_frac_peak=round(alg / my_us / 1e3 / PEAK, 3),
               max_abs_diff_vs_ref_kernel=diff)
    print(json.dumps(rec), flush=True)
    out.write(json.dumps(rec) + "\n")
    del Ws
