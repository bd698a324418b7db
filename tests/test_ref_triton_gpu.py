"""On-GPU cross-check against the REFERENCE's own Triton kernels, run unmodified from oracle/_ref/ (copied there by
oracle/fetch_ref.py; git-ignored, travels with the gpurun snapshot).  Skipped when the copy or Triton is unavailable.
The reference accumulates with fp16 global atomics (kernels/sparse_gemv.py:83), so it is the noisier side: we bound
|ours - ref| by the reference's own error against the fp32 truth plus our stated tolerance."""
import os
import sys

import pytest
import torch

from oracle import teal_oracle as O

pytestmark = pytest.mark.gpu
REF = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref")


def _ref_kernels():
    if not os.path.isdir(os.path.join(REF, "kernels")):
        pytest.skip("oracle/_ref not fetched")
    if REF not in sys.path:
        sys.path.insert(0, REF)
    try:
        from kernels.sparse_gemv import qkv_gemv, splitk_sparse_gemv
    except Exception as e:  # triton / msgspec missing or incompatible on this box
        pytest.skip(f"reference kernels not importable: {e}")
    return splitk_sparse_gemv, qkv_gemv


@pytest.mark.parametrize("Z,N", [(4096, 4096), (4096, 11008)])
def test_matches_reference_triton_kernel(Z, N):
    ref_gemv, _ = _ref_kernels()
    import teal_b200.ops as ops
    g = torch.Generator().manual_seed(Z + N)
    x = torch.randn(1, 1, Z, generator=g).half()
    W = (torch.randn(N, Z, generator=g) * 0.02).half()
    thr = float(x.abs().float().flatten().kthvalue(Z // 2).values)
    Wc = ops.to_column_major(W.cuda())
    try:
        yr = ref_gemv(x.cuda(), Wc, thr, 10).float().cpu().reshape(-1)
    except Exception as e:
        pytest.skip(f"reference Triton kernel failed to run here: {type(e).__name__}")
    ym = ops.splitk_sparse_gemv(x.cuda(), Wc, thr, 0).float().cpu().reshape(-1)
    truth = O.sparse_linear_f32(x, W, thr).reshape(-1)
    ref_err = float((yr - truth).abs().max())
    our_err = float((ym - truth).abs().max())
    assert our_err <= ref_err + 1e-3          # we are at least as close to the truth as the reference's GPU kernel
    assert float((ym - yr).abs().max()) <= ref_err + our_err + 1e-3
