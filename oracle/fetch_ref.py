"""Copies the few reference files the GPU-side comparators need next to the oracle (oracle/_ref/, git-ignored, travels to
the GPU box with gpurun; /root/reference does not exist there):
  kernels/      the reference's Triton kernels    -> tools/bench_ref_triton.py, tests/test_ref_triton_gpu.py
  utils/utils.py (SparsifyFn)                      -> bench.py --impl reference / cpu_baseline (kind "reference")
  gpt-fast/{model,tp,distribution}.py              -> tests/test_ref_caller_gpu.py (the drop-in under the reference's own caller)
Never committed, never imported by teal_b200/.  Run by __graft_entry__.build() when the reference is present."""
import os, shutil, sys
SRC = os.environ.get("TEAL_REFERENCE", "/root/reference")
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
FILES = [("kernels", "sparse_gemv.py"), ("kernels", "compile_wrapper.py"), ("kernels", "__init__.py"),
         ("utils", "utils.py"), ("utils", "__init__.py"),
         ("gpt-fast", "model.py"), ("gpt-fast", "tp.py"), ("gpt-fast", "distribution.py")]
def main():
    if not os.path.isdir(SRC):
        print("reference not present; nothing fetched"); return 0
    for d, f in FILES:
        os.makedirs(os.path.join(DST, d), exist_ok=True)
        shutil.copy(os.path.join(SRC, d, f), os.path.join(DST, d, f))
    print("fetched reference files into", DST); return 0
if __name__ == "__main__":
    sys.exit(main())
