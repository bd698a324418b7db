"""Copies the reference's two Triton kernel files next to the oracle (oracle/_ref/, git-ignored, travels to the GPU
box with gpurun) so that tools/bench_ref_triton.py can time the reference's own GPU kernel on the same B200 and
tests/test_ref_triton_gpu.py can cross-check parity against it.  Never committed, never imported by teal_b200/."""
import os, shutil, sys
SRC = os.environ.get("TEAL_REFERENCE", "/root/reference")
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
def main():
    if not os.path.isdir(SRC):
        print("reference not present; nothing fetched"); return 0
    os.makedirs(os.path.join(DST, "kernels"), exist_ok=True)
    for f in ("sparse_gemv.py", "compile_wrapper.py", "__init__.py"):
        shutil.copy(os.path.join(SRC, "kernels", f), os.path.join(DST, "kernels", f))
    print("fetched reference kernels into", DST); return 0
if __name__ == "__main__":
    sys.exit(main())
