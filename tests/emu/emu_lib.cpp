// libteal_emu: the C ABI of include/teal_b200.h served by the CPU emulation of the kernels (see cuda_emu.h).
// Built from the production sources api.cu, dispatch.cu, rowchunk_gemv.cu, decode_kernels.cu, step_program.cu compiled
// as C++ with -DTEAL_EMU; this file adds what the emulation leaves out and the emulator's own controls.
// TEST INFRASTRUCTURE ONLY: pointers are host pointers, "streams" are ignored, every call is synchronous.
#include "cuda_emu.h"

#include "../../teal_b200/csrc/common.cuh"

namespace teal {
// The TMA-bulk-ring and cp.async-ring kernels are hardware-validated design points built on mbarrier / cp.async
// machinery that the emulation does not model.
int launch_gemv(GemvParams&, void*, size_t, int, cudaStream_t) { return set_error(TEAL_ERR_UNSUPPORTED, "emulation: TMA kernel not modelled"); }
int launch_gather_gemv(GemvParams&, void*, size_t, int, cudaStream_t, int) {
  return set_error(TEAL_ERR_UNSUPPORTED, "emulation: cp.async kernel not modelled");
}
int launch_simple(const __half*, const __half*, __half*, int, int, long long, float, cudaStream_t) {
  return set_error(TEAL_ERR_UNSUPPORTED, "emulation: simple kernel not modelled");
}
}  // namespace teal

extern "C" {
int teal_emu_set_sm_count(int n) { emu::set_sm_count(n); return 0; }
int teal_emu_set_concurrent(int on) { emu::set_concurrent(on != 0); return 0; }
}
