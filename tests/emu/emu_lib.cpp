// libteal_emu: the C ABI of include/teal_b200.h served by the CPU emulation of the kernels (see cuda_emu.h).
// Built from the production sources api.cu, gemv.cu, decode_kernels.cu compiled as C++ with -DTEAL_EMU; this file adds
// the emulator's own controls.
// TEST INFRASTRUCTURE ONLY: pointers are host pointers, "streams" are ignored, every call is synchronous.
#include "cuda_emu.h"

#include "../../teal_b200/csrc/common.cuh"

extern "C" {
int teal_emu_set_sm_count(int n) { emu::set_sm_count(n); return 0; }
int teal_emu_set_concurrent(int on) { emu::set_concurrent(on != 0); return 0; }
int teal_emu_set_max_workers(int n) { emu::set_max_workers(n); return 0; }
}
