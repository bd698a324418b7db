// Shared device helpers (inline PTX for sm_100a) and host-side plumbing of libteal_b200.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/teal_b200.h"

namespace teal {

constexpr int kMaxSeg = 4;         // column regions per call (q | k | v, gate | up)
constexpr int kMaxCounters = 4096; // column tiles per call
constexpr int kSsqParts = TEAL_SSQ_PARTS;        // partial sums of squares per replica (unused tail stays zero)
constexpr int kSsqReplicas = TEAL_SSQ_REPLICAS;  // identical copies, so that ~4000 consumer warps do not share one L2 line
constexpr int kMaxTpWorld = 16;    // ranks of one tensor-parallel group (one NVSwitch domain)
constexpr int kListCap = 512;      // survivor-list entries (= input lanes) one warp may own

enum { kModePlain = 0, kModePairSilu = 1, kModeSurvivors = 2 };
enum { kInPlain = 0, kInRmsNorm = 1, kInGateUpAcc = 2 };  // how x is produced from the input vector
// what happens to the split-K sums: a tile finaliser writes fp16 (plain / + residual), or -- kOutDeferred -- nobody in this
// launch does: the fp32 sums stay in the caller's buffer and the CONSUMER kernel rounds them when it loads its input
// (attention for q | k | v, the down projection for the gate / up pair: consumers with no RMSNorm in front), so the
// producer's last CTA ends with its RED instead of fence -> arrival -> finaliser -> store.
enum { kOutPlain = 0, kOutAddResidual = 1, kOutDeferred = 2 };

// Row-parallel GEMV fused with its TP exchange (sparse_gemv_kernel<.., TP = true>), low-latency PUSH protocol: the
// finaliser of a column tile sends the rank's fp16 partial of that tile straight into every rank's receive buffer as
// 8-byte words {2 fp16 values, epoch} -- data and flag travel in ONE store (single-copy atomic), so there is no fence, no
// separate flag and no remote read: the receiver polls its LOCAL buffer until the word carries the current epoch -- then
// sums the `world` words of every column pair in rank order (bit-identical replicas), adds the residual.
struct TpFuse {
  uint2* recv[kMaxTpWorld];  // receive buffer of every rank (peer-mapped; [rank] is local): [world sources][N / 2] words
  unsigned* epoch;           // local [kSsqParts]: exchanges completed per tile on this buffer set
  int rank, world, n;        // n = N (columns of the projection = row length of a source's slot)
};

// ---- what the C ABI entry points hand to the launcher ------------------------------------------------------------
struct GemvSeg {
  const __half* w;  // fp16 [Z][ldw], first column of the region
  long long ldw;    // elements
  int ncols;        // columns in the region
  int out_col0;     // first column in y / accumulator space
  float thr;        // fp16-representable threshold of the region
};

struct GemvParams {
  const __half* x;
  __half* y;
  const __half* norm_w;     // RMSNorm weight (kInRmsNorm)
  const __half* resid;      // residual stream added in the epilogue (kOutAddResidual)
  __half* x_out;            // optional: the (normalised) x actually consumed
  const float* ssq_in;      // RMSNorm: per-tile sums of squares of x left by its producer
  float* ssq_out;           // optional: per-tile sums of squares of y (for the next RMSNorm)
  unsigned char* keep_out;  // optional: keep mask (parity hook)
  const TpFuse* tp;         // optional (host pointer): fuse the TP exchange into the finaliser (row-parallel o / down)
  const float* in_acc;      // kInGateUpAcc: deferred sums of the gate / up pair, gate at [0, Z), up at [in_acc_stride, +Z)
  float* raw_out;           // kOutDeferred: fp32 sums are left here (zero on entry; the consumer's successor clears it)
  float* clear;             // optional: clear_n floats this launch zeroes once its producer is done (deferred buffers
  int clear_n;              //   whose readers have all completed by then)
  int in_acc_stride;
  int Z, nseg, mode, in_mode, out_mode, n_ssq_in;
  float eps;
  GemvSeg seg[kMaxSeg];
};

// ---- what the kernel sees -----------------------------------------------------------------------------------------
struct KSeg {
  const __half* w;
  long long ldw;
  int ncols, acc_col0, y_col0;
  float thr;
  int ctile0, nctiles;
};

struct KParams {
  const __half* x;
  const __half* norm_w;
  const float* ssq_in;
  __half* y;
  const __half* resid;
  float* ssq_out;
  __half* x_out;
  unsigned char* keep_out;
  float* acc;                 // fp32 accumulators (workspace), zero on entry and on exit; kOutDeferred: the caller's buffer
  const float* in_acc;        // kInGateUpAcc
  float* clear;               // optional: floats [0, clear_n) zeroed after the wait, clear_per_cta per CTA
  unsigned* counters;         // per-tile arrival counters (workspace), zero on entry and on exit
  unsigned long long* probe;  // optional [2]: gathered / dense weight bytes
  long long* dbg_time;        // TIMED instantiations: [grid][16] globaltimer stamps
  int Z, mode, in_mode, out_mode, nseg;
  int n_ssq_in, rows_per_cta, lanes_per_warp, nkchunks, total_ctiles, target;
  int in_acc_stride, clear_n, clear_per_cta;
  float eps;
  KSeg seg[kMaxSeg];
  TpFuse tp;  // TP instantiations only
};

// ---------------------------------------------------------------------------------- host
struct DeviceInfo {
  cudaError_t err;
  int sm_count;
  int smem_optin;
};
const DeviceInfo& device_info();
int set_error(int code, const char* msg);
int set_cuda_error(cudaError_t e, const char* where);
float round_threshold_fp16(double t);

// Launch context of the calling thread (teal_set_pdl / teal_set_config / teal_set_bytes_probe / teal_set_debug_timeline):
// thread-local, so two decoders driven from two host threads never see each other's settings.
struct LaunchCtx {
  bool pdl = false;                      // launch with the programmatic-dependent-launch attribute
  int variant = -1;                      // forced kernel variant (benchmarks, parity tests), -1 = by size
  int lanes_per_warp = 0;                // forced K range per warp, 0 = by size
  unsigned long long* probe = nullptr;   // device [2]: gathered / dense weight bytes
  long long* timeline = nullptr;         // device: one slot per launch for the TIMED instantiations
  int timeline_slots = 0, timeline_next = 0;
};
LaunchCtx& launch_ctx();
constexpr int kTimelineMaxCtas = 4096;  // CTAs recorded per launch (16 int64 each)
constexpr size_t kTimelineSlotInts = (size_t)kTimelineMaxCtas * 16;

size_t workspace_bytes(int max_out_cols);
float* workspace_ssq_scratch(void* workspace, size_t ws_bytes);

int launch_sparse_gemv(const GemvParams& gp, void* workspace, size_t ws_bytes, int out_cols_total, cudaStream_t stream);
int launch_ssq(const __half* x, int Z, float* ssq_out, int nparts, cudaStream_t stream);
int launch_mask_to_indices(const unsigned char* keep, int Z, int* idx, int* cnt, cudaStream_t stream);
int launch_simple(const __half* x, const __half* wt, __half* y, int Z, int N, long long ldw, float thr, cudaStream_t stream);

int launch_embed(const __half* emb, const int* token, __half* x, float* ssq_out, int dim, cudaStream_t stream);
int launch_attn_decode(const __half* qkv, const float* qkv_acc, const __half* freqs, __half* k_cache, __half* v_cache, const int* pos,
                       __half* out, int n_head, int n_kv, int head_dim, int max_seq, cudaStream_t stream);
int launch_dense_rowmajor(const __half* x, const __half* norm_w, float eps, const __half* W, __half* y,
                          unsigned long long* argmax_packed, int Z, int V, long long ldw, cudaStream_t stream);
int launch_residual_add(__half* x, const __half* part, float* ssq_out, int dim, cudaStream_t stream);
int launch_tp_allreduce_residual(const void* const* bufs, void* const* flags, int rank, int ws, __half* x, float* ssq_out,
                                 unsigned* epoch_cell, int dim, cudaStream_t stream);
int launch_sample_topk(const __half* logits, int V, float temperature, int top_k, const __half* q_ext, unsigned long long seed,
                       const int* pos, unsigned long long* argmax_packed, __half* probs_out, cudaStream_t stream);
int launch_advance(unsigned long long* argmax_packed, int* token, int* pos, int* out_tokens, int out_cap, cudaStream_t stream);

// ---------------------------------------------------------------------------------- device
// TEAL_EMU = the CPU emulation build of tests/emu (test infrastructure): the inline-PTX helpers the emulated kernels
// use have a C++ twin; the nvcc build never sees them.
#if defined(__CUDACC__) || defined(TEAL_EMU)
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ uint64_t make_evict_first_policy() {
#ifdef TEAL_EMU
  return 0;
#else
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
#endif
}

#ifdef TEAL_EMU
__device__ __forceinline__ void fence_acq_rel_gpu() { __threadfence(); }
__device__ __forceinline__ long long globaltimer_ns() { return clock64(); }
__device__ __forceinline__ void griddep_wait() {}
__device__ __forceinline__ void griddep_launch_dependents() {}
__device__ __forceinline__ float4 ld_cg_f4(const float* p) { return make_float4(__ldcg(p), __ldcg(p + 1), __ldcg(p + 2), __ldcg(p + 3)); }
__device__ __forceinline__ void st_cg_f4(float* p, float4 v) { p[0] = v.x; p[1] = v.y; p[2] = v.z; p[3] = v.w; }
__device__ __forceinline__ void red_add_v4(float* p, float4 v) {
  atomicAdd(p, v.x); atomicAdd(p + 1, v.y); atomicAdd(p + 2, v.z); atomicAdd(p + 3, v.w);
}
__device__ __forceinline__ uint4 ld_nc_u4(const uint4* p) { uint4 r; memcpy(&r, p, sizeof r); return r; }
#else
__device__ __forceinline__ void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
__device__ __forceinline__ long long globaltimer_ns() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch_dependents() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
__device__ __forceinline__ uint4 ld_nc_u4(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ float4 ld_cg_f4(const float* p) {
  float4 r;
  asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p)
               : "memory");
  return r;
}
__device__ __forceinline__ void st_cg_f4(float* p, float4 v) {
  asm volatile("st.global.cg.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}
// fp32 vector reduction into global memory (performed at L2, no return value).
__device__ __forceinline__ void red_add_v4(float* p, float4 v) {
  asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y),
               "f"(v.z), "f"(v.w)
               : "memory");
}
#endif  // TEAL_EMU

// fp16(silu(fp16(g))) * fp16(u), rounded like the reference's `F.silu(gate) * up` on fp16 tensors
// (gpt-fast/model.py:259): gate and up are fp16 op outputs, silu rounds to fp16, the product rounds
// to fp16.
__device__ __forceinline__ __half silu_mul_f16(float gate_acc, float up_acc) {
  const float g = __half2float(__float2half_rn(gate_acc));
  const float u = __half2float(__float2half_rn(up_acc));
  const float s = __half2float(__float2half_rn(g / (1.0f + expf(-g))));
  return __float2half_rn(s * u);
}
#endif  // __CUDACC__

}  // namespace teal
