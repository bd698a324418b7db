// Shared device helpers (inline PTX for sm_100a) and host-side plumbing for libteal_b200.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/teal_b200.h"

namespace teal {

constexpr int kConsumerThreads = 512;             // 16 consumer warps
constexpr int kProducerWarps = 4;                 // TMA-issuing warps
constexpr int kThreads = kConsumerThreads + 32 * kProducerWarps;
constexpr int kMaxSeg = 4;                        // column regions per call (q,k,v / gate,up)
constexpr int kMaxThr = 4;                        // distinct thresholds per call
constexpr int kMaxStages = 32;
constexpr int kMaxPieces = 8;
constexpr int kMaxCounters = 4096;                // column tiles per call
constexpr int kNumTmaConfigs = 6;                 // tile configurations of the TMA-bulk-ring kernel (config ids 1..6)

enum { kModePlain = 0, kModePairSilu = 1, kModeSurvivors = 2 };
enum { kInPlain = 0, kInRmsNorm = 1 };        // how x is produced from the input vector
enum { kOutPlain = 0, kOutAddResidual = 1 };  // what the finaliser adds before writing fp16
constexpr int kMaxTpWorld = 16;               // ranks of one tensor-parallel group (one NVSwitch domain)

struct GemvSeg {
  const __half* w;   // fp16 [Z][ldw], first column of the region
  long long ldw;     // elements
  int ncols;         // columns in the region
  int out_col0;      // first column in y / accumulator space
  int thr_slot;      // index into GemvParams::thr
  int ntiles;        // filled by the launcher
  int tile0;         // filled by the launcher (first completion counter)
  int pad;
};

struct GemvParams {
  const __half* x;
  __half* y;
  float* acc;          // fp32 accumulators (workspace), zero on entry and on exit
  unsigned* counters;  // per-tile completion counters (workspace), zero on entry and on exit
  int* dbg_idx;
  int* dbg_cnt;
  long long* dbg_time;  // optional [grid][16] phase timestamps (profiling builds of the bench tools)
  const __half* norm_w;  // RMSNorm weight (kInRmsNorm)
  const __half* resid;   // residual stream added in the epilogue (kOutAddResidual)
  __half* x_out;         // optional: the (normalised) x actually consumed
  const float* ssq_in;   // RMSNorm: per-tile sums of squares of x left by its producer
  int n_ssq_in;
  float* ssq_out;        // optional: per-tile sums of squares of y (for the next RMSNorm)
  unsigned char* keep_out;  // optional: keep mask (parity hook)
  int Z, nseg, nthr, mode;
  int nstages, list_cap, chunk, xs_elems;
  int in_mode, out_mode;
  float eps;
  int tile_n, sub_n, cg, rg;  // gather_gemv tiling (filled by its launcher)
  float thr[kMaxThr];  // fp16-representable thresholds
  GemvSeg seg[kMaxSeg];
  // appended (the offsets above are those of the hardware-validated build of the kernels that take GemvParams by value)
  const struct TpFuse* tp;  // optional (host pointer): fuse the TP exchange into the finaliser (row-parallel o / down)
  int chain;                // 0 | 1 = producer of a flag-chained pair (gate/up) | 2 = its consumer (down) | 3 = QKV producer of a
                            // chained attention block | 4 = its o-proj consumer: see RcParams
  const struct ChainAttn* ca;  // chain 3 / 4 (host pointer)
  int chain_wait_all;          // chain 1 / 3: do not griddepcontrol.wait, wait for the previous block's whole output instead
  int chain_release_all;       // chain 2 / 4: grid size of the next RMSNorm consumer to release the whole output to (0 = none)
  int chain_rows, chain_ctiles, chain_cols;  // consumer K chunk, consumer column tiles, producer tile width (lanes of h)
};

// ---- gather_gemv (cp.async per-warp rings) ------------------------------------------------------
constexpr int kMaxTiles = 16;

struct GatherTile {
  const __half* w;   // fp16 [Z][ldw], first column of the tile
  long long ldw;
  int width;         // columns (multiple of 8)
  int acc_col0;      // first column in the fp32 accumulator
  int y_col0;        // first column in y / resid
  int gate_acc0;     // pair mode: accumulator columns of the gate / up halves of this tile
  int up_acc0;
  int thr_slot;
  int cta0, nctas;   // CTAs [cta0, cta0 + nctas) split this tile's survivor list evenly
  unsigned magic;    // ceil(2^32 / nctas), 0 when nctas == 1
  int cidx, target;  // completion counter and the arrival count that completes it
  int pad;
};

struct GatherParams {
  const __half* x;
  __half* y;
  float* acc;
  unsigned* counters;
  const __half* norm_w;
  const __half* resid;
  __half* x_out;
  int* dbg_idx;
  int* dbg_cnt;
  long long* dbg_time;
  unsigned long long* probe;  // optional [2]: gathered / dense weight bytes, accumulated
  int Z, mode, in_mode, out_mode;
  int ntiles, tile_n, sub_n, cg, rg;
  int chunk, list_cap, region_bytes;
  float eps;
  float thr[kMaxThr];
  GatherTile tile[kMaxTiles];
};

// ---- rowchunk_gemv (many small CTAs) --------------------------------------------------------------
constexpr int kSsqParts = 64;  // fixed length of a sum-of-squares partials buffer (unused tail stays zero)

// Row-parallel GEMV fused with its TP exchange (rowchunk_gemv_kernel<.., TP = true>): the finaliser of a column tile
// writes the rank's fp16 partial of that tile into its peer-mapped buffer, publishes flags[peer][rank * kSsqParts + tile]
// = epoch to every peer, waits for the peers' flags of the same tile, sums the partials in rank order, adds the residual.
struct TpFuse {
  __half* part[kMaxTpWorld];     // this exchange's partial buffer [N] of every rank (peer-mapped; [rank] is local)
  unsigned* flags[kMaxTpWorld];  // flag array [world * kSsqParts] of every rank
  unsigned* epoch;               // local [kSsqParts]: exchanges completed per tile
  int rank, world;
};

// Geometry of a flag-chained attention block (QKV -> attention -> o-proj), shared by the three launches: which QKV column
// tiles a head's q / k / v slices live in, and how many o-proj CTAs read a head.
struct ChainAttn {
  unsigned* ready_qkv;  // [QKV column tiles]: attention CTAs that still have to read the tile (0 = not produced yet)
  unsigned* ready_att;  // [n_head]: o-proj CTAs that still have to read the head's output
  int n_head, n_kv, hd;
  int cols;             // QKV tile width
  int nseg, seg_col0[3], seg_ncols[3], seg_tile0[3];  // QKV column regions (1 = uniform thresholds, 3 = q | k | v) and their first tiles
  int rows_o, ctiles_o; // o-proj: K chunk per CTA, column tiles
  unsigned char readers[64];  // per QKV tile: attention CTAs that read it (chain_qkv_readers, filled on the host)
};
__host__ __device__ inline int chain_tile_of(const ChainAttn& c, int col) {  // QKV column -> column tile
  int s = 0;
  for (int j = 1; j < c.nseg; ++j) s += (col >= c.seg_col0[j]) ? 1 : 0;
  return c.seg_tile0[s] + (col - c.seg_col0[s]) / c.cols;
}
// column range [a, b) of slice `which` (0 q, 1 k, 2 v) of query head h
__host__ __device__ inline void chain_slice(const ChainAttn& c, int h, int which, int* a, int* b) {
  const int kvh = h / (c.n_head / c.n_kv);
  *a = which == 0 ? h * c.hd : which == 1 ? (c.n_head + kvh) * c.hd : (c.n_head + c.n_kv + kvh) * c.hd;
  *b = *a + c.hd;
}
// attention CTAs that read QKV tile `tile` (one count per (head, slice) whose columns touch it; a slice spans <= 2 tiles)
__host__ __device__ inline unsigned chain_qkv_readers(const ChainAttn& c, int tile) {
  unsigned n = 0;
  for (int h = 0; h < c.n_head; ++h)
    for (int w = 0; w < 3; ++w) {
      int a, b;
      chain_slice(c, h, w, &a, &b);
      n += (chain_tile_of(c, a) <= tile && tile <= chain_tile_of(c, b - 1)) ? 1u : 0u;
    }
  return n;
}
__host__ __device__ inline unsigned chain_att_readers(const ChainAttn& c, int h) {  // o-proj CTAs whose K chunk overlaps head h
  return (unsigned)((((h + 1) * c.hd - 1) / c.rows_o - (h * c.hd) / c.rows_o + 1) * c.ctiles_o);
}

struct RcSeg {
  const __half* w;
  long long ldw;
  int ncols, acc_col0, y_col0;
  float thr;
  int ctile0, nctiles;
};

struct RcParams {
  const __half* x;
  const __half* norm_w;
  const float* ssq_in;
  __half* y;
  const __half* resid;
  float* ssq_out;
  __half* x_out;
  unsigned char* keep_out;
  float* acc;
  unsigned* counters;
  unsigned long long* probe;
  int Z, mode, in_mode, out_mode, nseg;
  int n_ssq_in, rows_per_cta, nkchunks, total_ctiles, target;
  float eps;
  RcSeg seg[kMaxSeg];
  // generation 2 only (appended: the offsets of everything above are those of the hardware-validated build)
  int lanes_per_warp;  // contiguous input lanes owned by one warp (<= 32 * kRc2MaxPass)
  int spec_rows;       // speculative L2 prefetch: first rows of every warp's lane range ...
  int spec_ctas;       // ... issued by CTAs [0, spec_ctas) only (the first wave), before griddepcontrol.wait
  int spec_mode;       // 1 = prefetch.global.L2 per 128-byte line, 2 = cp.async.bulk.prefetch.L2 per row segment
  long long* dbg_time; // TIMED instantiations: [grid][16] globaltimer stamps (teal_set_debug_timestamps)
  TpFuse tp;           // TP instantiations only
  // CHAIN instantiations (flag-chained launches: the consumer grid never executes griddepcontrol.wait, it waits for the
  // producer's tiles): ready[j] = number of consumer CTAs that still have to read h tile j (0 = not produced yet)
  unsigned* chain_ready;
  int chain_rows, chain_ctiles, chain_cols, chain_group, chain_rot;
  unsigned* chain_all_in;   // CIN == 2: consumer CTAs that still have to read the producer's whole output (0 = not final yet)
  unsigned* chain_all_out;  // COUT == 4: the cell of the NEXT RMSNorm consumer, ...
  unsigned* chain_done;     // ... the count of final tiles of this launch, ...
  int chain_all_readers;    // ... and that consumer's grid size
  ChainAttn ca;        // CHAIN == 3 (QKV producer of a chained attention block)
};
constexpr int kRc2MaxPass = 4;  // 32-lane groups one warp of a generation-2 CTA may own

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

size_t workspace_bytes(int max_out_cols);
float* workspace_ssq_scratch(void* workspace, size_t ws_bytes);
int set_config(int id);
void set_debug_timestamps(long long* buf);
long long* debug_timestamps();
constexpr int kTimelineMaxCtas = 4096;                       // CTAs recorded per launch (16 int64 each)
constexpr size_t kTimelineSlotInts = (size_t)kTimelineMaxCtas * 16;
void set_debug_timeline(long long* buf, int slots);
long long* next_debug_timeline_slot();
int forced_config();
void set_pdl(bool on);
bool pdl_enabled();
int sub_cols_override();
bool wave_balance_enabled();
void chain_describe(int dim, int inter, int* rows_c, int* ctiles_c, int* cols_p);
int chain_consumer_ctas(int kind, int dim, int a, int b, int c, bool uniform);
int chain_attn_describe(ChainAttn* ca, void* workspace, size_t ws_bytes, int dim, int n_head, int n_kv, int head_dim, bool uniform);
int set_tuning(int knob, int value);
int tuning(int knob);
void set_bytes_probe(unsigned long long* p);
unsigned long long* bytes_probe();
int launch_gather_gemv(GemvParams& p, void* workspace, size_t ws_bytes, int out_cols_total, cudaStream_t stream,
                       int tiles_per_matrix);
int launch_rowchunk(GemvParams& gp, void* workspace, size_t ws_bytes, int out_cols_total, cudaStream_t stream, int variant);
int launch_ssq(const __half* x, int Z, float* ssq_out, int nparts, cudaStream_t stream);
int launch_mask_to_indices(const unsigned char* keep, int Z, int* idx, int* cnt, cudaStream_t stream);
int dispatch_gemv(GemvParams& p, void* workspace, size_t ws_bytes, int out_cols_total, cudaStream_t stream);
int launch_gemv(GemvParams& p, void* workspace, size_t ws_bytes, int out_cols_total, cudaStream_t stream);
int launch_simple(const __half* x, const __half* wt, __half* y, int Z, int N, long long ldw, float thr,
                  cudaStream_t stream);

int launch_embed(const __half* emb, const int* token, __half* x, float* ssq_out, int dim, cudaStream_t stream);
int launch_attn_decode(const __half* qkv, const __half* freqs, __half* k_cache, __half* v_cache, const int* pos,
                       __half* out, int n_head, int n_kv, int head_dim, int max_seq, cudaStream_t stream);
int launch_attn_decode_chain(const __half* qkv, const __half* freqs, __half* k_cache, __half* v_cache, const int* pos, __half* out,
                             int max_seq, const ChainAttn& ca, cudaStream_t stream);
int launch_dense_rowmajor(const __half* x, const __half* norm_w, float eps, const __half* W, __half* y,
                          unsigned long long* argmax_packed, int Z, int V, long long ldw, cudaStream_t stream);
int launch_residual_add(__half* x, const __half* part, float* ssq_out, int dim, cudaStream_t stream);
int launch_tp_allreduce_residual(const void* const* bufs, void* const* flags, int rank, int ws, __half* x, float* ssq_out,
                                 unsigned* epoch_cell, int dim, cudaStream_t stream);
int launch_sample_topk(const __half* logits, int V, float temperature, int top_k, const __half* q_ext, unsigned long long seed,
                       const int* pos, unsigned long long* argmax_packed, __half* probs_out, cudaStream_t stream);
int launch_advance(unsigned long long* argmax_packed, int* token, int* pos, int* out_tokens, int out_cap,
                   cudaStream_t stream);

// ---------------------------------------------------------------------------------- device
// TEAL_EMU = the CPU emulation build of tests/emu (test infrastructure): the inline-PTX helpers the emulated kernels
// use have a C++ twin; the nvcc build never sees them.
#if defined(__CUDACC__) || defined(TEAL_EMU)
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok;
}
// Spin with a watchdog: a protocol bug traps after ~2 s instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (((++spins) & 0xfffu) == 0 && clock64() - t0 > 4000000000ll) __trap();
  }
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
// 1-D TMA bulk copy global -> shared, completion signalled on an mbarrier (bytes % 16 == 0).
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar,
                                         uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
      ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(policy)
      : "memory");
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
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
__device__ __forceinline__ void consumer_bar_sync() {
  asm volatile("bar.sync 1, %0;" ::"n"(kConsumerThreads) : "memory");
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
