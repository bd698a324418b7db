// rowchunk_gemv: fused [RMSNorm ->] threshold -> warp-level survivor compaction -> gather-GEMV
// [-> +residual | SiLU(gate)*up] for sm_100a, built from MANY SMALL CTAs.
//
// Why this shape (measured on B200): a batch-1 projection moves 17-117 MB, i.e.
// it lasts 3-18 us at HBM speed, so every serialised latency inside the kernel (x load, compaction,
// first DRAM access, cross-CTA reduction, completion protocol) is worth 5-15 % of its runtime.  One
// fat CTA per SM runs those phases in lock-step on all 148 SMs and leaves HBM idle during each of
// them.  Here a CTA is 4 warps and owns (one 256/512-column tile) x (one K-chunk of 128-256 input
// lanes): ~3600 CTAs per projection, 6-8 resident per SM, all at different phases, so HBM always has
// somebody streaming; the hardware CTA scheduler doubles as the load balancer for the random
// survivor counts.
//
//   per warp: 32 lanes of x -> (optional RMSNorm scale) -> |x| > thr -> ballot = survivor set
//             -> for each surviving lane m (warp-uniform): LDG.128 of W[m][tile] (512 B coalesced per
//                instruction, up to 8 rows in flight per lane), FFMA2 in fp32 against x[m] (shuffle)
//   per CTA : 4 warp partials -> shared memory -> fp32 vector RED into the L2-resident accumulator
//   per tile: arrival counter; the last CTA converts the tile to fp16, applies the epilogue
//             (+residual / SiLU(gate)*up), writes the tile's sum of squares for the next RMSNorm and
//             re-zeroes accumulator + counter (no memset launch, nothing to clean up afterwards).
//
// The RMSNorm of the consumer is fused without any CTA reading the whole vector: the producer of the
// residual stream (embedding, o-proj, down-proj finalisers) leaves per-tile sums of squares, the
// consumer adds <= 64 floats.
//
// Replaces kernels/sparse_gemv.py:45-83 (splitk_sparse_gemv_kernel + its Y.zero_() pre-hook) and
// :147-194 (qkv_kernel) of the reference; thresholds are selected per column region exactly.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "common.cuh"

namespace teal {

namespace {


__device__ __forceinline__ void rc_ffma2(float2& acc, const float2 w, const float2 x) {
#ifdef TEAL_EMU
  acc.x = fmaf(w.x, x.x, acc.x);
  acc.y = fmaf(w.y, x.y, acc.y);
#else
  asm("fma.rn.f32x2 %0, %1, %2, %0;"
      : "+l"(*reinterpret_cast<unsigned long long*>(&acc))
      : "l"(*reinterpret_cast<const unsigned long long*>(&w)), "l"(*reinterpret_cast<const unsigned long long*>(&x)));
#endif
}
// streaming 16-byte load: read-only path, no L1 allocation, L2 evict-first (weights are read once per token)
__device__ __forceinline__ uint4 ld_stream_u4(const void* p, uint64_t policy) {
  uint4 r;
#ifdef TEAL_EMU
  (void)policy;
  memcpy(&r, p, sizeof r);
#else
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p), "l"(policy));
#endif
  return r;
}

// the same without the L2 cache-policy descriptor (the descriptor costs two R2UR per load)
__device__ __forceinline__ uint4 ld_stream_nohint_u4(const void* p) {
  uint4 r;
#ifdef TEAL_EMU
  memcpy(&r, p, sizeof r);
#else
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
#endif
  return r;
}
// f32 += f16 * f16 (FHFMA): the product of two fp16 values is exact in fp32, so this is fma(float(w), float(x), acc)
__device__ __forceinline__ void rc_fhfma2(float2& acc, uint32_t w2, unsigned short xh) {
#ifdef TEAL_EMU
  const float x = __half2float(__ushort_as_half(xh));
  acc.x = fmaf(__half2float(__ushort_as_half((unsigned short)(w2 & 0xffffu))), x, acc.x);
  acc.y = fmaf(__half2float(__ushort_as_half((unsigned short)(w2 >> 16))), x, acc.y);
#else
  asm("{\n"
      ".reg .b16 lo, hi;\n"
      "mov.b32 {lo, hi}, %2;\n"
      "fma.rn.f32.f16 %0, lo, %3, %0;\n"
      "fma.rn.f32.f16 %1, hi, %3, %1;\n"
      "}\n"
      : "+f"(acc.x), "+f"(acc.y)
      : "r"(w2), "h"(xh));
#endif
}
#ifdef TEAL_EMU
// the emulation touches every byte a prefetch names, so an out-of-range speculative address is caught by ASan / a fault
__device__ __forceinline__ void rc_prefetch_l2(const void* p) { (void)*reinterpret_cast<const volatile char*>(p); }
__device__ __forceinline__ void rc_bulk_prefetch_l2(const void* p, uint32_t bytes) {
  for (uint32_t i = 0; i < bytes; i += 16) (void)*(reinterpret_cast<const volatile char*>(p) + i);
}
__device__ __forceinline__ __half rc_ldcg_half(const __half* p) { return __ushort_as_half(__ldcg(reinterpret_cast<const unsigned short*>(p))); }
__device__ __forceinline__ uint32_t rc_lds_u32(const uint32_t* p) { return *reinterpret_cast<const volatile uint32_t*>(p); }
__device__ __forceinline__ uint64_t rc_row_address(uint64_t base, uint32_t row, uint32_t row_bytes) { return base + (uint64_t)row * row_bytes; }
#else
__device__ __forceinline__ void rc_prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void rc_bulk_prefetch_l2(const void* p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
__device__ __forceinline__ __half rc_ldcg_half(const __half* p) {
  unsigned short v;
  asm volatile("ld.global.cg.u16 %0, [%1];" : "=h"(v) : "l"(p) : "memory");
  return __ushort_as_half(v);
}
// volatile asm on purpose: the compiler must neither duplicate the address arithmetic per 16-byte load nor move the
// list read across the weight loads
__device__ __forceinline__ uint32_t rc_lds_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
  return v;
}
__device__ __forceinline__ uint64_t rc_row_address(uint64_t base, uint32_t row, uint32_t row_bytes) {
  uint64_t a;
  asm volatile("mad.wide.u32 %0, %1, %2, %3;" : "=l"(a) : "r"(row), "r"(row_bytes), "l"(base));
  return a;
}
#endif

// cp.async (LDGSTS) ring helpers of the RING instantiations
#ifdef TEAL_EMU
__device__ __forceinline__ void rc_cp_async16(unsigned char* base, uint32_t off, const void* src, uint64_t) { memcpy(base + off, src, 16); }
__device__ __forceinline__ void rc_cp_async_commit() {}
template <int N> __device__ __forceinline__ void rc_cp_async_wait() {}
__device__ __forceinline__ uint4 rc_lds_u4(const unsigned char* base, uint32_t off) { uint4 r; memcpy(&r, base + off, 16); return r; }
#else
__device__ __forceinline__ void rc_cp_async16(unsigned char* base, uint32_t off, const void* src, uint64_t policy) {
  asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(smem_u32(base) + off), "l"(src), "l"(policy) : "memory");
}
__device__ __forceinline__ void rc_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void rc_cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ uint4 rc_lds_u4(const unsigned char* base, uint32_t off) {
  uint4 r;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(smem_u32(base) + off) : "memory");
  return r;
}
#endif
constexpr int kRcRingList = 128;  // list entries per warp of a RING instantiation (4 lane groups)

// gpu-scope flag traffic of the flag-chained launches (CHAIN instantiations)
#ifdef TEAL_EMU
__device__ __forceinline__ void rc_st_release_gpu(unsigned* p, unsigned v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }
__device__ __forceinline__ unsigned rc_ld_acquire_gpu(const unsigned* p) { ::emu::yield(); return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
#else
__device__ __forceinline__ void rc_st_release_gpu(unsigned* p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned rc_ld_acquire_gpu(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
#endif
// system-scope flag traffic of the fused TP exchange (TP instantiations)
#ifdef TEAL_EMU
__device__ __forceinline__ void rc_st_release_sys(unsigned* p, unsigned v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }
__device__ __forceinline__ unsigned rc_ld_acquire_sys(const unsigned* p) { ::emu::yield(); return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
__device__ __forceinline__ uint4 rc_ld_volatile_u4(const void* p) { uint4 r; memcpy(&r, p, sizeof r); return r; }
#else
__device__ __forceinline__ void rc_st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned rc_ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint4 rc_ld_volatile_u4(const void* p) {
  uint4 r;
  asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
  return r;
}
#endif

}  // namespace

// VEC: 16-byte vectors per lane per row (column tile = 256 * VEC); U: rows in flight per lane;
// WARPS: warps per CTA (each owns 32-lane groups of the CTA's K chunk).
// TIMED: profiling instantiation (teal_set_debug_timestamps): thread 0 of every CTA records globaltimer stamps of its
// phases -- tools/rowchunk_timeline.py turns them into the wave / head / tail picture of a launch.  The production
// instantiation (TIMED = false) is SASS-identical to the build without the stamps (tools/sass_diff.py).
// TP: the row-parallel projections under tensor parallelism (o / down, gpt-fast/tp.py:117,130) with the exchange step
// (tp.py:120-121,139-140) fused into the tile finaliser -- see TpFuse.  Opt-in, not yet run on hardware.
// OPT (bit mask, experiments on the validated structure, one change at a time -- TEAL_TUNE_V1_OPTS):
//   1 = the first lane group's x and norm weights are requested before / together with the sum-of-squares partials
//       (one L2 round trip in the head instead of two); 2 = FHFMA arithmetic instead of conversions + FFMA2;
//   4 = 32-bit row offsets (one IMAD.WIDE.U32 per row instead of 64-bit multiplies per 16-byte load).
// CHAIN (flag-chained launches; opt-in, not yet run on hardware): 1 = producer of the MLP pair
//   (gate/up): CTAs run group-major (a group of tile pairs x all K chunks, so tiles complete early) and the finaliser of
//   pair tile j releases ready[j] = #consumer CTAs; 2 = consumer (down): launched with the PDL attribute but WITHOUT
//   griddepcontrol.wait -- every CTA waits for the h tiles that cover its K chunk, reads them at L2 and counts
//   ready[j] down again (the cell is back to 0 when the step is over: nothing to reset, graph-replayable).  Safe
//   because every producer CTA issued launch_dependents when it started: by the time a consumer CTA exists, all
//   producer CTAs are resident or done, and everything older than the producer is complete.
//   CHAIN = 10 * CIN + COUT.  CIN: 0 = griddepcontrol.wait, 1 = wait for the producer's tiles covering my K chunk (o-proj, down),
//   2 = wait for the producer's WHOLE output (the RMSNorm consumers QKV and gate/up: `chain_all_in` holds the number of consumer
//   CTAs once every tile of the producer is final).  COUT: 0 = none, 1 = release h tile pairs (gate/up), 3 = release QKV tiles,
//   4 = count final tiles and release the whole output to the next RMSNorm consumer (o-proj, down).
// RING (> 0: slots per warp): the weight rows land in a per-warp shared-memory ring through cp.async (LDGSTS, L2 evict-first)
//   instead of registers -- RING - 1 rows stay in flight per warp all the time (cp.async.wait_group is an in-order count, a
//   register ring refilled row by row serialises on the load scoreboards), the bytes in flight per SM are bounded by shared
//   memory (7 CTAs x 32 KB) instead of by the register file, and the survivors of the warp's lane groups are compacted into a
//   per-warp list {local row, fp16 x} in shared memory first.
template <int VEC, int U, int WARPS, bool TIMED = false, bool TP = false, int OPT = 0, int CHAIN = 0, int RING = 0>
__global__ void __launch_bounds__(WARPS * 32, ((OPT != 0 || CHAIN != 0) && RING == 0 && !TP && !TIMED) ? 28 / WARPS : 0)
rowchunk_gemv_kernel(const __grid_constant__ RcParams p) {
  constexpr int COLS = 256 * VEC;
  constexpr int CIN = CHAIN / 10, COUT = CHAIN % 10;
  // (the chained twins are profiled through the same hook, decided at run time: they are experimental anyway)
  constexpr bool kCanStamp = TIMED || CHAIN != 0;
  long long* const dbg = (kCanStamp && (TIMED || p.dbg_time)) ? p.dbg_time + (size_t)blockIdx.x * 16 : nullptr;
  auto stamp = [&](int i) {
    if constexpr (kCanStamp) {
      if (threadIdx.x == 0 && (TIMED || dbg)) dbg[i] = globaltimer_ns();
    }
  };
  stamp(0);
  constexpr int kRcWarps = WARPS;
  constexpr int kRcThreads = WARPS * 32;
  constexpr size_t kRedBytes = (size_t)kRcWarps * COLS * 4, kRingBytes = (size_t)RING * kRcWarps * COLS * 2;
  __shared__ __align__(16) unsigned char s_raw[kRedBytes > kRingBytes ? kRedBytes : kRingBytes];
  float (*const s_red)[COLS] = reinterpret_cast<float (*)[COLS]>(s_raw);  // (RING: the ring first, the warp partials afterwards)
  __shared__ uint32_t s_list[RING > 0 ? kRcWarps : 1][RING > 0 ? kRcRingList : 1];
  __shared__ float s_ssq[kRcWarps];
  __shared__ int s_last;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // ---- which (column tile, K chunk) is mine: column tile varies fastest so that concurrently running
  // CTAs read neighbouring pieces of the same weight rows ---------------------------------------------
  int ct = (int)(blockIdx.x % (unsigned)p.total_ctiles);
  int kc = (int)(blockIdx.x / (unsigned)p.total_ctiles);
  if constexpr (COUT == 1) {  // group-major: order position -> (gate tile j, up tile j neighbours), K chunk slow inside a group
    const int T = p.total_ctiles, G = p.chain_group, per = G * p.nkchunks;
    const int g = (int)blockIdx.x / per, r = (int)blockIdx.x - g * per;
    const int gsz = min(G, T - g * G);
    kc = r / gsz;
    const int pos = g * G + (r - kc * gsz);
    ct = (pos & 1) ? (T >> 1) + (pos >> 1) : (pos >> 1);
  }
  if constexpr (COUT == 3) {  // group-major, k / v tiles first (every head needs its kv head's tiles), then the q tiles
    const int T = p.total_ctiles, G = p.chain_group, per = G * p.nkchunks;
    const int g = (int)blockIdx.x / per, r = (int)blockIdx.x - g * per;
    const int gsz = min(G, T - g * G);
    kc = r / gsz;
    ct = g * G + (r - kc * gsz) + p.chain_rot;
    if (ct >= T) ct -= T;
  }
  int s = 0;
#pragma unroll
  for (int j = 1; j < kMaxSeg; ++j) s += (j < p.nseg && ct >= p.seg[j].ctile0) ? 1 : 0;
  const RcSeg& sg = p.seg[s];
  const int col0 = (ct - sg.ctile0) * COLS;
  const int wcols = min(COLS, sg.ncols - col0);
  const float thr = sg.thr;
  bool vact[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) vact[v] = lane * 8 + 256 * v < wcols;
  const __half* wbase = sg.w + col0 + lane * 8;
  const long long ldw = sg.ldw;
  const uint64_t policy = make_evict_first_policy();

  [[maybe_unused]] __half early_nw = __ushort_as_half((unsigned short)0), early_x = __ushort_as_half((unsigned short)0);
  [[maybe_unused]] const int early_k = kc * p.rows_per_cta + warp * 32 + lane;
  [[maybe_unused]] const bool early_ok = early_k < min(p.Z, (kc + 1) * p.rows_per_cta);
  if constexpr ((OPT & 1) != 0) {
    if (early_ok && p.in_mode == kInRmsNorm) early_nw = p.norm_w[early_k];  // a constant: before the producer is done
  }

  if constexpr (CIN == 1) {
    griddep_launch_dependents();
    // the h tiles covering my K chunk: one polling thread per tile
    const int cb = kc * p.rows_per_cta, ce = min(p.Z, cb + p.rows_per_cta);
    const int t0 = cb / p.chain_cols, t1 = (ce - 1) / p.chain_cols;
    if (tid <= t1 - t0) {
      const unsigned* cell = p.chain_ready + t0 + tid;
      const long long w0 = clock64();
      while (rc_ld_acquire_gpu(cell) == 0u) {
        if (clock64() - w0 > 4000000000ll) __trap();  // a protocol bug traps instead of hanging the GPU
      }
    }
    __syncthreads();
  } else if constexpr (CIN == 2) {
    griddep_launch_dependents();
    if (tid == 0) {  // the producer's whole output (residual stream + sum-of-squares partials)
      const long long w0 = clock64();
      while (rc_ld_acquire_gpu(p.chain_all_in) == 0u) {
        __nanosleep(64);  // ~1000 resident CTAs poll this one cell: keep the line's request rate well below one L2 slice's
        if (clock64() - w0 > 4000000000ll) __trap();
      }
    }
    __syncthreads();
  } else {
    griddep_wait();                 // PDL: everything above overlapped the previous kernel's tail
    griddep_launch_dependents();    // the next kernel may start launching once all our CTAs are running
  }
  stamp(1);
  if constexpr ((OPT & 1) != 0) {
    if (early_ok) early_x = rc_ldcg_half(p.x + early_k);  // in flight together with the sum-of-squares partials below
  }

  // ---- RMSNorm scale from the producer's per-tile sums of squares (gpt-fast/model.py:286-291) --------
  float rs = 1.f;
  if (p.in_mode == kInRmsNorm) {
    float t = 0.f;
    for (int i = lane; i < p.n_ssq_in; i += 32) t += __ldcg(p.ssq_in + i);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) t += __shfl_xor_sync(0xffffffffu, t, d);
    rs = rsqrtf(t / (float)p.Z + p.eps);
  }

  float2 acc[VEC][4];
#pragma unroll
  for (int v = 0; v < VEC; ++v)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[v][c] = make_float2(0.f, 0.f);

  const int kbeg = kc * p.rows_per_cta;
  const int kend = min(p.Z, kbeg + p.rows_per_cta);
  int my_nnz = 0;
  if constexpr (RING > 0) {
    // ---- per-warp survivor list in shared memory --------------------------------------------------------------------
    uint32_t* const list = s_list[warp];
    int n = 0;
    for (int k0 = kbeg + warp * 32; k0 < kend; k0 += kRcWarps * 32) {
      const int k = k0 + lane;
      float xf = 0.f;
      if (k < kend) {
        const __half xh = p.x[k];
        if (p.in_mode == kInRmsNorm) {
          xf = __half2float(__hmul(__float2half_rn(__half2float(xh) * rs), p.norm_w[k]));
          if (p.x_out && ct == 0) p.x_out[k] = __float2half_rn(xf);
        } else {
          xf = __half2float(xh);
        }
      }
      const bool keep = (k < kend) && !(fabsf(xf) <= thr);  // strict >, NaN kept (it must poison y like mask*x)
      const unsigned mask = __ballot_sync(0xffffffffu, keep);
      if (p.keep_out && ct == 0 && k < kend) p.keep_out[k] = keep ? 1 : 0;
      if (keep) list[n + __popc(mask & ((1u << lane) - 1u))] = (uint32_t)(k - kbeg) | ((uint32_t)__half_as_ushort(__float2half_rn(xf)) << 16);
      n += __popc(mask);
    }
    __syncwarp();
    my_nnz = n;
    if (p.mode == kModeSurvivors) n = 0;  // parity hook: thresholds only, no weights
    stamp(2);
    // ---- cp.async ring: RING - 1 rows in flight, every lane copies and later reads its own 16-byte pieces -----------
    constexpr uint32_t ROWB = COLS * 2;
    const uint32_t ring0 = (uint32_t)warp * (RING * ROWB) + (uint32_t)lane * 16u;  // byte offset into s_raw
    const __half* const wrow0 = wbase + (long long)kbeg * ldw;
    auto issue = [&](int i) {
      const uint32_t ent = list[i];
      const __half* src = wrow0 + (long long)(ent & 0xffffu) * ldw;
      const uint32_t dst = ring0 + (uint32_t)(i % RING) * ROWB;
#pragma unroll
      for (int v = 0; v < VEC; ++v)
        if (vact[v]) rc_cp_async16(s_raw, dst + 512u * v, src + 256 * v, policy);
    };
#pragma unroll
    for (int i = 0; i < RING - 1; ++i) {
      if (i < n) issue(i);
      rc_cp_async_commit();
    }
    for (int c = 0; c < n; ++c) {
      if (c + RING - 1 < n) issue(c + RING - 1);
      rc_cp_async_commit();
      rc_cp_async_wait<RING - 1>();
      const uint32_t ent = list[c];
      const unsigned short xb = (unsigned short)(ent >> 16);
      const float xs = __half2float(__ushort_as_half(xb));
      const float2 x2 = make_float2(xs, xs);
      const uint32_t slot = ring0 + (uint32_t)(c % RING) * ROWB;
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        if (vact[v]) {
          const uint4 wv = rc_lds_u4(s_raw, slot + 512u * v);
          if constexpr ((OPT & 2) != 0) {
            rc_fhfma2(acc[v][0], wv.x, xb);
            rc_fhfma2(acc[v][1], wv.y, xb);
            rc_fhfma2(acc[v][2], wv.z, xb);
            rc_fhfma2(acc[v][3], wv.w, xb);
          } else {
            rc_ffma2(acc[v][0], __half22float2(*reinterpret_cast<const __half2*>(&wv.x)), x2);
            rc_ffma2(acc[v][1], __half22float2(*reinterpret_cast<const __half2*>(&wv.y)), x2);
            rc_ffma2(acc[v][2], __half22float2(*reinterpret_cast<const __half2*>(&wv.z)), x2);
            rc_ffma2(acc[v][3], __half22float2(*reinterpret_cast<const __half2*>(&wv.w)), x2);
          }
        }
      }
    }
    __syncthreads();  // the warp partials below overwrite the rings
  } else
  for (int k0 = kbeg + warp * 32; k0 < kend; k0 += kRcWarps * 32) {
    const int k = k0 + lane;
    float xf = 0.f;
    if (k < kend) {
      const bool early = (OPT & 1) != 0 && k0 == kbeg + warp * 32;
      const __half xh = early ? early_x : (CIN != 0 ? rc_ldcg_half(p.x + k) : p.x[k]);  // (chained: written by a grid that is still running -> L2)
      if (p.in_mode == kInRmsNorm) {
        // fp32 norm, cast to fp16, multiply by the fp16 weight in fp16
        xf = __half2float(__hmul(__float2half_rn(__half2float(xh) * rs), early ? early_nw : p.norm_w[k]));
        if (p.x_out && ct == 0) p.x_out[k] = __float2half_rn(xf);
      } else {
        xf = __half2float(xh);
      }
    }
    const bool keep = (k < kend) && !(fabsf(xf) <= thr);  // strict >, NaN kept (it must poison y like mask*x)
    unsigned mask = __ballot_sync(0xffffffffu, keep);
    if (p.keep_out && ct == 0 && k < kend) p.keep_out[k] = keep ? 1 : 0;
    my_nnz += __popc(mask);
    if (p.mode == kModeSurvivors) mask = 0u;  // parity hook: thresholds only, no weights
    if (k0 == kbeg + warp * 32) stamp(2);     // first survivor set known
    [[maybe_unused]] bool first_batch = kCanStamp && k0 == kbeg + warp * 32;
    while (mask) {
      int bit[U];
      uint4 wv[U][VEC];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        bit[u] = mask ? (__ffs(mask) - 1) : -1;
        mask &= mask - 1;  // no-op when mask == 0
        if (bit[u] >= 0) {
          const __half* src;
          if constexpr ((OPT & 4) != 0)
            src = reinterpret_cast<const __half*>(rc_row_address(reinterpret_cast<uint64_t>(wbase), (uint32_t)(k0 + bit[u]), (uint32_t)ldw * 2u));
          else
            src = wbase + (long long)(k0 + bit[u]) * ldw;
#pragma unroll
          for (int v = 0; v < VEC; ++v)
            if (vact[v]) wv[u][v] = (OPT & 8) ? ld_stream_nohint_u4(src + 256 * v) : ld_stream_u4(src + 256 * v, policy);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (bit[u] >= 0) {
          const float xs = __shfl_sync(0xffffffffu, xf, bit[u]);
          const float2 x2 = make_float2(xs, xs);
          [[maybe_unused]] const unsigned short xb = __half_as_ushort(__float2half_rn(xs));  // exact: xs is an fp16 value
#pragma unroll
          for (int v = 0; v < VEC; ++v) {
            if constexpr ((OPT & 2) != 0) {
              if (vact[v]) {
                rc_fhfma2(acc[v][0], wv[u][v].x, xb);
                rc_fhfma2(acc[v][1], wv[u][v].y, xb);
                rc_fhfma2(acc[v][2], wv[u][v].z, xb);
                rc_fhfma2(acc[v][3], wv[u][v].w, xb);
              }
            } else if (vact[v]) {
              rc_ffma2(acc[v][0], __half22float2(*reinterpret_cast<const __half2*>(&wv[u][v].x)), x2);
              rc_ffma2(acc[v][1], __half22float2(*reinterpret_cast<const __half2*>(&wv[u][v].y)), x2);
              rc_ffma2(acc[v][2], __half22float2(*reinterpret_cast<const __half2*>(&wv[u][v].z)), x2);
              rc_ffma2(acc[v][3], __half22float2(*reinterpret_cast<const __half2*>(&wv[u][v].w)), x2);
            }
          }
        }
      }
      if constexpr (kCanStamp) {
        if (first_batch && (TIMED || dbg)) {  // the first rows have arrived from DRAM and been consumed
          first_batch = false;
          if (threadIdx.x == 0) dbg[3] = globaltimer_ns() + (long long)(acc[0][0].x == 1.2345e-30f);
        }
      }
    }
  }
  stamp(4);
  if (p.probe && lane == 0) {  // measurement probe: gathered vs dense weight bytes
    atomicAdd(p.probe, (unsigned long long)my_nnz * (unsigned long long)wcols * 2ull);
    const int rows = max(0, min(kend, kbeg + p.rows_per_cta) - kbeg);
    if (warp == 0) atomicAdd(p.probe + 1, (unsigned long long)rows * (unsigned long long)wcols * 2ull);
  }
  if (p.mode == kModeSurvivors) return;

  // ---- CTA partial: 4 warps -> smem -> fp32 vector RED into the L2 accumulator ------------------------
#pragma unroll
  for (int v = 0; v < VEC; ++v) {
    float* d = &s_red[warp][lane * 8 + 256 * v];
    *reinterpret_cast<float4*>(d) = make_float4(acc[v][0].x, acc[v][0].y, acc[v][1].x, acc[v][1].y);
    *reinterpret_cast<float4*>(d + 4) = make_float4(acc[v][2].x, acc[v][2].y, acc[v][3].x, acc[v][3].y);
  }
  __syncthreads();
  if constexpr (CIN == 1) {  // every warp has consumed its lanes of h: this CTA is done with the tiles it waited for
    const int cb = kc * p.rows_per_cta, ce = min(p.Z, cb + p.rows_per_cta);
    const int t0 = cb / p.chain_cols, t1 = (ce - 1) / p.chain_cols;
    if (tid <= t1 - t0) atomicSub(p.chain_ready + t0 + tid, 1u);
  }
  if constexpr (CIN == 2) {  // ... with its lanes of the residual stream and the partials
    if (tid == 0) atomicSub(p.chain_all_in, 1u);
  }
  const int acc0 = sg.acc_col0 + col0;
  for (int c = tid * 4; c < wcols; c += kRcThreads * 4) {
    float4 a = *reinterpret_cast<const float4*>(&s_red[0][c]);
#pragma unroll
    for (int w = 1; w < kRcWarps; ++w) {
      const float4 b = *reinterpret_cast<const float4*>(&s_red[w][c]);
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    if (a.x != 0.f || a.y != 0.f || a.z != 0.f || a.w != 0.f) red_add_v4(p.acc + acc0 + c, a);
  }
  // ---- arrival counter: REDs of all threads happen-before the barrier, thread 0's release fence is
  // cumulative over them; the last arrival acquires and finalises the tile ------------------------------
  __syncthreads();
  stamp(5);
  const int pair = p.mode == kModePairSilu;
  const int cidx = pair ? (ct - sg.ctile0) : ct;
  if (tid == 0) {
    fence_acq_rel_gpu();
    const unsigned old = atomicAdd(&p.counters[cidx], 1u);
    const int last = ((int)old + 1 == p.target) ? 1 : 0;
    if (last) fence_acq_rel_gpu();
    s_last = last;
    if (kCanStamp && (TIMED || dbg)) {
      unsigned smid = blockIdx.x;
#ifndef TEAL_EMU
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
#endif
      dbg[6] = globaltimer_ns();
      dbg[8] = smid;
      dbg[9] = my_nnz;
      dbg[10] = last;
      dbg[7] = 0;
    }
  }
  __syncthreads();
  if (!s_last) return;

  float ssq = 0.f;
  if (pair) {
    const int g0 = p.seg[0].acc_col0 + col0, u0 = p.seg[1].acc_col0 + col0;
    for (int c = tid * 4; c < wcols; c += kRcThreads * 4) {
      const float4 g = ld_cg_f4(p.acc + g0 + c), u = ld_cg_f4(p.acc + u0 + c);
      st_cg_f4(p.acc + g0 + c, make_float4(0.f, 0.f, 0.f, 0.f));
      st_cg_f4(p.acc + u0 + c, make_float4(0.f, 0.f, 0.f, 0.f));
      const __half2 h0 = __halves2half2(silu_mul_f16(g.x, u.x), silu_mul_f16(g.y, u.y));
      const __half2 h1 = __halves2half2(silu_mul_f16(g.z, u.z), silu_mul_f16(g.w, u.w));
      uint2 o;
      o.x = *reinterpret_cast<const uint32_t*>(&h0);
      o.y = *reinterpret_cast<const uint32_t*>(&h1);
      *reinterpret_cast<uint2*>(p.y + col0 + c) = o;
    }
    if constexpr (COUT == 1) {  // h tile `cidx` is final: release it to the consumer CTAs whose K chunks overlap it
      __syncthreads();
      if (tid == 0) {
        const int lo = cidx * COLS, hi = min(sg.ncols, lo + COLS) - 1;
        const unsigned readers = (unsigned)((hi / p.chain_rows - lo / p.chain_rows + 1) * p.chain_ctiles);
        __threadfence();
        rc_st_release_gpu(p.chain_ready + cidx, readers);
      }
    }
  } else if constexpr (TP) {
    // ---- fused TP exchange of this column tile -----------------------------------------------------------------
    const int y0 = sg.y_col0 + col0;
    const TpFuse& tp = p.tp;
    __half* mine = tp.part[tp.rank];
    for (int c = tid * 4; c < wcols; c += kRcThreads * 4) {  // 1. my fp16 partial -> my peer-mapped buffer
      const float4 a = ld_cg_f4(p.acc + acc0 + c);
      st_cg_f4(p.acc + acc0 + c, make_float4(0.f, 0.f, 0.f, 0.f));
      const __half2 h0 = __floats2half2_rn(a.x, a.y), h1 = __floats2half2_rn(a.z, a.w);
      uint2 o;
      o.x = *reinterpret_cast<const uint32_t*>(&h0);
      o.y = *reinterpret_cast<const uint32_t*>(&h1);
      *reinterpret_cast<uint2*>(mine + y0 + c) = o;
    }
    __syncthreads();  // the partial's stores happen-before the publishing threads' system-scope release
    if (tid < tp.world) {  // 2. one thread per peer: tell it, then wait for its partial of the same tile
      const unsigned e = tp.epoch[cidx] + 1u;
      __threadfence_system();
      rc_st_release_sys(tp.flags[tid] + tp.rank * kSsqParts + cidx, e);
      const unsigned* theirs = tp.flags[tp.rank] + tid * kSsqParts + cidx;
      const long long t0 = clock64();
      while ((int)(rc_ld_acquire_sys(theirs) - e) < 0) {
        if (clock64() - t0 > 4000000000ll) __trap();  // a missing peer traps instead of hanging the GPU
      }
    }
    __syncthreads();
    if (tid == 0) tp.epoch[cidx] += 1u;
    for (int c = tid * 8; c < wcols; c += kRcThreads * 8) {  // 3. sum in rank order (replicas bit-identical), + residual
      float a[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] = 0.f;
      for (int r = 0; r < tp.world; ++r) {
        const uint4 v = rc_ld_volatile_u4(tp.part[r] + y0 + c);
        const __half2* h2 = reinterpret_cast<const __half2*>(&v);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = __half22float2(h2[j]);
          a[2 * j] += f.x;
          a[2 * j + 1] += f.y;
        }
      }
      uint4 xv = *reinterpret_cast<const uint4*>(p.resid + y0 + c);
      __half2* x2 = reinterpret_cast<__half2*>(&xv);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        x2[j] = __hadd2(x2[j], __floats2half2_rn(a[2 * j], a[2 * j + 1]));  // y rounds to fp16, then the fp16 residual add
        const float2 f = __half22float2(x2[j]);
        ssq = fmaf(f.x, f.x, fmaf(f.y, f.y, ssq));
      }
      *reinterpret_cast<uint4*>(p.y + y0 + c) = xv;
    }
  } else {
    const int y0 = sg.y_col0 + col0;
    for (int c = tid * 4; c < wcols; c += kRcThreads * 4) {
      const float4 a = ld_cg_f4(p.acc + acc0 + c);
      st_cg_f4(p.acc + acc0 + c, make_float4(0.f, 0.f, 0.f, 0.f));
      __half2 h0 = __floats2half2_rn(a.x, a.y), h1 = __floats2half2_rn(a.z, a.w);
      if (p.out_mode == kOutAddResidual) {  // h = x + y in fp16 (gpt-fast/model.py:160-161)
        // (chained twins: the residual stream was written by a grid that may still be running -> read it at L2)
        const uint2 r = (CHAIN != 0) ? __ldcg(reinterpret_cast<const uint2*>(p.resid + y0 + c)) : *reinterpret_cast<const uint2*>(p.resid + y0 + c);
        h0 = __hadd2(*reinterpret_cast<const __half2*>(&r.x), h0);
        h1 = __hadd2(*reinterpret_cast<const __half2*>(&r.y), h1);
      }
      uint2 o;
      o.x = *reinterpret_cast<const uint32_t*>(&h0);
      o.y = *reinterpret_cast<const uint32_t*>(&h1);
      *reinterpret_cast<uint2*>(p.y + y0 + c) = o;
      const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
      ssq = fmaf(f0.x, f0.x, fmaf(f0.y, f0.y, fmaf(f1.x, f1.x, fmaf(f1.y, f1.y, ssq))));
    }
    if constexpr (COUT == 3) {  // QKV tile `cidx` is final: release it to the attention CTAs whose q / k / v slices touch it
      __syncthreads();
      if (tid == 0) {
        const unsigned readers = p.ca.readers[cidx];  // (head, slice) pairs touching this tile: chain_qkv_readers, from the host
        __threadfence();
        rc_st_release_gpu(p.ca.ready_qkv + cidx, readers);
      }
    }
  }
  if (p.ssq_out) {  // this tile's contribution to the next RMSNorm (fixed reduction order: deterministic)
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) ssq += __shfl_xor_sync(0xffffffffu, ssq, d);
    if (lane == 0) s_ssq[warp] = ssq;
    __syncthreads();
    if (tid == 0) {
      float a = 0.f;
#pragma unroll
      for (int w = 0; w < kRcWarps; ++w) a += s_ssq[w];
      p.ssq_out[cidx] = a;
      // every slot of the kSsqParts-long buffer is (re)defined by exactly one tile on every call: slot i by tile i,
      // the unused tail i + T, i + 2T, ... is zeroed -- a previous producer with more tiles leaves nothing stale
      for (int sidx = cidx + p.total_ctiles; sidx < kSsqParts; sidx += p.total_ctiles) p.ssq_out[sidx] = 0.f;
    }
  }
  if constexpr (COUT == 4) {  // this tile of the residual stream (and its sum of squares) is final: count it; the last one
    __syncthreads();          // releases the whole vector to the CTAs of the next RMSNorm consumer
    if (tid == 0) {
      __threadfence();
      const unsigned done = atomicAdd(p.chain_done, 1u) + 1u;
      if ((int)done == p.total_ctiles) {
        *p.chain_done = 0u;
        __threadfence();
        rc_st_release_gpu(p.chain_all_out, (unsigned)p.chain_all_readers);
      }
    }
  }
  if (tid == 0) p.counters[cidx] = 0;
  stamp(7);  // finaliser done (only the last arrival of a tile gets here)
}

// ------------------------------------------------------------------------------------------------------
// Generation 2 (opt-in: teal_set_tuning(TEAL_TUNE_GENERATION, 2); compiled and SASS-checked, NOT yet run on
// hardware -- the default dispatch never selects it).  Same tiling, workspace protocol and epilogue as
// rowchunk_gemv_kernel, so every parity test applies unchanged; what differs is the per-CTA critical path,
// which is what a 3-18 us projection is made of:
//   * survivors are compacted into a per-warp list in SHARED MEMORY (fp16 bits of x | lane offset << 16): the
//     gather loop is one broadcast LDS + one IMAD.WIDE per row instead of ffs / brev / shfl / 64-bit address
//     arithmetic per 16-byte load (~75 -> ~25 instructions per row of a 512-column tile);
//   * the U rows in flight per lane form a ROLLING register ring: row i+U is requested the moment row i has been
//     consumed, instead of U loads / one full DRAM latency / U consumes;
//   * FH: fp16 x fp16 + fp32 in one FHFMA (fma.rn.f32.f16, exact product, one rounding: bit-identical to
//     converting both factors and using FFMA), no conversion instructions;
//   * x, the RMSNorm weights and the sum-of-squares partials are requested together (one L2 round trip, not
//     three); the norm weights -- constants -- before griddepcontrol.wait;
//   * optional speculation (spec_rows > 0): while the CTA waits for its producer (HBM is idle in the producer's
//     tail), the first-wave CTAs pull the first rows of their lane ranges into L2 without knowing which survive.
template <int VEC, int WARPS>
__device__ __forceinline__ void rc2_finish(const RcParams& p, const RcSeg& sg, int ct, int col0, int wcols,
                                           const float2 (&acc)[VEC][4], float (*s_red)[256 * VEC], float* s_ssq, int* s_last) {
  constexpr int kThreadsCta = WARPS * 32;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
#pragma unroll
  for (int v = 0; v < VEC; ++v) {
    float* d = &s_red[warp][lane * 8 + 256 * v];
    *reinterpret_cast<float4*>(d) = make_float4(acc[v][0].x, acc[v][0].y, acc[v][1].x, acc[v][1].y);
    *reinterpret_cast<float4*>(d + 4) = make_float4(acc[v][2].x, acc[v][2].y, acc[v][3].x, acc[v][3].y);
  }
  __syncthreads();
  const int acc0 = sg.acc_col0 + col0;
  for (int c = tid * 4; c < wcols; c += kThreadsCta * 4) {
    float4 a = *reinterpret_cast<const float4*>(&s_red[0][c]);
#pragma unroll
    for (int w = 1; w < WARPS; ++w) {
      const float4 b = *reinterpret_cast<const float4*>(&s_red[w][c]);
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    if (a.x != 0.f || a.y != 0.f || a.z != 0.f || a.w != 0.f) red_add_v4(p.acc + acc0 + c, a);
  }
  __syncthreads();  // REDs of all threads happen-before the barrier; thread 0's release fence is cumulative over them
  const int pair = p.mode == kModePairSilu;
  const int cidx = pair ? (ct - sg.ctile0) : ct;
  if (tid == 0) {
    fence_acq_rel_gpu();
    const unsigned old = atomicAdd(&p.counters[cidx], 1u);
    const int last = ((int)old + 1 == p.target) ? 1 : 0;
    if (last) fence_acq_rel_gpu();
    *s_last = last;
  }
  __syncthreads();
  if (!*s_last) return;

  float ssq = 0.f;
  if (pair) {
    const int g0 = p.seg[0].acc_col0 + col0, u0 = p.seg[1].acc_col0 + col0;
    for (int c = tid * 4; c < wcols; c += kThreadsCta * 4) {
      const float4 g = ld_cg_f4(p.acc + g0 + c), u = ld_cg_f4(p.acc + u0 + c);
      st_cg_f4(p.acc + g0 + c, make_float4(0.f, 0.f, 0.f, 0.f));
      st_cg_f4(p.acc + u0 + c, make_float4(0.f, 0.f, 0.f, 0.f));
      const __half2 h0 = __halves2half2(silu_mul_f16(g.x, u.x), silu_mul_f16(g.y, u.y));
      const __half2 h1 = __halves2half2(silu_mul_f16(g.z, u.z), silu_mul_f16(g.w, u.w));
      uint2 o;
      o.x = *reinterpret_cast<const uint32_t*>(&h0);
      o.y = *reinterpret_cast<const uint32_t*>(&h1);
      *reinterpret_cast<uint2*>(p.y + col0 + c) = o;
    }
  } else {
    const int y0 = sg.y_col0 + col0;
    for (int c = tid * 4; c < wcols; c += kThreadsCta * 4) {
      const float4 a = ld_cg_f4(p.acc + acc0 + c);
      st_cg_f4(p.acc + acc0 + c, make_float4(0.f, 0.f, 0.f, 0.f));
      __half2 h0 = __floats2half2_rn(a.x, a.y), h1 = __floats2half2_rn(a.z, a.w);
      if (p.out_mode == kOutAddResidual) {  // h = x + y in fp16 (gpt-fast/model.py:160-161)
        const uint2 r = *reinterpret_cast<const uint2*>(p.resid + y0 + c);
        h0 = __hadd2(*reinterpret_cast<const __half2*>(&r.x), h0);
        h1 = __hadd2(*reinterpret_cast<const __half2*>(&r.y), h1);
      }
      uint2 o;
      o.x = *reinterpret_cast<const uint32_t*>(&h0);
      o.y = *reinterpret_cast<const uint32_t*>(&h1);
      *reinterpret_cast<uint2*>(p.y + y0 + c) = o;
      const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
      ssq = fmaf(f0.x, f0.x, fmaf(f0.y, f0.y, fmaf(f1.x, f1.x, fmaf(f1.y, f1.y, ssq))));
    }
  }
  if (p.ssq_out) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) ssq += __shfl_xor_sync(0xffffffffu, ssq, d);
    if (lane == 0) s_ssq[warp] = ssq;
    __syncthreads();
    if (tid == 0) {
      float a = 0.f;
#pragma unroll
      for (int w = 0; w < WARPS; ++w) a += s_ssq[w];
      p.ssq_out[cidx] = a;
      for (int sidx = cidx + p.total_ctiles; sidx < kSsqParts; sidx += p.total_ctiles) p.ssq_out[sidx] = 0.f;
    }
  }
  if (tid == 0) p.counters[cidx] = 0;
}

// Warp-wise completion (generation 2, WW): every WARP is an independent worker -- it adds its own partial to the L2
// accumulator straight from registers and arrives at the tile's counter on its own (target = K chunks x warps), so
// there is no shared-memory reduction, no __syncthreads, and the RED -> fence -> counter round trips of one warp overlap
// the other warps' streaming instead of idling the whole CTA.  The last-arriving warp finalises the tile (16 columns
// per lane of a 512-column tile).  Costs WARPS x more RED traffic at L2 (2 KB per 16 KB of weights).
template <int VEC>
__device__ __forceinline__ void rc2_finish_warp(const RcParams& p, const RcSeg& sg, int ct, int col0, int wcols,
                                                const float2 (&acc)[VEC][4]) {
  constexpr int COLS = 256 * VEC;
  const int lane = threadIdx.x & 31;
  const int acc0 = sg.acc_col0 + col0;
#pragma unroll
  for (int v = 0; v < VEC; ++v) {
    const int c = lane * 8 + 256 * v;
    if (c < wcols) {
      const float4 a = make_float4(acc[v][0].x, acc[v][0].y, acc[v][1].x, acc[v][1].y);
      const float4 b = make_float4(acc[v][2].x, acc[v][2].y, acc[v][3].x, acc[v][3].y);
      if (a.x != 0.f || a.y != 0.f || a.z != 0.f || a.w != 0.f) red_add_v4(p.acc + acc0 + c, a);
      if (b.x != 0.f || b.y != 0.f || b.z != 0.f || b.w != 0.f) red_add_v4(p.acc + acc0 + c + 4, b);
    }
  }
  __syncwarp();  // the lanes' REDs happen-before lane 0's cumulative release fence
  const int pair = p.mode == kModePairSilu;
  const int cidx = pair ? (ct - sg.ctile0) : ct;
  int last = 0;
  if (lane == 0) {
    fence_acq_rel_gpu();
    const unsigned old = atomicAdd(&p.counters[cidx], 1u);
    last = ((int)old + 1 == p.target) ? 1 : 0;
    if (last) fence_acq_rel_gpu();
  }
  last = __shfl_sync(0xffffffffu, last, 0);
  __syncwarp();  // lane 0's acquire fence happens-before the other lanes' accumulator reads
  if (!last) return;

  float ssq = 0.f;
  if (pair) {
    const int g0 = p.seg[0].acc_col0 + col0, u0 = p.seg[1].acc_col0 + col0;
    for (int c = lane * 4; c < wcols; c += 128) {
      const float4 g = ld_cg_f4(p.acc + g0 + c), u = ld_cg_f4(p.acc + u0 + c);
      st_cg_f4(p.acc + g0 + c, make_float4(0.f, 0.f, 0.f, 0.f));
      st_cg_f4(p.acc + u0 + c, make_float4(0.f, 0.f, 0.f, 0.f));
      const __half2 h0 = __halves2half2(silu_mul_f16(g.x, u.x), silu_mul_f16(g.y, u.y));
      const __half2 h1 = __halves2half2(silu_mul_f16(g.z, u.z), silu_mul_f16(g.w, u.w));
      uint2 o;
      o.x = *reinterpret_cast<const uint32_t*>(&h0);
      o.y = *reinterpret_cast<const uint32_t*>(&h1);
      *reinterpret_cast<uint2*>(p.y + col0 + c) = o;
    }
  } else {
    const int y0 = sg.y_col0 + col0;
    for (int c = lane * 4; c < wcols; c += 128) {
      const float4 a = ld_cg_f4(p.acc + acc0 + c);
      st_cg_f4(p.acc + acc0 + c, make_float4(0.f, 0.f, 0.f, 0.f));
      __half2 h0 = __floats2half2_rn(a.x, a.y), h1 = __floats2half2_rn(a.z, a.w);
      if (p.out_mode == kOutAddResidual) {  // h = x + y in fp16 (gpt-fast/model.py:160-161)
        const uint2 r = *reinterpret_cast<const uint2*>(p.resid + y0 + c);
        h0 = __hadd2(*reinterpret_cast<const __half2*>(&r.x), h0);
        h1 = __hadd2(*reinterpret_cast<const __half2*>(&r.y), h1);
      }
      uint2 o;
      o.x = *reinterpret_cast<const uint32_t*>(&h0);
      o.y = *reinterpret_cast<const uint32_t*>(&h1);
      *reinterpret_cast<uint2*>(p.y + y0 + c) = o;
      const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
      ssq = fmaf(f0.x, f0.x, fmaf(f0.y, f0.y, fmaf(f1.x, f1.x, fmaf(f1.y, f1.y, ssq))));
    }
  }
  if (p.ssq_out) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) ssq += __shfl_xor_sync(0xffffffffu, ssq, d);
    if (lane == 0) {
      p.ssq_out[cidx] = ssq;
      for (int sidx = cidx + p.total_ctiles; sidx < kSsqParts; sidx += p.total_ctiles) p.ssq_out[sidx] = 0.f;
    }
  }
  if (lane == 0) p.counters[cidx] = 0;
  (void)COLS;
}


// The gather loop of generation 2.  `list` (shared-space address) holds n entries {fp16 bits of x | lane offset << 16};
// row i+U is requested as soon as row i has been consumed (ring slot i % U, statically indexed by the unrolled u).
template <int VEC, int U, bool FH, bool FULL, bool BATCH = false>
__device__ __forceinline__ void rc2_gather(float2 (&acc)[VEC][4], const uint32_t* list, int n, uint64_t wrow, uint32_t ldwb,
                                           uint64_t policy, const bool (&vact)[VEC]) {
  uint4 buf[U][VEC];
  uint32_t ent[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    ent[u] = 0u;
#pragma unroll
    for (int v = 0; v < VEC; ++v) buf[u][v] = make_uint4(0u, 0u, 0u, 0u);
  }
#define RC2_ISSUE(u, i)                                                                                          \
  {                                                                                                              \
    ent[u] = rc_lds_u32(list + (i));                                                                             \
    const uint64_t src = rc_row_address(wrow, ent[u] >> 16, ldwb);                                               \
    _Pragma("unroll") for (int v = 0; v < VEC; ++v) if (FULL || vact[v])                                         \
        buf[u][v] = ld_stream_u4(reinterpret_cast<const char*>(src) + 512 * v, policy);                          \
  }
  if constexpr (BATCH) {  // all requests of a batch are issued before its first row is consumed, nothing is refilled in between
    for (int base = 0; base < n; base += U) {
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (base + u < n) RC2_ISSUE(u, base + u)
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (base + u < n) {
          const unsigned short xb = (unsigned short)(ent[u] & 0xffffu);
#pragma unroll
          for (int v = 0; v < VEC; ++v) {
            if (FULL || vact[v]) {
              if (FH) {
                rc_fhfma2(acc[v][0], buf[u][v].x, xb);
                rc_fhfma2(acc[v][1], buf[u][v].y, xb);
                rc_fhfma2(acc[v][2], buf[u][v].z, xb);
                rc_fhfma2(acc[v][3], buf[u][v].w, xb);
              } else {
                const float xs = __half2float(__ushort_as_half(xb));
                const float2 x2 = make_float2(xs, xs);
                rc_ffma2(acc[v][0], __half22float2(*reinterpret_cast<const __half2*>(&buf[u][v].x)), x2);
                rc_ffma2(acc[v][1], __half22float2(*reinterpret_cast<const __half2*>(&buf[u][v].y)), x2);
                rc_ffma2(acc[v][2], __half22float2(*reinterpret_cast<const __half2*>(&buf[u][v].z)), x2);
                rc_ffma2(acc[v][3], __half22float2(*reinterpret_cast<const __half2*>(&buf[u][v].w)), x2);
              }
            }
          }
        }
      }
    }
    return;
  }
#pragma unroll
  for (int u = 0; u < U; ++u)
    if (u < n) RC2_ISSUE(u, u)
  for (int base = 0; base < n; base += U) {
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = base + u;
      if (i < n) {  // warp-uniform
        const unsigned short xb = (unsigned short)(ent[u] & 0xffffu);
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
          if (FH) {
            rc_fhfma2(acc[v][0], buf[u][v].x, xb);
            rc_fhfma2(acc[v][1], buf[u][v].y, xb);
            rc_fhfma2(acc[v][2], buf[u][v].z, xb);
            rc_fhfma2(acc[v][3], buf[u][v].w, xb);
          } else {
            const float xs = __half2float(__ushort_as_half(xb));
            const float2 x2 = make_float2(xs, xs);
            rc_ffma2(acc[v][0], __half22float2(*reinterpret_cast<const __half2*>(&buf[u][v].x)), x2);
            rc_ffma2(acc[v][1], __half22float2(*reinterpret_cast<const __half2*>(&buf[u][v].y)), x2);
            rc_ffma2(acc[v][2], __half22float2(*reinterpret_cast<const __half2*>(&buf[u][v].z)), x2);
            rc_ffma2(acc[v][3], __half22float2(*reinterpret_cast<const __half2*>(&buf[u][v].w)), x2);
          }
        }
        if (i + U < n) RC2_ISSUE(u, i + U)
      }
    }
  }
#undef RC2_ISSUE
}

template <int VEC, int U, int WARPS, bool FH, bool WW, bool BATCH = false>
__global__ void __launch_bounds__(WARPS * 32, BATCH ? 28 / WARPS : 0) rowchunk_gemv2_kernel(const __grid_constant__ RcParams p) {
  constexpr int COLS = 256 * VEC;
  constexpr int kCap = kRc2MaxPass * 32;
  __shared__ __align__(16) float s_red[WW ? 1 : WARPS][WW ? 4 : COLS];  // warp-wise completion needs no staging
  __shared__ uint32_t s_list[WARPS][kCap];  // the warp's compacted survivors: fp16 bits of x | (lane offset << 16)
  __shared__ float s_ssq[WARPS];
  __shared__ int s_last;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ct = (int)(blockIdx.x % (unsigned)p.total_ctiles);
  const int kc = (int)(blockIdx.x / (unsigned)p.total_ctiles);
  int s = 0;
#pragma unroll
  for (int j = 1; j < kMaxSeg; ++j) s += (j < p.nseg && ct >= p.seg[j].ctile0) ? 1 : 0;
  const RcSeg& sg = p.seg[s];
  const int col0 = (ct - sg.ctile0) * COLS;
  const int wcols = min(COLS, sg.ncols - col0);
  const float thr = sg.thr;
  bool vact[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) vact[v] = lane * 8 + 256 * v < wcols;
  const uint32_t ldwb = (uint32_t)sg.ldw * 2u;  // bytes per weight row (the launcher checks that it fits)
  const uint64_t policy = make_evict_first_policy();

  // ---- the warp's contiguous lane range [kb, ke) of the CTA's K chunk --------------------------------------
  const int kbeg = kc * p.rows_per_cta;
  const int kend = min(p.Z, kbeg + p.rows_per_cta);
  const int kb = min(kend, kbeg + warp * p.lanes_per_warp);
  const int ke = min(kend, kb + p.lanes_per_warp);
  const char* wrow0 = reinterpret_cast<const char*>(sg.w + col0 + lane * 8) + (uint64_t)(uint32_t)kb * ldwb;
  const bool rms = p.in_mode == kInRmsNorm;

  // ---- before the producer is done: constants, and (first wave only) speculative rows into L2 ---------------
  __half nw[kRc2MaxPass];
#pragma unroll
  for (int q = 0; q < kRc2MaxPass; ++q) {
    const int k = kb + q * 32 + lane;
    nw[q] = (rms && k < ke) ? p.norm_w[k] : __ushort_as_half((unsigned short)0);
  }
  if (p.spec_rows > 0 && (int)blockIdx.x < p.spec_ctas && p.mode != kModeSurvivors) {
    const int nrows = min(p.spec_rows, ke - kb);
    if (p.spec_mode == 2) {  // one bulk prefetch per row segment, lane r takes row r
      if (lane < nrows)
        rc_bulk_prefetch_l2(reinterpret_cast<const char*>(sg.w + col0) + (uint64_t)(uint32_t)(kb + lane) * ldwb, (uint32_t)wcols * 2u);
    } else {                 // one prefetch per 128-byte line: a warp-wide instruction covers 32 / LPR rows
      constexpr int LPR = COLS / 64;
      const int line = lane % LPR;
      if (line * 64 < wcols)
        for (int r = lane / LPR; r < nrows; r += 32 / LPR)
          rc_prefetch_l2(reinterpret_cast<const char*>(sg.w + col0 + line * 64) + (uint64_t)(uint32_t)(kb + r) * ldwb);
    }
  }

  griddep_wait();
  griddep_launch_dependents();

  // ---- x lanes and the producer's sum-of-squares partials: all requests in flight together (L2) -------------
  __half xh[kRc2MaxPass];
#pragma unroll
  for (int q = 0; q < kRc2MaxPass; ++q) {
    const int k = kb + q * 32 + lane;
    xh[q] = (k < ke) ? rc_ldcg_half(p.x + k) : __ushort_as_half((unsigned short)0);
  }
  float rs = 1.f;
  if (rms) {
    float t = 0.f;
    for (int i = lane; i < p.n_ssq_in; i += 32) t += __ldcg(p.ssq_in + i);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) t += __shfl_xor_sync(0xffffffffu, t, d);
    rs = rsqrtf(t / (float)p.Z + p.eps);
  }

  // ---- threshold + compaction into the warp's shared-memory list --------------------------------------------
  int n = 0;
#pragma unroll
  for (int q = 0; q < kRc2MaxPass; ++q) {
    const int k = kb + q * 32 + lane;
    if (kb + q * 32 < ke) {  // warp-uniform
      __half xv = xh[q];
      if (rms) {  // fp32 norm, cast to fp16, multiply by the fp16 weight in fp16 (gpt-fast/model.py:286-291)
        xv = __hmul(__float2half_rn(__half2float(xv) * rs), nw[q]);
        if (p.x_out && ct == 0 && k < ke) p.x_out[k] = xv;
      }
      const bool keep = (k < ke) && !(fabsf(__half2float(xv)) <= thr);  // strict >, NaN kept
      const unsigned mask = __ballot_sync(0xffffffffu, keep);
      if (p.keep_out && ct == 0 && k < ke) p.keep_out[k] = keep ? 1 : 0;
      if (keep) s_list[warp][n + __popc(mask & ((1u << lane) - 1u))] = (uint32_t)__half_as_ushort(xv) | ((uint32_t)(q * 32 + lane) << 16);
      n += __popc(mask);
    }
  }
  __syncwarp();
  const int my_nnz = n;
  if (p.mode == kModeSurvivors) n = 0;  // parity hook: thresholds only, no weights

  // ---- gather: U rows in flight per lane, a rolling ring of registers ---------------------------------------
  float2 acc[VEC][4];
#pragma unroll
  for (int v = 0; v < VEC; ++v)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[v][c] = make_float2(0.f, 0.f);
  const uint32_t* list = &s_list[warp][0];
  const uint64_t wrow = reinterpret_cast<uint64_t>(wrow0);
  if (wcols == COLS)  // CTA-uniform: a full tile needs no per-lane column predicates
    rc2_gather<VEC, U, FH, true, BATCH>(acc, list, n, wrow, ldwb, policy, vact);
  else
    rc2_gather<VEC, U, FH, false, BATCH>(acc, list, n, wrow, ldwb, policy, vact);
  if (p.probe && lane == 0) {  // measurement probe: gathered vs dense weight bytes
    atomicAdd(p.probe, (unsigned long long)my_nnz * (unsigned long long)wcols * 2ull);
    if (warp == 0) atomicAdd(p.probe + 1, (unsigned long long)max(0, kend - kbeg) * (unsigned long long)wcols * 2ull);
  }
  if (p.mode == kModeSurvivors) return;
  if constexpr (WW) {
    rc2_finish_warp<VEC>(p, sg, ct, col0, wcols, acc);
  } else {
    rc2_finish<VEC, WARPS>(p, sg, ct, col0, wcols, acc, s_red, s_ssq, &s_last);
  }
}

// Sum-of-squares partials of a vector that did not come out of a GEMV finaliser (embedding row, tests).
__global__ void ssq_kernel(const __half* __restrict__ x, int Z, float* __restrict__ ssq_out, int nparts) {
  griddep_wait();
  // part i covers [i*len, (i+1)*len)
  const int len = (Z + nparts - 1) / nparts;
  const int i = blockIdx.x;
  float t = 0.f;
  for (int k = i * len + threadIdx.x; k < min(Z, (i + 1) * len); k += blockDim.x) {
    const float f = __half2float(x[k]);
    t = fmaf(f, f, t);
  }
  __shared__ float sp[32];
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) t += __shfl_xor_sync(0xffffffffu, t, d);
  if ((threadIdx.x & 31) == 0) sp[threadIdx.x >> 5] = t;
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) a += sp[w];
    ssq_out[i] = a;
  }
  griddep_launch_dependents();
}

// keep-mask -> sorted index list (single CTA; parity hook only)
__global__ void mask_to_indices_kernel(const unsigned char* __restrict__ keep, int Z, int* __restrict__ idx, int* __restrict__ cnt) {
  __shared__ int s_base;
  if (threadIdx.x == 0) s_base = 0;
  __syncthreads();
  for (int k0 = 0; k0 < Z; k0 += blockDim.x) {
    const int k = k0 + threadIdx.x;
    const bool kp = k < Z && keep[k];
    const unsigned m = __ballot_sync(0xffffffffu, kp);
    __shared__ int s_w[32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) s_w[warp] = __popc(m);
    __syncthreads();
    int off = s_base;
    for (int w = 0; w < warp; ++w) off += s_w[w];
    if (kp) idx[off + __popc(m & ((1u << lane) - 1u))] = k;
    __syncthreads();
    if (threadIdx.x == 0) {
      int t = 0;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += s_w[w];
      s_base += t;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) cnt[0] = s_base;
}

// ------------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------------
int launch_rowchunk(GemvParams& gp, void* workspace, size_t ws_bytes, int out_cols_total, cudaStream_t stream,
                    int variant) {
  if (gp.Z <= 0 || gp.Z > (1 << 20)) return set_error(TEAL_ERR_BAD_ARG, "Z out of range");
  RcParams p;
  memset(&p, 0, sizeof(p));
  p.x = gp.x; p.y = gp.y; p.norm_w = gp.norm_w; p.resid = gp.resid; p.x_out = gp.x_out;
  p.ssq_in = gp.ssq_in; p.n_ssq_in = gp.n_ssq_in; p.ssq_out = gp.ssq_out; p.keep_out = gp.keep_out;
  p.probe = bytes_probe();
  p.Z = gp.Z; p.mode = gp.mode; p.in_mode = gp.in_mode; p.out_mode = gp.out_mode; p.eps = gp.eps;
  p.nseg = gp.nseg;
  if (gp.mode != kModeSurvivors) {
    if (!workspace || ws_bytes < workspace_bytes(out_cols_total)) return set_error(TEAL_ERR_WORKSPACE, "workspace too small");
    if ((reinterpret_cast<uintptr_t>(gp.y) & 7)) return set_error(TEAL_ERR_ALIGNMENT, "y must be 8-byte aligned");
    p.counters = reinterpret_cast<unsigned*>(workspace);
    p.acc = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(workspace) + (size_t)kMaxCounters * 4);
  }
  if (gp.in_mode == kInRmsNorm && (!gp.norm_w || !gp.ssq_in || gp.n_ssq_in <= 0))
    return set_error(TEAL_ERR_BAD_ARG, "RMSNorm mode needs norm_w and the producer's sum-of-squares partials");
  if (gp.ssq_out && gp.mode == kModePairSilu) return set_error(TEAL_ERR_UNSUPPORTED, "ssq_out is not defined for the gate/up pair");
  if (gp.out_mode == kOutAddResidual && (!gp.resid || (reinterpret_cast<uintptr_t>(gp.resid) & 7)))
    return set_error(TEAL_ERR_ALIGNMENT, "residual must be non-null and 8-byte aligned");
  struct Variant { int vec, u, warps, rows, gen, ww; void (*kern)(const RcParams); };
#define TEAL_RC(V, UU, W, R) {V, UU, W, R, 1, 0, rowchunk_gemv_kernel<V, UU, W>}
#define TEAL_RC2(V, UU, W, R, FH) {V, UU, W, R, 2, 0, rowchunk_gemv2_kernel<V, UU, W, FH, false>}
#define TEAL_RC2W(V, UU, W, R) {V, UU, W, R, 2, 1, rowchunk_gemv2_kernel<V, UU, W, true, true>}
#define TEAL_RC2B(V, UU, W, R, FH) {V, UU, W, R, 2, 0, rowchunk_gemv2_kernel<V, UU, W, FH, false, true>}
#define TEAL_RCR(V, W, R, D, O) {V, D, W, R, 4, 0, rowchunk_gemv_kernel<V, 1, W, false, false, O, 0, D>}
  static const Variant kVariants[] = {
      TEAL_RC(1, 8, 4, 128),   // 0
      TEAL_RC(2, 4, 4, 128),   // 1
      TEAL_RC(1, 8, 2, 64),    // 2
      TEAL_RC(2, 4, 2, 64),    // 3
      TEAL_RC(1, 16, 4, 128),  // 4
      TEAL_RC(2, 8, 4, 128),   // 5
      TEAL_RC(1, 16, 2, 64),   // 6
      TEAL_RC(2, 8, 2, 64),    // 7
      TEAL_RC(1, 8, 1, 32),    // 8
      TEAL_RC(2, 8, 1, 32),    // 9
      TEAL_RC(1, 8, 4, 256),   // 10
      TEAL_RC(1, 16, 1, 32),   // 11
      // generation 2 (compacted list in shared memory, rolling register ring; not yet run on hardware)
      TEAL_RC2(2, 4, 4, 128, false),  // 12: the shape of variant 1, FFMA2 arithmetic
      TEAL_RC2(2, 4, 4, 128, true),   // 13: ... FHFMA arithmetic (generation-2 default for >= 40 M-element launches)
      TEAL_RC2(1, 8, 2, 64, false),   // 14: the shape of variant 2
      TEAL_RC2(1, 8, 2, 64, true),    // 15: (generation-2 default for the small projections)
      TEAL_RC2(2, 4, 4, 256, true),   // 16: two 32-lane groups per warp (half as many REDs / arrivals)
      TEAL_RC2(1, 8, 4, 128, true),   // 17
      TEAL_RC2(2, 6, 4, 128, true),   // 18: 6 KB in flight per warp
      TEAL_RC2(1, 8, 2, 128, true),   // 19
      // generation 2, warp-wise completion (every warp an independent worker: no smem reduction, no __syncthreads)
      TEAL_RC2W(2, 4, 4, 128),        // 20: 32 lanes per warp
      TEAL_RC2W(1, 8, 2, 64),         // 21
      TEAL_RC2W(2, 4, 4, 256),        // 22: 64 lanes per warp (half the REDs and arrivals)
      TEAL_RC2W(1, 8, 4, 256),        // 23
      // cp.async rings in shared memory (RING instantiations of the generation-1 kernel)
      TEAL_RCR(2, 4, 128, 8, 0),      // 24: the shape of variant 1, 8 x 1 KB per warp
      TEAL_RCR(1, 2, 64, 8, 0),       // 25: the shape of variant 2, 8 x 512 B per warp
      TEAL_RCR(2, 4, 128, 6, 0),      // 26
      TEAL_RCR(1, 4, 128, 8, 0),      // 27
      TEAL_RCR(1, 2, 64, 16, 0),      // 28
      TEAL_RCR(2, 4, 128, 8, 2),      // 29: variant 24 with FHFMA
      // generation 2's lean instruction stream (list + one IMAD.WIDE per row) with generation 1's batch structure
      TEAL_RC2B(2, 4, 4, 128, true),  // 30
      TEAL_RC2B(1, 8, 2, 64, true),   // 31
      TEAL_RC2B(2, 4, 4, 128, false), // 32
      TEAL_RC2B(1, 8, 2, 64, false),  // 33
      TEAL_RC2B(1, 8, 4, 128, true),  // 34
      TEAL_RC2B(2, 4, 2, 64, true),   // 35
      TEAL_RC2B(2, 4, 4, 256, true),  // 36
      TEAL_RC2B(1, 8, 2, 128, true),  // 37
      TEAL_RC2B(1, 8, 2, 256, true),  // 38
      TEAL_RC2B(1, 8, 4, 256, true),  // 39
      TEAL_RC2B(1, 8, 4, 512, true),  // 40
      TEAL_RC2B(1, 8, 1, 64, true),   // 41
      TEAL_RC2B(1, 8, 1, 128, true),  // 42
      TEAL_RC2B(2, 4, 2, 128, true),  // 43
      TEAL_RC2B(2, 4, 2, 256, true),  // 44
      TEAL_RC2B(2, 4, 4, 512, true),  // 45
  };
#undef TEAL_RC
#undef TEAL_RC2
#undef TEAL_RC2W
#undef TEAL_RCR
#undef TEAL_RC2B
  constexpr int kNumVariants = sizeof(kVariants) / sizeof(kVariants[0]);
  if (variant < 0 || variant >= kNumVariants) variant = 0;
  const Variant& var = kVariants[variant];
  const int vec = var.vec;
  const int cols = 256 * vec;
  int rows = var.rows;
  int ctile0 = 0;
  for (int s = 0; s < gp.nseg; ++s) {
    const GemvSeg& g = gp.seg[s];
    if (g.ncols <= 0 || (g.ncols & 7) || (g.ldw & 7) || (g.out_col0 & 7) || (reinterpret_cast<uintptr_t>(g.w) & 15))
      return set_error(TEAL_ERR_ALIGNMENT, "weights: N, ldw, region offsets must be multiples of 8 and 16-byte aligned");
    RcSeg& r = p.seg[s];
    r.w = g.w; r.ldw = g.ldw; r.ncols = g.ncols; r.acc_col0 = g.out_col0;
    r.y_col0 = g.out_col0; r.thr = gp.thr[g.thr_slot];
    r.ctile0 = ctile0; r.nctiles = (g.ncols + cols - 1) / cols;
    ctile0 += r.nctiles;
  }
  if (gp.mode == kModeSurvivors) {  // threshold + keep-mask only: one pseudo column tile
    p.nseg = 1;
    p.seg[0] = RcSeg{};
    p.seg[0].thr = gp.thr[0];
    p.seg[0].ncols = 8;
    p.seg[0].nctiles = 1;
    ctile0 = 1;
  }
  p.total_ctiles = ctile0;
  const bool pair = gp.mode == kModePairSilu;
  if (pair && (gp.nseg != 2 || p.seg[0].nctiles != p.seg[1].nctiles)) return set_error(TEAL_ERR_BAD_ARG, "pair mode needs two equal segments");
  const int ncounters = pair ? p.seg[0].nctiles : ctile0;
  if (ncounters > kMaxCounters) return set_error(TEAL_ERR_UNSUPPORTED, "too many column tiles");
  if (gp.ssq_out && ctile0 > kSsqParts) return set_error(TEAL_ERR_UNSUPPORTED, "ssq_out: more column tiles than TEAL_SSQ_PARTS");
  // Wave quantisation: pick the K-chunk so that the grid is (just under) a whole number of full waves of
  // resident CTAs -- 1792 CTAs on 1036 slots ran 1.73 waves, i.e. the second wave left 27 % of the chip idle.
  int rows_forced = forced_config() ? 0 : tuning(vec == 2 ? TEAL_TUNE_ROWS_WIDE : TEAL_TUNE_ROWS_NARROW);  // experiments only
  if (gp.chain == 2 || gp.chain == 4) rows_forced = gp.chain_rows;  // the producer was told this K chunk (ready[] counts depend on it)
  if ((gp.chain == 1 || gp.chain == 3) && gp.chain_wait_all) rows_forced = var.rows;  // ... and the previous block this grid size
  if (rows_forced > 0) {
    rows = rows_forced;
    if (var.gen == 2 && rows > 32 * kRc2MaxPass * var.warps) rows = 32 * kRc2MaxPass * var.warps;
  } else if (gp.mode != kModeSurvivors && wave_balance_enabled()) {
    static int occ_cache[kNumVariants] = {0};  // (was a fixed [16]: variants >= 16 wrote past it -- the exit-time crash of
                                               // profiles/r01_generations_quick_ab.jsonl's run)
    if (!occ_cache[variant]) {
      int occ = 0;
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, var.kern, var.warps * 32, 0) != cudaSuccess || occ < 1) occ = 1;
      occ_cache[variant] = occ;
    }
    const DeviceInfo& di = device_info();
    const long long slots = (long long)di.sm_count * occ_cache[variant];
    int best_rows = rows;
    long long best_cost = 1ll << 60;
    for (int m = 1; m <= 8; ++m) {
      long long nk = m * slots / ctile0;
      if (nk < 1) continue;
      int r = (int)((gp.Z + nk - 1) / nk);
      if (r < 32 * var.warps) r = 32 * var.warps;
      if (var.gen == 2 && r > 32 * kRc2MaxPass * var.warps) continue;  // a generation-2 warp owns <= kRc2MaxPass lane groups
      if (var.gen == 4 && r > kRcRingList * var.warps) continue;       // a RING warp lists <= kRcRingList lanes
      const long long cost = (r > rows) ? (r - rows) : 2ll * (rows - r);  // prefer the nominal chunk, lean towards larger
      if (cost < best_cost) { best_cost = cost; best_rows = r; }
    }
    rows = best_rows;
  }
  if (var.gen == 4 && rows > kRcRingList * var.warps) rows = kRcRingList * var.warps;
  p.rows_per_cta = rows;
  p.nkchunks = (gp.Z + rows - 1) / rows;
  if (var.gen == 2) {  // (variants of generations 2 and 3: the same kernel template)
    if (rows > 32 * kRc2MaxPass * var.warps) return set_error(TEAL_ERR_UNSUPPORTED, "rowchunk generation 2: K chunk too long");
    for (int s = 0; s < p.nseg && gp.mode != kModeSurvivors; ++s)
      if (p.seg[s].ldw <= 0 || p.seg[s].ldw >= (1ll << 30)) return set_error(TEAL_ERR_UNSUPPORTED, "rowchunk generation 2: ldw out of range");
    p.lanes_per_warp = (rows + var.warps - 1) / var.warps;
    p.spec_rows = tuning(TEAL_TUNE_SPEC_ROWS);
    p.spec_mode = tuning(TEAL_TUNE_SPEC_MODE);
    if (p.spec_rows > 0) {  // only the CTAs that start while the producer drains (the first wave) speculate
      int occ = 0;
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, var.kern, var.warps * 32, 0) != cudaSuccess || occ < 1) occ = 1;
      p.spec_ctas = device_info().sm_count * occ;
    }
  }
  p.target = (pair ? 2 * p.nkchunks : p.nkchunks) * (var.ww ? var.warps : 1);
  const long long grid = (long long)p.total_ctiles * p.nkchunks;
  if (grid > 0x7fffffffll) return set_error(TEAL_ERR_UNSUPPORTED, "grid too large");

  void (*kern)(const RcParams) = var.kern;
  if (gp.chain) {  // flag-chained MLP pair (gate/up -> down): CHAIN twins of the two default variants
    if (grid <= kTimelineMaxCtas) p.dbg_time = next_debug_timeline_slot();  // (null unless teal_set_debug_timeline armed it)
    constexpr int kChainCounter0 = 2048, kChainReady0 = 3072, kChainAcc0 = 65536;  // consumer's own regions of the workspace
    if (variant != 1 && variant != 2) return set_error(TEAL_ERR_UNSUPPORTED, "chained launches: only the default row-chunk variants");
    if ((gp.chain == 1 || gp.chain == 3) && out_cols_total > kChainAcc0)
      return set_error(TEAL_ERR_UNSUPPORTED, "chained producer: its accumulators would reach into the consumer's region");
    if (gp.chain != 3 && (gp.chain_rows < 32 || gp.chain_ctiles < 1 || gp.chain_cols < 64 || gp.tp))
      return set_error(TEAL_ERR_BAD_ARG, "chained launches: bad consumer description");
    p.chain_ready = reinterpret_cast<unsigned*>(workspace) + kChainReady0;
    p.chain_rows = gp.chain_rows; p.chain_ctiles = gp.chain_ctiles; p.chain_cols = gp.chain_cols;
    // whole-output hand-offs between the blocks: cell A = o-proj -> gate/up, cell B = down -> the next layer's QKV
    unsigned* const all_cells = reinterpret_cast<unsigned*>(workspace) + kChainReady0 + 384;
    if (gp.chain == 1 || gp.chain == 3) {
      p.chain_all_in = gp.chain_wait_all ? all_cells + (gp.chain == 1 ? 0 : 1) : nullptr;
    } else if (gp.chain_release_all > 0) {
      p.chain_all_out = all_cells + (gp.chain == 4 ? 0 : 1);
      p.chain_done = all_cells + (gp.chain == 4 ? 2 : 3);
      p.chain_all_readers = gp.chain_release_all;
    }
    const bool wa = gp.chain_wait_all != 0, ra = gp.chain_release_all > 0;
    if (gp.chain == 3) {  // QKV producer of a chained attention block
      if (!gp.ca || pair || gp.mode != kModePlain || cols != gp.ca->cols || ctile0 > 64 || gp.nseg != gp.ca->nseg)
        return set_error(TEAL_ERR_BAD_ARG, "chained QKV: announced tiling");
      p.ca = *gp.ca;
      p.ca.ready_qkv = reinterpret_cast<unsigned*>(workspace) + kChainReady0 + 128;
      p.ca.ready_att = reinterpret_cast<unsigned*>(workspace) + kChainReady0 + 256;
      p.chain_group = 4;
      p.chain_rot = chain_tile_of(p.ca, p.ca.n_head * p.ca.hd);  // first k tile
      kern = variant == 1 ? (wa ? rowchunk_gemv_kernel<2, 4, 4, false, false, 0, 23> : rowchunk_gemv_kernel<2, 4, 4, false, false, 0, 3>)
                          : (wa ? rowchunk_gemv_kernel<1, 8, 2, false, false, 0, 23> : rowchunk_gemv_kernel<1, 8, 2, false, false, 0, 3>);
    } else if (gp.chain == 1) {
      if (!pair || cols != gp.chain_cols || ncounters > kMaxCounters - kChainReady0 || ncounters > kChainCounter0)
        return set_error(TEAL_ERR_BAD_ARG, "chained producer: gate/up pair whose tile width is the announced one");
      p.chain_group = 8;  // order positions per group = 4 tile pairs: the first h tiles are final after 1/7th of an 8B launch
      kern = variant == 1 ? (wa ? rowchunk_gemv_kernel<2, 4, 4, false, false, 0, 21> : rowchunk_gemv_kernel<2, 4, 4, false, false, 0, 1>)
                          : (wa ? rowchunk_gemv_kernel<1, 8, 2, false, false, 0, 21> : rowchunk_gemv_kernel<1, 8, 2, false, false, 0, 1>);
    } else {
      if (gp.chain == 4) p.chain_ready = reinterpret_cast<unsigned*>(workspace) + kChainReady0 + 256;  // ready_att: one cell per head
      const int ntiles_in = (gp.Z + gp.chain_cols - 1) / gp.chain_cols;
      if (gp.mode != kModePlain || gp.in_mode != kInPlain || ntiles_in > kMaxCounters - kChainReady0 || ctile0 > kChainReady0 - kChainCounter0 ||
          ws_bytes < (size_t)kMaxCounters * 4 + ((size_t)kChainAcc0 + out_cols_total) * 4 + (size_t)kSsqParts * 4 || gp.chain_ctiles != ctile0)
        return set_error(TEAL_ERR_BAD_ARG, "chained consumer: plain-input projection, announced tiling, workspace for a second accumulator region");
      p.counters += kChainCounter0;  // the producer grid is still running: its counters / accumulators are not ours
      p.acc += kChainAcc0;
      kern = variant == 1 ? (ra ? rowchunk_gemv_kernel<2, 4, 4, false, false, 0, 14> : rowchunk_gemv_kernel<2, 4, 4, false, false, 0, 10>)
                          : (ra ? rowchunk_gemv_kernel<1, 8, 2, false, false, 0, 14> : rowchunk_gemv_kernel<1, 8, 2, false, false, 0, 10>);
    }
  } else if (gp.tp) {  // row-parallel projection with the TP exchange fused into the tile finaliser (TP twins of the defaults)
    if (variant != 1 && variant != 2) return set_error(TEAL_ERR_UNSUPPORTED, "fused TP exchange: only the default row-chunk variants");
    if (gp.mode != kModePlain || gp.out_mode != kOutAddResidual || gp.nseg != 1 || (gp.seg[0].ncols & 7) || (reinterpret_cast<uintptr_t>(gp.y) & 15) ||
        (reinterpret_cast<uintptr_t>(gp.resid) & 15))
      return set_error(TEAL_ERR_BAD_ARG, "fused TP exchange: one region, residual output, 16-byte aligned vectors");
    if (gp.tp->world < 1 || gp.tp->world > kMaxTpWorld || gp.tp->rank < 0 || gp.tp->rank >= gp.tp->world || !gp.tp->epoch ||
        gp.tp->world > var.warps * 32 || ctile0 > kSsqParts)
      return set_error(TEAL_ERR_BAD_ARG, "fused TP exchange: bad rank / world / tile count");
    for (int r = 0; r < gp.tp->world; ++r)
      if (!gp.tp->part[r] || !gp.tp->flags[r] || (reinterpret_cast<uintptr_t>(gp.tp->part[r]) & 15))
        return set_error(TEAL_ERR_ALIGNMENT, "fused TP exchange: null / unaligned peer pointer");
    p.tp = *gp.tp;
    kern = variant == 1 ? rowchunk_gemv_kernel<2, 4, 4, false, true> : rowchunk_gemv_kernel<1, 8, 2, false, true>;
  } else if (const int opts = forced_config() ? 0 : tuning(TEAL_TUNE_V1_OPTS); opts && (variant == 1 || variant == 2)) {
    // one experimental change at a time on the validated structure (see the OPT template parameter)
    using K = void (*)(const RcParams);
#define TEAL_OPT_ROW(V, UU, W) {nullptr, rowchunk_gemv_kernel<V, UU, W, false, false, 1>, rowchunk_gemv_kernel<V, UU, W, false, false, 2>, \
                                rowchunk_gemv_kernel<V, UU, W, false, false, 3>, rowchunk_gemv_kernel<V, UU, W, false, false, 4>,          \
                                rowchunk_gemv_kernel<V, UU, W, false, false, 5>, rowchunk_gemv_kernel<V, UU, W, false, false, 6>,          \
                                rowchunk_gemv_kernel<V, UU, W, false, false, 7>, rowchunk_gemv_kernel<V, UU, W, false, false, 10>,         \
                                rowchunk_gemv_kernel<V, UU, W, false, false, 14>, rowchunk_gemv_kernel<V, UU, W, false, false, 8>,        \
                                rowchunk_gemv_kernel<V, UU, W, false, false, 12>}
    static const K kOpt[2][12] = {TEAL_OPT_ROW(2, 4, 4), TEAL_OPT_ROW(1, 8, 2)};
#undef TEAL_OPT_ROW
    if (opts == 4 || opts == 5 || opts == 6 || opts == 7 || opts == 9 || opts == 11)
      for (int s = 0; s < p.nseg && gp.mode != kModeSurvivors; ++s)
        if (p.seg[s].ldw <= 0 || p.seg[s].ldw >= (1ll << 30)) return set_error(TEAL_ERR_UNSUPPORTED, "32-bit row offsets: ldw out of range");
    kern = kOpt[variant - 1][opts];  // 8, 9 = FHFMA + {no L2 hint, no hint + 32-bit offsets}; 10, 11 = the same with conversions + FFMA2
  } else if ((variant == 1 || variant == 2) && grid <= kTimelineMaxCtas) {  // profiling: the two default variants have a TIMED twin
    if (long long* slot = next_debug_timeline_slot()) {           // teal_set_debug_timeline: one slot per launch, in launch order
      p.dbg_time = slot;
      kern = variant == 1 ? rowchunk_gemv_kernel<2, 4, 4, true> : rowchunk_gemv_kernel<1, 8, 2, true>;
    }
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(var.warps * 32);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, p);
  if (e != cudaSuccess) return set_cuda_error(e, "rowchunk_gemv launch");
  return TEAL_OK;
}

int launch_ssq(const __half* x, int Z, float* ssq_out, int nparts, cudaStream_t stream) {
  if (!x || !ssq_out || Z <= 0 || nparts <= 0 || nparts > 64) return set_error(TEAL_ERR_BAD_ARG, "ssq: bad argument");
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(nparts);
  cfg.blockDim = dim3(256);
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  cudaError_t e = cudaLaunchKernelEx(&cfg, ssq_kernel, x, Z, ssq_out, nparts);
  if (e != cudaSuccess) return set_cuda_error(e, "ssq launch");
  return TEAL_OK;
}

int launch_mask_to_indices(const unsigned char* keep, int Z, int* idx, int* cnt, cudaStream_t stream) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(1);
  cfg.blockDim = dim3(1024);
  cfg.stream = stream;
  cudaError_t e = cudaLaunchKernelEx(&cfg, mask_to_indices_kernel, keep, Z, idx, cnt);
  if (e != cudaSuccess) return set_cuda_error(e, "mask_to_indices launch");
  return TEAL_OK;
}

}  // namespace teal
