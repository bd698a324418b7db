// sparse_gemv: fused [RMSNorm ->] threshold -> survivor compaction in shared memory -> gather-GEMV
// [-> +residual | SiLU(gate)*up | TP exchange] for sm_100a.  The one production kernel of the library: all seven
// projections of a decoder layer go through it (kernels/sparse_gemv.py:45-83 splitk_sparse_gemv_kernel + its Y.zero_()
// pre-hook, :147-194 qkv_kernel of the reference; thresholds are selected per column region exactly).
//
// Shape of the kernel, and the B200 measurements behind it (profiles/r02_*):
//   * a batch-1 projection moves 17-117 MB, i.e. it lasts 3-18 us at HBM speed: every latency that is serialised
//     inside the kernel counts.  CTAs are SMALL (2-4 warps, 72 registers, 7-14 resident per SM) so that the SMs
//     always hold CTAs in different phases, and a CTA owns one 256/512-column tile x one K chunk of the input lanes;
//     the K chunk is sized so that the whole grid is (at most) one wave of resident CTAs.
//   * per warp: its lanes of x -> (RMSNorm scale) -> |x| > thr -> ballot -> the survivors are COMPACTED into a list
//     {fp16 x, local row} in shared memory.  The gather loop then costs one broadcast LDS + one IMAD.WIDE per row,
//     one coalesced LDG.128 (L1 no-allocate, L2 evict-first) per 512 B of the row and one FHFMA (fp16 x fp16 + fp32,
//     exact product) per weight: ~27 SASS instructions per 1 KB row instead of ~97 with per-row ffs / shuffle / 64-bit
//     multiplies / conversions.  That matters because a warp that is busy consuming has no loads in flight: with the
//     lean loop the warps spend their time waiting on HBM, which is what keeps HBM busy (+14 % tokens/s).
//   * loads are issued in BATCHES of U rows per lane (all requests of a batch before its first row is consumed).
//     A register ring refilled row by row serialises on the load scoreboards (measured: 2x slower), cp.async rings
//     in shared memory cost CTAs per SM and were slower too (profiles/r02_bench_gemv_ring.jsonl).
//   * split-K reduction: warp partials -> shared memory -> one fp32 vector RED per 4 columns into an L2-resident
//     accumulator; an arrival counter per tile; the last CTA converts the tile to fp16, applies the epilogue, leaves
//     the tile's sum of squares for the next RMSNorm and re-zeroes accumulator + counter (no memset launch anywhere).
//     A thread-block-cluster / DSMEM reduction was built and measured (tools/micro/cluster_gemv_bench.cu): it needs
//     fat CTAs to cover Z inside one cluster and two ~0.7 us cluster barriers, and lost on every shape.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "common.cuh"

namespace teal {

namespace {

// streaming 16-byte load: read-only path, no L1 allocation, L2 evict-first (a weight row is read once per token)
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
// f32 += f16 * f16 (FHFMA): the product of two fp16 values is exact in fp32, so this is fma(float(w), float(x), acc)
__device__ __forceinline__ void fhfma2(float2& acc, uint32_t w2, unsigned short xh) {
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
__device__ __forceinline__ void prefetch_l2(const void* p) { (void)*reinterpret_cast<const volatile char*>(p); }
__device__ __forceinline__ __half ldcg_half(const __half* p) { return __ushort_as_half(__ldcg(reinterpret_cast<const unsigned short*>(p))); }
__device__ __forceinline__ uint32_t lds_u32(const uint32_t* p) { return *reinterpret_cast<const volatile uint32_t*>(p); }
__device__ __forceinline__ uint64_t row_address(uint64_t base, uint32_t row, uint32_t row_bytes) { return base + (uint64_t)row * row_bytes; }
__device__ __forceinline__ void st_word_sys(uint2* p, uint2 v) {
  __atomic_store_n(reinterpret_cast<unsigned long long*>(p), (unsigned long long)v.x | ((unsigned long long)v.y << 32), __ATOMIC_RELEASE);
}
__device__ __forceinline__ uint2 ld_word_sys(const uint2* p) {
  ::emu::yield();
  const unsigned long long w = __atomic_load_n(reinterpret_cast<const unsigned long long*>(p), __ATOMIC_ACQUIRE);
  return make_uint2((unsigned)w, (unsigned)(w >> 32));
}
#else
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ __half ldcg_half(const __half* p) {
  unsigned short v;
  asm volatile("ld.global.cg.u16 %0, [%1];" : "=h"(v) : "l"(p) : "memory");
  return __ushort_as_half(v);
}
// volatile asm on purpose: the compiler must neither duplicate the address arithmetic per 16-byte load nor move the
// list read across the weight loads
__device__ __forceinline__ uint32_t lds_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
  return v;
}
__device__ __forceinline__ uint64_t row_address(uint64_t base, uint32_t row, uint32_t row_bytes) {
  uint64_t a;
  asm volatile("mad.wide.u32 %0, %1, %2, %3;" : "=l"(a) : "r"(row), "r"(row_bytes), "l"(base));
  return a;
}
// the fused TP exchange's 8-byte {data, epoch} words (TP instantiations): one store / one load each, never split
__device__ __forceinline__ void st_word_sys(uint2* p, uint2 v) {
  asm volatile("st.volatile.global.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(v.x), "r"(v.y) : "memory");
}
__device__ __forceinline__ uint2 ld_word_sys(const uint2* p) {
  uint2 v;
  asm volatile("ld.volatile.global.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
  return v;
}
#endif

// The gather loop of one warp: `n` list entries {fp16 x | local row << 16}, rows of `wrow` (the lane's 16-byte column piece
// of the warp's first lane), U rows x VEC vectors in flight per lane.  FULL: the tile has all its columns (no per-lane
// column predicates).
template <int VEC, int U, bool FULL>
__device__ __forceinline__ void gather(float2 (&acc)[VEC][4], const uint32_t* list, int n, uint64_t wrow, uint32_t ldwb,
                                       uint64_t policy, const bool (&vact)[VEC]) {
  uint4 buf[U][VEC];
  uint32_t ent[U];
  for (int base = 0; base < n; base += U) {
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (base + u < n) {  // warp-uniform
        ent[u] = lds_u32(list + base + u);
        const uint64_t src = row_address(wrow, ent[u] >> 16, ldwb);
#pragma unroll
        for (int v = 0; v < VEC; ++v)
          if (FULL || vact[v]) buf[u][v] = ld_stream_u4(reinterpret_cast<const char*>(src) + 512 * v, policy);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (base + u < n) {
        const unsigned short xb = (unsigned short)(ent[u] & 0xffffu);
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
          if (FULL || vact[v]) {
            fhfma2(acc[v][0], buf[u][v].x, xb);
            fhfma2(acc[v][1], buf[u][v].y, xb);
            fhfma2(acc[v][2], buf[u][v].z, xb);
            fhfma2(acc[v][3], buf[u][v].w, xb);
          }
        }
      }
    }
  }
}

}  // namespace

// VEC: 16-byte vectors per lane per row (column tile = 256 * VEC); U: rows in flight per lane; WARPS: warps per CTA, each
// owns a contiguous range of p.lanes_per_warp input lanes of the CTA's K chunk.
// TP: the row-parallel projections under tensor parallelism (o / down, gpt-fast/tp.py:117,130) with the exchange step
// (tp.py:120-121,139-140) fused into the tile finaliser -- see TpFuse.
// TIMED: profiling twin (teal_set_debug_timeline): thread 0 of every CTA records globaltimer stamps of its phases;
// tools/gemv_timeline.py turns them into the head / gather / tail picture of a launch.
template <int VEC, int U, int WARPS, bool TP, bool TIMED>
__global__ void __launch_bounds__(WARPS * 32, 28 / WARPS) sparse_gemv_kernel(const __grid_constant__ KParams p) {
  constexpr int COLS = 256 * VEC;
  constexpr int NT = WARPS * 32;
  constexpr int kPass = 4;  // 32-lane groups whose x / norm-weight requests are in flight together
  __shared__ __align__(16) float s_red[WARPS][COLS];
  __shared__ uint32_t s_list[WARPS][kListCap];  // the warp's compacted survivors: fp16 bits of x | (local row << 16)
  __shared__ float s_ssq[WARPS];
  __shared__ int s_last;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  long long* const dbg = TIMED ? p.dbg_time + (size_t)blockIdx.x * 16 : nullptr;
  auto stamp = [&](int i) {
    if constexpr (TIMED) {
      if (tid == 0) dbg[i] = globaltimer_ns();
    }
  };
  stamp(0);
  // ---- which (column tile, K chunk) is mine: the column tile varies fastest, so CTAs that run together read
  // neighbouring pieces of the same weight rows -----------------------------------------------------------------
  const int ct = (int)(blockIdx.x % (unsigned)p.total_ctiles);
  const int kc = (int)(blockIdx.x / (unsigned)p.total_ctiles);
  int s = 0;
#pragma unroll
  for (int j = 1; j < kMaxSeg; ++j) s += (j < p.nseg && ct >= p.seg[j].ctile0) ? 1 : 0;
  const KSeg& sg = p.seg[s];
  const int col0 = (ct - sg.ctile0) * COLS;
  const int wcols = min(COLS, sg.ncols - col0);
  const float thr = sg.thr;
  bool vact[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) vact[v] = lane * 8 + 256 * v < wcols;
  const uint32_t ldwb = (uint32_t)sg.ldw * 2u;  // bytes per weight row (the launcher checks that it fits)
  const uint64_t policy = make_evict_first_policy();

  // ---- the warp's contiguous lane range [kb, ke) of the CTA's K chunk ---------------------------------------------
  const int kbeg = kc * p.rows_per_cta;
  const int kend = min(p.Z, kbeg + p.rows_per_cta);
  const int kb = min(kend, kbeg + warp * p.lanes_per_warp);
  const int ke = min(kend, kb + p.lanes_per_warp);
  const uint64_t wrow = reinterpret_cast<uint64_t>(sg.w + col0 + lane * 8) + (uint64_t)(uint32_t)kb * ldwb;
  const bool rms = p.in_mode == kInRmsNorm;

  // ---- before the producer is done: the norm weights are constants -- the first pass's into registers, the lines of
  // the later passes towards L2 -----------------------------------------------------------------------------------------
  __half xh[kPass], nw[kPass];
  auto load_nw = [&](int q0) {
#pragma unroll
    for (int q = 0; q < kPass; ++q) {
      const int k = kb + (q0 + q) * 32 + lane;
      nw[q] = (rms && k < ke) ? p.norm_w[k] : __ushort_as_half((unsigned short)0);
    }
  };
  auto load_x = [&](int q0) {
#pragma unroll
    for (int q = 0; q < kPass; ++q) {
      const int k = kb + (q0 + q) * 32 + lane;
      xh[q] = (k < ke) ? ldcg_half(p.x + k) : __ushort_as_half((unsigned short)0);
    }
  };
  // kInGateUpAcc: the lanes are h = silu(gate) * up of the producer's deferred fp32 sums (gpt-fast/model.py:259), rounded
  // exactly as the pair finaliser rounds them; all requests of a pass in flight before the first is used
  auto load_x_gate_up = [&](int q0) {
    float g[kPass], u[kPass];
#pragma unroll
    for (int q = 0; q < kPass; ++q) {
      const int k = kb + (q0 + q) * 32 + lane;
      g[q] = (k < ke) ? __ldcg(p.in_acc + k) : 0.f;
      u[q] = (k < ke) ? __ldcg(p.in_acc + p.in_acc_stride + k) : 0.f;
    }
#pragma unroll
    for (int q = 0; q < kPass; ++q) xh[q] = silu_mul_f16(g[q], u[q]);
  };
  const bool from_gate_up = p.in_mode == kInGateUpAcc;
  load_nw(0);
  if (rms && kb + kPass * 32 + lane * 64 < ke) prefetch_l2(p.norm_w + kb + kPass * 32 + lane * 64);

  griddep_wait();               // PDL: everything above overlapped the previous kernel's tail
  griddep_launch_dependents();  // the next kernel may move in behind us and wait for our completion
  stamp(1);

  // ---- x lanes and the producer's sum-of-squares partials: all requests of the first pass in flight together ------
  if (from_gate_up) load_x_gate_up(0); else load_x(0);
  if (p.clear) {  // deferred buffers whose readers finished with the producer we just waited for: my slice back to zero
    const int c1 = min(p.clear_n, (int)(blockIdx.x + 1) * p.clear_per_cta);
    for (int c = (int)blockIdx.x * p.clear_per_cta + tid * 4; c < c1; c += NT * 4) st_cg_f4(p.clear + c, make_float4(0.f, 0.f, 0.f, 0.f));
  }
  float rs = 1.f;
  if (rms) {  // gpt-fast/model.py:286-291
    float t = 0.f;
    const float* ssq = p.ssq_in + (size_t)((blockIdx.x * WARPS + warp) & (kSsqReplicas - 1)) * kSsqParts;  // my replica
    static_assert(kSsqParts == 64, "two requests per lane, both in flight together");
    {
      const float a = __ldcg(ssq + lane), b = __ldcg(ssq + lane + 32);
      t = a + b;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) t += __shfl_xor_sync(0xffffffffu, t, d);
    rs = rsqrtf(t / (float)p.Z + p.eps);
  }

  // ---- threshold + compaction into the warp's shared-memory list ---------------------------------------------------
  uint32_t* const list = &s_list[warp][0];
  int n = 0;
  const int ngroups = (ke - kb + 31) >> 5;
  for (int q0 = 0; q0 < ngroups; q0 += kPass) {
    if (q0 > 0) {
      load_nw(q0);
      if (from_gate_up) load_x_gate_up(q0); else load_x(q0);
    }
#pragma unroll
    for (int q = 0; q < kPass; ++q) {
      if (q0 + q < ngroups) {  // warp-uniform
        const int k = kb + (q0 + q) * 32 + lane;
        __half xv = xh[q];
        if (rms) {  // fp32 norm, cast to fp16, multiply by the fp16 weight in fp16
          xv = __hmul(__float2half_rn(__half2float(xv) * rs), nw[q]);
          if (p.x_out && ct == 0 && k < ke) p.x_out[k] = xv;
        }
        const bool keep = (k < ke) && !(fabsf(__half2float(xv)) <= thr);  // strict >, NaN kept (it must poison y like mask*x)
        const unsigned mask = __ballot_sync(0xffffffffu, keep);
        if (p.keep_out && ct == 0 && k < ke) p.keep_out[k] = keep ? 1 : 0;
        if (keep) list[n + __popc(mask & ((1u << lane) - 1u))] = (uint32_t)__half_as_ushort(xv) | ((uint32_t)((q0 + q) * 32 + lane) << 16);
        n += __popc(mask);
      }
    }
  }
  __syncwarp();
  const int my_nnz = n;
  if (p.mode == kModeSurvivors) n = 0;  // parity hook: thresholds only, no weights
  stamp(2);

  // ---- gather -----------------------------------------------------------------------------------------------------
  float2 acc[VEC][4];
#pragma unroll
  for (int v = 0; v < VEC; ++v)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[v][c] = make_float2(0.f, 0.f);
  if (wcols == COLS)  // CTA-uniform: a full tile needs no per-lane column predicates
    gather<VEC, U, true>(acc, list, n, wrow, ldwb, policy, vact);
  else
    gather<VEC, U, false>(acc, list, n, wrow, ldwb, policy, vact);
  stamp(4);
  if (p.probe && lane == 0) {  // measurement probe: gathered vs dense weight bytes
    atomicAdd(p.probe, (unsigned long long)my_nnz * (unsigned long long)wcols * 2ull);
    if (warp == 0) atomicAdd(p.probe + 1, (unsigned long long)max(0, kend - kbeg) * (unsigned long long)wcols * 2ull);
  }
  if (p.mode == kModeSurvivors) return;

  // ---- CTA partial: warp partials -> shared memory -> fp32 vector RED into the L2 accumulator --------------------
#pragma unroll
  for (int v = 0; v < VEC; ++v) {
    float* d = &s_red[warp][lane * 8 + 256 * v];
    *reinterpret_cast<float4*>(d) = make_float4(acc[v][0].x, acc[v][0].y, acc[v][1].x, acc[v][1].y);
    *reinterpret_cast<float4*>(d + 4) = make_float4(acc[v][2].x, acc[v][2].y, acc[v][3].x, acc[v][3].y);
  }
  __syncthreads();
  const int acc0 = sg.acc_col0 + col0;
  for (int c = tid * 4; c < wcols; c += NT * 4) {
    float4 a = *reinterpret_cast<const float4*>(&s_red[0][c]);
#pragma unroll
    for (int w = 1; w < WARPS; ++w) {
      const float4 b = *reinterpret_cast<const float4*>(&s_red[w][c]);
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    if (a.x != 0.f || a.y != 0.f || a.z != 0.f || a.w != 0.f) red_add_v4(p.acc + acc0 + c, a);
  }
  if (p.out_mode == kOutDeferred) {  // the consumer kernel rounds the sums (grid completion makes them visible to it)
    if constexpr (TIMED) {
      if (tid == 0) {
        unsigned smid = blockIdx.x;
#ifndef TEAL_EMU
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
#endif
        dbg[5] = dbg[6] = globaltimer_ns();  // no arrival, no finaliser: [6] = [5], [7] stays 0
        dbg[7] = 0;
        dbg[8] = smid;
        dbg[9] = my_nnz;
        dbg[10] = 0;
      }
    }
    return;
  }
  // ---- arrival counter: the REDs of all threads happen-before the barrier, thread 0's release fence is cumulative
  // over them; the last arrival acquires and finalises the tile ------------------------------------------------------
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
    if constexpr (TIMED) {
      unsigned smid = blockIdx.x;
#ifndef TEAL_EMU
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
#endif
      dbg[6] = globaltimer_ns();
      dbg[7] = 0;
      dbg[8] = smid;
      dbg[9] = my_nnz;
      dbg[10] = last;
    }
  }
  __syncthreads();
  if (!s_last) return;

  float ssq = 0.f;
  if (pair) {
    const int g0 = p.seg[0].acc_col0 + col0, u0 = p.seg[1].acc_col0 + col0;
    for (int c = tid * 4; c < wcols; c += NT * 4) {
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
  } else if constexpr (TP) {
    // ---- fused TP exchange of this column tile (push, 8-byte {2 x fp16, epoch} words: see TpFuse) --------------------
    const int y0 = sg.y_col0 + col0;
    const TpFuse& tp = p.tp;
    const unsigned e = tp.epoch[cidx] + 1u;
    for (int c = tid * 2; c < wcols; c += NT * 2) {  // 1. my partial of the tile -> every rank's slot for me
      float2 a;
      a.x = __ldcg(p.acc + acc0 + c);
      a.y = __ldcg(p.acc + acc0 + c + 1);
      __stcg(p.acc + acc0 + c, 0.f);
      __stcg(p.acc + acc0 + c + 1, 0.f);
      const __half2 h = __floats2half2_rn(a.x, a.y);
      const uint2 word = make_uint2(*reinterpret_cast<const uint32_t*>(&h), e);
      const size_t slot = ((size_t)tp.rank * tp.n + y0 + c) >> 1;
      for (int q = 0; q < tp.world; ++q) st_word_sys(tp.recv[q] + slot, word);
    }
    for (int c = tid * 2; c < wcols; c += NT * 2) {  // 2. the same column pairs of every source, in rank order
      float2 a = make_float2(0.f, 0.f);
      for (int r = 0; r < tp.world; ++r) {
        const uint2* src = tp.recv[tp.rank] + (((size_t)r * tp.n + y0 + c) >> 1);
        uint2 w = ld_word_sys(src);
        if (w.y != e) {
          const long long t0 = clock64();
          while ((w = ld_word_sys(src)).y != e) {
            if (clock64() - t0 > 4000000000ll) __trap();  // a missing peer traps instead of hanging the GPU
          }
        }
        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w.x));
        a.x += f.x;
        a.y += f.y;
      }
      __half2 xv = *reinterpret_cast<const __half2*>(p.resid + y0 + c);
      xv = __hadd2(xv, __floats2half2_rn(a.x, a.y));  // y rounds to fp16, then the fp16 residual add
      *reinterpret_cast<__half2*>(p.y + y0 + c) = xv;
      const float2 f = __half22float2(xv);
      ssq = fmaf(f.x, f.x, fmaf(f.y, f.y, ssq));
    }
    __syncthreads();  // every thread has read the epoch
    if (tid == 0) tp.epoch[cidx] = e;
  } else {
    const int y0 = sg.y_col0 + col0;
    for (int c = tid * 4; c < wcols; c += NT * 4) {
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
  if (p.ssq_out) {  // this tile's contribution to the next RMSNorm (fixed reduction order: deterministic)
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) ssq += __shfl_xor_sync(0xffffffffu, ssq, d);
    if (lane == 0) s_ssq[warp] = ssq;
    __syncthreads();
    if (tid < kSsqReplicas) {  // one thread per replica of the partials buffer
      float a = 0.f;
#pragma unroll
      for (int w = 0; w < WARPS; ++w) a += s_ssq[w];
      float* rep = p.ssq_out + tid * kSsqParts;
      rep[cidx] = a;
      // every slot of a replica is (re)defined by exactly one tile on every call: slot i by tile i, the unused tail
      // i + T, i + 2T, ... is zeroed -- a previous producer with more tiles leaves nothing stale
      for (int sidx = cidx + p.total_ctiles; sidx < kSsqParts; sidx += p.total_ctiles) rep[sidx] = 0.f;
    }
  }
  if (tid == 0) p.counters[cidx] = 0;
  stamp(7);  // finaliser done (only the last arrival of a tile gets here)
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
    sp[0] = a;
  }
  __syncthreads();
  if (threadIdx.x < kSsqReplicas) ssq_out[threadIdx.x * kSsqParts + i] = sp[0];
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

// Independent slow cross-check (one warp per 256 columns, no K split, no workspace): tests triage with it.
__global__ void sparse_gemv_simple_kernel(const __half* __restrict__ x, const __half* __restrict__ wt, __half* __restrict__ y, int Z,
                                          int N, long long ldw, float thr) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int col = warp * 256 + lane * 8;
  if (col >= N) return;
  float acc[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) acc[c] = 0.f;
  for (int m = 0; m < Z; ++m) {
    const float xv = __half2float(x[m]);
    if (!(fabsf(xv) <= thr)) {
      const uint4 v = *reinterpret_cast<const uint4*>(wt + (size_t)m * ldw + col);
      const __half2* h2 = reinterpret_cast<const __half2*>(&v);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __half22float2(h2[j]);
        acc[2 * j] = fmaf(f.x, xv, acc[2 * j]);
        acc[2 * j + 1] = fmaf(f.y, xv, acc[2 * j + 1]);
      }
    }
  }
#pragma unroll
  for (int c = 0; c < 8; ++c) y[col + c] = __float2half_rn(acc[c]);
}

// ------------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------------
namespace {

struct Variant {
  int vec, u, warps;
  void (*kern)(const KParams);
  void (*kern_tp)(const KParams);
  void (*kern_timed)(const KParams);
};
#define TEAL_V(V, UU, W) {V, UU, W, sparse_gemv_kernel<V, UU, W, false, false>, sparse_gemv_kernel<V, UU, W, true, false>, \
                          sparse_gemv_kernel<V, UU, W, false, true>}
const Variant kVariants[] = {
    TEAL_V(1, 8, 2),  // 0: 256-column tiles, 2 warps -- the default for every projection measured (profiles/r02_*gen5*)
    TEAL_V(1, 8, 4),  // 1
    TEAL_V(2, 4, 4),  // 2: 512-column tiles
    TEAL_V(2, 4, 2),  // 3
    TEAL_V(1, 8, 1),  // 4
};
#undef TEAL_V
constexpr int kNumVariants = sizeof(kVariants) / sizeof(kVariants[0]);

int occupancy_of(int variant) {
  static thread_local int cache[kNumVariants] = {0};
  if (!cache[variant]) {
    int occ = 0;
    const Variant& v = kVariants[variant];
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, v.kern, v.warps * 32, 0) != cudaSuccess || occ < 1) occ = 1;
    cache[variant] = occ;
  }
  return cache[variant];
}

}  // namespace

int gemv_num_variants() { return kNumVariants; }

int launch_sparse_gemv(const GemvParams& gp, void* workspace, size_t ws_bytes, int out_cols_total, cudaStream_t stream) {
  if (gp.Z <= 0 || gp.Z > (1 << 20)) return set_error(TEAL_ERR_BAD_ARG, "Z out of range");
  LaunchCtx& ctx = launch_ctx();
  KParams p;
  memset(&p, 0, sizeof(p));
  p.x = gp.x; p.y = gp.y; p.norm_w = gp.norm_w; p.resid = gp.resid; p.x_out = gp.x_out;
  p.ssq_in = gp.ssq_in; p.n_ssq_in = gp.n_ssq_in; p.ssq_out = gp.ssq_out; p.keep_out = gp.keep_out;
  p.probe = ctx.probe;
  p.Z = gp.Z; p.mode = gp.mode; p.in_mode = gp.in_mode; p.out_mode = gp.out_mode; p.eps = gp.eps;
  p.nseg = gp.nseg;
  const bool deferred = gp.out_mode == kOutDeferred;
  if (deferred) {
    if (!gp.raw_out || (reinterpret_cast<uintptr_t>(gp.raw_out) & 15) || gp.ssq_out || gp.tp || gp.mode == kModeSurvivors)
      return set_error(TEAL_ERR_BAD_ARG, "deferred output: a 16-byte aligned fp32 buffer, no ssq_out, no TP exchange");
    p.acc = gp.raw_out;
  } else if (gp.mode != kModeSurvivors) {
    if (!workspace || ws_bytes < workspace_bytes(out_cols_total)) return set_error(TEAL_ERR_WORKSPACE, "workspace too small");
    if ((reinterpret_cast<uintptr_t>(gp.y) & 7)) return set_error(TEAL_ERR_ALIGNMENT, "y must be 8-byte aligned");
    p.counters = reinterpret_cast<unsigned*>(workspace);
    p.acc = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(workspace) + (size_t)kMaxCounters * 4);
  }
  if (gp.in_mode == kInGateUpAcc) {
    if (!gp.in_acc || gp.in_acc_stride < gp.Z) return set_error(TEAL_ERR_BAD_ARG, "gate/up accumulator input: null / stride < Z");
    p.in_acc = gp.in_acc;
    p.in_acc_stride = gp.in_acc_stride;
  }
  if (gp.clear_n < 0 || (gp.clear_n > 0 && (!gp.clear || (gp.clear_n & 3) || (reinterpret_cast<uintptr_t>(gp.clear) & 15))))
    return set_error(TEAL_ERR_ALIGNMENT, "clear range: 16-byte aligned, a multiple of 4 floats");
  if (gp.clear_n > 0 && gp.in_mode == kInGateUpAcc) {  // CTAs clear their slices while others still load their input
    const float* lo = gp.in_acc;
    const float* hi = gp.in_acc + (size_t)gp.in_acc_stride + gp.Z;
    if (gp.clear < hi && lo < gp.clear + gp.clear_n) return set_error(TEAL_ERR_BAD_ARG, "clear range overlaps the launch's own input");
  }
  if (gp.clear_n > 0 && deferred && gp.clear < gp.raw_out + out_cols_total && gp.raw_out < gp.clear + gp.clear_n)
    return set_error(TEAL_ERR_BAD_ARG, "clear range overlaps the launch's own output");
  if (gp.in_mode == kInRmsNorm && (!gp.norm_w || !gp.ssq_in || gp.n_ssq_in != kSsqParts))
    return set_error(TEAL_ERR_BAD_ARG, "RMSNorm mode needs norm_w and the producer's sum-of-squares partials");
  if (gp.ssq_out && gp.mode == kModePairSilu) return set_error(TEAL_ERR_UNSUPPORTED, "ssq_out is not defined for the gate/up pair");
  if (gp.out_mode == kOutAddResidual && (!gp.resid || (reinterpret_cast<uintptr_t>(gp.resid) & 7)))
    return set_error(TEAL_ERR_ALIGNMENT, "residual must be non-null and 8-byte aligned");

  int variant = (ctx.variant >= 0 && ctx.variant < kNumVariants) ? ctx.variant : 0;
  const Variant& var = kVariants[variant];
  const int cols = 256 * var.vec;
  int ctile0 = 0;
  for (int s = 0; s < gp.nseg; ++s) {
    const GemvSeg& g = gp.seg[s];
    if (g.ncols <= 0 || (g.ncols & 7) || (g.ldw & 7) || (g.out_col0 & 7) || (reinterpret_cast<uintptr_t>(g.w) & 15))
      return set_error(TEAL_ERR_ALIGNMENT, "weights: N, ldw, region offsets must be multiples of 8 and 16-byte aligned");
    if (g.ldw <= 0 || g.ldw >= (1ll << 30)) return set_error(TEAL_ERR_UNSUPPORTED, "ldw out of range");
    KSeg& r = p.seg[s];
    r.w = g.w; r.ldw = g.ldw; r.ncols = g.ncols; r.acc_col0 = g.out_col0;
    r.y_col0 = g.out_col0; r.thr = g.thr;
    r.ctile0 = ctile0; r.nctiles = (g.ncols + cols - 1) / cols;
    ctile0 += r.nctiles;
  }
  if (gp.mode == kModeSurvivors) {  // threshold + keep-mask only: one pseudo column tile
    p.nseg = 1;
    p.seg[0] = KSeg{};
    p.seg[0].thr = gp.seg[0].thr;
    p.seg[0].ncols = 8;
    p.seg[0].nctiles = 1;
    p.seg[0].ldw = 8;
    ctile0 = 1;
  }
  p.total_ctiles = ctile0;
  const bool pair = gp.mode == kModePairSilu;
  if (pair && (gp.nseg != 2 || p.seg[0].nctiles != p.seg[1].nctiles)) return set_error(TEAL_ERR_BAD_ARG, "pair mode needs two equal segments");
  const int ncounters = pair ? p.seg[0].nctiles : ctile0;
  if (ncounters > kMaxCounters) return set_error(TEAL_ERR_UNSUPPORTED, "too many column tiles");
  if (gp.ssq_out && ctile0 > kSsqParts) return set_error(TEAL_ERR_UNSUPPORTED, "ssq_out: more column tiles than TEAL_SSQ_PARTS");

  // K chunk: as many CTAs as fit in ONE wave of resident CTAs, never less than one 32-lane group per warp -- measured on
  // B200 (profiles/r02_quick_ab_wave_fraction.jsonl: grid = 50 / 70 / 90 / 100 / 120 / 150 / 200 % of the slots gives 466 / 497 /
  // 520 / 521 / 464 / 484 / 494 tok/s): a second wave starts with the chip already streaming and ends with a partly idle
  // one; inside one wave, more (shorter) CTAs beat fewer (longer) ones.
  int lpw = ctx.lanes_per_warp;
  if (lpw <= 0) {
    const DeviceInfo& di = device_info();
    const long long slots = (long long)(di.sm_count > 0 ? di.sm_count : 148) * occupancy_of(variant);
    const long long nk_max = slots / ctile0 > 0 ? slots / ctile0 : 1;  // K chunks per column tile in one wave
    long long rows = (gp.Z + nk_max - 1) / nk_max;
    if (rows < 32 * var.warps) rows = 32 * var.warps;
    lpw = (int)((rows + var.warps - 1) / var.warps);
    if (lpw > kListCap) lpw = kListCap;  // (huge Z: several waves)
  }
  if (lpw < 1 || lpw > kListCap) return set_error(TEAL_ERR_BAD_ARG, "lanes per warp must be in [1, 512]");
  p.lanes_per_warp = lpw;
  p.rows_per_cta = lpw * var.warps;
  p.nkchunks = (gp.Z + p.rows_per_cta - 1) / p.rows_per_cta;
  p.target = pair ? 2 * p.nkchunks : p.nkchunks;
  const long long grid = (long long)p.total_ctiles * p.nkchunks;
  if (grid > 0x7fffffffll) return set_error(TEAL_ERR_UNSUPPORTED, "grid too large");
  if (gp.clear_n > 0) {
    p.clear = gp.clear;
    p.clear_n = gp.clear_n;
    p.clear_per_cta = (int)((((long long)gp.clear_n + grid - 1) / grid + 3) & ~3ll);
  }

  void (*kern)(const KParams) = var.kern;
  if (gp.tp) {  // row-parallel projection with the TP exchange fused into the tile finaliser
    if (gp.mode != kModePlain || gp.out_mode != kOutAddResidual || gp.nseg != 1 || (gp.seg[0].ncols & 7))
      return set_error(TEAL_ERR_BAD_ARG, "fused TP exchange: one region, residual output");
    if (gp.tp->world < 1 || gp.tp->world > kMaxTpWorld || gp.tp->rank < 0 || gp.tp->rank >= gp.tp->world || !gp.tp->epoch ||
        ctile0 > kSsqParts || gp.tp->n != gp.seg[0].ncols || (gp.tp->n & 1))
      return set_error(TEAL_ERR_BAD_ARG, "fused TP exchange: bad rank / world / tile count / row length");
    for (int r = 0; r < gp.tp->world; ++r)
      if (!gp.tp->recv[r] || (reinterpret_cast<uintptr_t>(gp.tp->recv[r]) & 7))
        return set_error(TEAL_ERR_ALIGNMENT, "fused TP exchange: null / unaligned peer pointer");
    p.tp = *gp.tp;
    kern = var.kern_tp;
  } else if (ctx.timeline && ctx.timeline_next < ctx.timeline_slots && grid <= kTimelineMaxCtas) {
    // profiling (teal_set_debug_timeline): one slot per launch, in launch order
    p.dbg_time = ctx.timeline + (size_t)(ctx.timeline_next++) * kTimelineSlotInts;
    kern = var.kern_timed;
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
  cfg.numAttrs = ctx.pdl ? 1 : 0;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, p);
  if (e != cudaSuccess) return set_cuda_error(e, "sparse_gemv launch");
  return TEAL_OK;
}

int launch_ssq(const __half* x, int Z, float* ssq_out, int nparts, cudaStream_t stream) {
  if (!x || !ssq_out || Z <= 0 || nparts <= 0 || nparts > kSsqParts) return set_error(TEAL_ERR_BAD_ARG, "ssq: bad argument");
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(nparts);
  cfg.blockDim = dim3(256);
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = launch_ctx().pdl ? 1 : 0;
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

int launch_simple(const __half* x, const __half* wt, __half* y, int Z, int N, long long ldw, float thr, cudaStream_t stream) {
  if (N & 7) return set_error(TEAL_ERR_ALIGNMENT, "N must be a multiple of 8");
  const int warps = (N + 255) / 256;
  const int threads = 128;
  const int blocks = (warps * 32 + threads - 1) / threads;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(blocks);
  cfg.blockDim = dim3(threads);
  cfg.stream = stream;
  cudaError_t e = cudaLaunchKernelEx(&cfg, sparse_gemv_simple_kernel, x, wt, y, Z, N, ldw, thr);
  if (e != cudaSuccess) return set_cuda_error(e, "simple launch");
  return TEAL_OK;
}

size_t workspace_bytes(int max_out_cols) {
  size_t acc = ((size_t)max_out_cols * 4 + 255) & ~(size_t)255;
  return (size_t)kMaxCounters * 4 + acc + (size_t)kSsqParts * kSsqReplicas * 4;  // [counters][acc][ssq scratch]
}
float* workspace_ssq_scratch(void* workspace, size_t ws_bytes) {
  return reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(workspace) + ws_bytes - (size_t)kSsqParts * kSsqReplicas * 4);
}

}  // namespace teal
