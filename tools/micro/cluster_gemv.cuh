// cluster_gemv: fused [RMSNorm ->] threshold -> shared-memory survivor compaction -> gather-GEMV
// [-> +residual | SiLU(gate)*up] for sm_100a, with the split-K reduction done INSIDE a thread-block cluster.
//
// Why (measured on B200, profiles/r02_call1_*): with an fp32 accumulator in L2 (RED + fence + arrival counter + last-arriver
// finaliser) every CTA spends ~1 us after its last row on the reduction chain and the last CTA of a tile another ~0.8 us,
// all CTAs in lock-step -- 1.8 us of a 3-18 us projection with HBM idle.  Here the C CTAs of one cluster own one column
// tile and split the Z input lanes between them: partials meet in distributed shared memory (one cluster barrier,
// ~0.2 us), CTA r of the cluster sums column slice r in rank order and writes the final fp16 values.  No global
// accumulator, no counters, no atomics, no fences, no workspace to keep zeroed -- and a fixed summation order, so the
// result is bit-reproducible from run to run (the RED order was not).
//
//   per CTA : [wait for the producer (PDL)] -> x lanes of my K range -> (RMSNorm scale) -> |x| > thr -> ballots ->
//             prefix over the 32-lane groups -> compacted list {row, fp16 x} in shared memory (ascending rows)
//   per warp: list entries warp, warp + W, ...: a rolling ring of U rows in flight per lane (LDG.128 evict-first,
//             512 B coalesced per instruction and row), fp32 accumulation
//   per CTA : warp partials -> shared memory -> CTA partial [COLS]
//   cluster : barrier -> rank r reduces columns [r*COLS/C, (r+1)*COLS/C) over the C partials (DSMEM loads), epilogue,
//             fp16 store, sum of squares of its slice for the next RMSNorm -> barrier -> exit
//
// Replaces kernels/sparse_gemv.py:45-83 (splitk_sparse_gemv_kernel + its Y.zero_() pre-hook) and :147-194 (qkv_kernel)
// of the reference; thresholds are selected per column region exactly.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../teal_b200/csrc/common.cuh"

namespace teal {

constexpr int kCgMaxSeg = 4;       // column regions per launch (q | k | v, or gate | up)
constexpr int kCgMaxGroups = 128;  // 32-lane groups of one CTA's K range (<= 4096 lanes)

struct CgSeg {
  const __half* w;  // fp16 [Z][ldw], first column of the region
  long long ldw;    // elements
  int ncols;        // columns of the region
  int y_col0;       // first output column of the region
  int tile0;        // first column tile of the region
  float thr;        // fp16-representable threshold
};

struct CgParams {
  const __half* x;        // [Z] input lanes
  const __half* norm_w;   // RMSNorm weight (in_mode == kInRmsNorm)
  const float* ssq_in;    // RMSNorm: partial sums of squares of x left by its producer
  __half* y;              // output (plain / residual: [sum of region widths]; pair: [ncols])
  const __half* resid;    // residual stream (out_mode == kOutAddResidual)
  float* ssq_out;         // optional: partial sums of squares of y, slot [tile * C + rank]; slots up to ssq_slots zeroed
  __half* x_out;          // optional: the (normalised) x actually consumed
  unsigned char* keep_out;       // optional: keep mask (parity hook)
  unsigned long long* probe;     // optional [2]: gathered / dense weight bytes
  long long* dbg_time;           // optional [grid][8]: globaltimer stamps of the CTA's phases (tools/micro/cluster_gemv_bench)
  int Z, nseg, mode, in_mode, out_mode;
  int n_ssq_in, ssq_slots;
  int lanes_per_cta;      // K range of one CTA (multiple of 32); C * lanes_per_cta >= Z
  int ntiles;             // column tiles (pair: tiles of ONE matrix)
  float eps;
  CgSeg seg[kCgMaxSeg];
};

namespace cg {

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_nctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_arrive_release() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait_acquire() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t map_to_rank(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ float4 ld_dsmem_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared::cluster.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}
// streaming 16-byte load: read-only path, no L1 allocation, L2 evict-first (a weight row is read once per token)
__device__ __forceinline__ uint4 ld_stream(const void* p, uint64_t policy) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p), "l"(policy));
  return r;
}
__device__ __forceinline__ void ffma2(float2& acc, const float2 w, const float2 x) {
  asm("fma.rn.f32x2 %0, %1, %2, %0;"
      : "+l"(*reinterpret_cast<unsigned long long*>(&acc))
      : "l"(*reinterpret_cast<const unsigned long long*>(&w)), "l"(*reinterpret_cast<const unsigned long long*>(&x)));
}
// f32 += f16 * f16: the product of two fp16 values is exact in fp32, so this equals fma(float(w), float(x), acc)
__device__ __forceinline__ void fhfma2(float2& acc, uint32_t w2, unsigned short xh) {
  asm("{\n"
      ".reg .b16 lo, hi;\n"
      "mov.b32 {lo, hi}, %2;\n"
      "fma.rn.f32.f16 %0, lo, %3, %0;\n"
      "fma.rn.f32.f16 %1, hi, %3, %1;\n"
      "}\n"
      : "+f"(acc.x), "+f"(acc.y)
      : "r"(w2), "h"(xh));
}

}  // namespace cg

// VEC: 16-byte vectors per lane per row (column tile = 256 * VEC); U: rows in flight per lane; WARPS: warps per CTA;
// PAIR: the CTA gathers its tile of the gate matrix, then of the up matrix (own survivor list each: the two thresholds
// may differ) and the epilogue writes fp16(silu(fp16 g)) * fp16 u; FH: FHFMA arithmetic instead of conversions + FFMA2;
// DB: two batches of U rows per lane, the second requested before the first is consumed.
template <int VEC, int U, int WARPS, bool PAIR, bool FH, bool DB = false>
__global__ void __launch_bounds__(WARPS * 32) cluster_gemv_kernel(const __grid_constant__ CgParams p) {
  constexpr int COLS = 256 * VEC, NT = WARPS * 32, NSUB = PAIR ? 2 : 1;
  extern __shared__ __align__(16) unsigned char s_dyn[];  // survivor list {u16 local row, fp16 x}; later the warp partials
  __shared__ __align__(16) float s_part[NSUB][COLS];      // this CTA's partial, read by the cluster peers through DSMEM
  __shared__ unsigned s_gmask[kCgMaxGroups];
  __shared__ unsigned short s_goff[kCgMaxGroups + 1];
  __shared__ float s_ssq[WARPS];
  uint32_t* const s_list = reinterpret_cast<uint32_t*>(s_dyn);                                  // [lanes_per_cta]
  unsigned short* const s_xn = reinterpret_cast<unsigned short*>(s_dyn + 4 * (size_t)p.lanes_per_cta);  // [lanes_per_cta] fp16 bits
  float* const s_red = reinterpret_cast<float*>(s_dyn);                                         // [WARPS][COLS], after the gathers

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t C = cg::cluster_nctarank(), rank = cg::cluster_ctarank();
  const int tile = (int)(blockIdx.x / C);
  const int kbeg = (int)rank * p.lanes_per_cta, kend = min(p.Z, kbeg + p.lanes_per_cta);
  const int ngroups = max(0, (kend - kbeg + 31) >> 5);
  const uint64_t policy = make_evict_first_policy();
  long long* const dbg = p.dbg_time ? p.dbg_time + (size_t)blockIdx.x * 8 : nullptr;
  auto stamp = [&](int i) {
    if (dbg && tid == 0) dbg[i] = globaltimer_ns();
  };
  stamp(0);

  // which region is my tile in (plain: one of nseg; pair: seg[0] = gate, seg[1] = up, same tile of both)
  int sidx = 0;
  if (!PAIR) {
#pragma unroll
    for (int j = 1; j < kCgMaxSeg; ++j) sidx += (j < p.nseg && tile >= p.seg[j].tile0) ? 1 : 0;
  }
  const int col0 = (tile - (PAIR ? 0 : p.seg[sidx].tile0)) * COLS;
  const int wcols = min(COLS, p.seg[sidx].ncols - col0);
  bool vact[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) vact[v] = lane * 8 + 256 * v < wcols;

  // the norm weights are constants: pull my lanes' lines towards L2 before the producer is done
  if (p.in_mode == kInRmsNorm && tid * 64 < kend - kbeg) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.norm_w + kbeg + tid * 64));
  griddep_wait();               // PDL: everything above overlapped the previous kernel's tail
  griddep_launch_dependents();  // all our CTAs are resident (one wave): the next kernel may move in behind us
  stamp(1);

  // ---- RMSNorm scale from the producer's partial sums of squares (gpt-fast/model.py:286-291) -------------------------
  float rs = 1.f;
  if (p.in_mode == kInRmsNorm) {
    float t = 0.f;
    for (int i = lane; i < p.n_ssq_in; i += 32) t += __ldcg(p.ssq_in + i);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) t += __shfl_xor_sync(0xffffffffu, t, d);
    rs = rsqrtf(t / (float)p.Z + p.eps);
  }
  // ---- my lanes of x (normalised: fp32 norm, cast to fp16, times the fp16 weight in fp16) -> shared memory ------------
#pragma unroll 4
  for (int g = warp; g < ngroups; g += WARPS) {
    const int k = kbeg + g * 32 + lane;
    __half xv = __ushort_as_half((unsigned short)0);
    if (k < kend) {
      xv = __ushort_as_half(__ldcg(reinterpret_cast<const unsigned short*>(p.x) + k));
      if (p.in_mode == kInRmsNorm) {
        xv = __hmul(__float2half_rn(__half2float(xv) * rs), p.norm_w[k]);
        if (p.x_out && tile == 0) p.x_out[k] = xv;
      }
    }
    s_xn[g * 32 + lane] = __half_as_ushort(xv);
  }

  float2 acc[NSUB][VEC][4];
#pragma unroll
  for (int s = 0; s < NSUB; ++s)
#pragma unroll
    for (int v = 0; v < VEC; ++v)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[s][v][c] = make_float2(0.f, 0.f);

  unsigned long long my_bytes = 0;
#pragma unroll
  for (int sub = 0; sub < NSUB; ++sub) {
    const CgSeg& sg = p.seg[PAIR ? sub : sidx];
    const float thr = sg.thr;
    // ---- survivor compaction: ballots -> prefix over the groups -> list in ascending row order ----------------------
    if (sub > 0) __syncthreads();  // the previous list has been consumed by every warp
    for (int g = warp; g < ngroups; g += WARPS) {  // (each warp re-reads only what it wrote itself)
      const int k = kbeg + g * 32 + lane;
      const float xf = __half2float(__ushort_as_half(s_xn[g * 32 + lane]));
      const bool keep = k < kend && !(fabsf(xf) <= thr);  // strict >, NaN kept (it must poison y like mask*x)
      const unsigned m = __ballot_sync(0xffffffffu, keep);
      if (lane == 0) s_gmask[g] = m;
      if (p.keep_out && tile == 0 && sub == 0 && k < kend) p.keep_out[k] = keep ? 1 : 0;
    }
    __syncthreads();
    if (warp == 0) {  // exclusive prefix of the group counts (<= 128 groups: 4 per lane)
      int c[kCgMaxGroups / 32], tot = 0;
#pragma unroll
      for (int j = 0; j < kCgMaxGroups / 32; ++j) {
        const int g = lane * (kCgMaxGroups / 32) + j;
        c[j] = g < ngroups ? __popc(s_gmask[g]) : 0;
        tot += c[j];
      }
      int incl = tot;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int o = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += o;
      }
      int run = incl - tot;
#pragma unroll
      for (int j = 0; j < kCgMaxGroups / 32; ++j) {
        s_goff[lane * (kCgMaxGroups / 32) + j] = (unsigned short)run;
        run += c[j];
      }
      if (lane == 31) s_goff[kCgMaxGroups] = (unsigned short)run;
    }
    __syncthreads();
    const int n = s_goff[kCgMaxGroups];
    for (int g = warp; g < ngroups; g += WARPS) {
      const unsigned m = s_gmask[g];
      if ((m >> lane) & 1u)
        s_list[s_goff[g] + __popc(m & ((1u << lane) - 1u))] = (uint32_t)(g * 32 + lane) | ((uint32_t)s_xn[g * 32 + lane] << 16);
    }
    __syncthreads();
    if (sub == 0) stamp(2);
    if (tid == 0) my_bytes += (unsigned long long)n * (unsigned long long)wcols * 2ull;
    if (p.mode == kModeSurvivors) continue;

    // ---- gather: batches of U rows in flight per lane (all requests of a batch are issued before the first is consumed;
    // a register ring refilled row by row serialises on the load scoreboards -- measured, profiles/r02_cluster_gemv_1) ------
    const __half* wb = sg.w + (long long)kbeg * sg.ldw + col0 + lane * 8;
    const long long ldw = sg.ldw;
    auto issue = [&](uint4 (&buf)[U][VEC], uint32_t (&ent)[U], int e0) {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int e = e0 + u * WARPS;
        if (e < n) {
          ent[u] = s_list[e];
          const __half* src = wb + (long long)(ent[u] & 0xffffu) * ldw;
#pragma unroll
          for (int v = 0; v < VEC; ++v)
            if (vact[v]) buf[u][v] = cg::ld_stream(src + 256 * v, policy);
        }
      }
    };
    auto consume = [&](const uint4 (&buf)[U][VEC], const uint32_t (&ent)[U], int e0) {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (e0 + u * WARPS < n) {
          const unsigned short xb = (unsigned short)(ent[u] >> 16);
          const float xs = __half2float(__ushort_as_half(xb));
          const float2 x2 = make_float2(xs, xs);
#pragma unroll
          for (int v = 0; v < VEC; ++v) {
            if (vact[v]) {
              if constexpr (FH) {
                cg::fhfma2(acc[sub][v][0], buf[u][v].x, xb);
                cg::fhfma2(acc[sub][v][1], buf[u][v].y, xb);
                cg::fhfma2(acc[sub][v][2], buf[u][v].z, xb);
                cg::fhfma2(acc[sub][v][3], buf[u][v].w, xb);
              } else {
                cg::ffma2(acc[sub][v][0], __half22float2(*reinterpret_cast<const __half2*>(&buf[u][v].x)), x2);
                cg::ffma2(acc[sub][v][1], __half22float2(*reinterpret_cast<const __half2*>(&buf[u][v].y)), x2);
                cg::ffma2(acc[sub][v][2], __half22float2(*reinterpret_cast<const __half2*>(&buf[u][v].z)), x2);
                cg::ffma2(acc[sub][v][3], __half22float2(*reinterpret_cast<const __half2*>(&buf[u][v].w)), x2);
              }
            }
          }
        }
      }
    };
    if constexpr (DB) {  // two batches: B is requested before A is consumed
      uint4 bufA[U][VEC], bufB[U][VEC];
      uint32_t entA[U], entB[U];
      issue(bufA, entA, warp);
      for (int e0 = warp; e0 < n; e0 += 2 * U * WARPS) {
        issue(bufB, entB, e0 + U * WARPS);
        consume(bufA, entA, e0);
        issue(bufA, entA, e0 + 2 * U * WARPS);
        consume(bufB, entB, e0 + U * WARPS);
      }
    } else {
      uint4 buf[U][VEC];
      uint32_t ent[U];
      for (int e0 = warp; e0 < n; e0 += U * WARPS) {
        issue(buf, ent, e0);
        consume(buf, ent, e0);
      }
    }
  }
  if (p.probe && tid == 0) {  // measurement probe: gathered vs dense weight bytes
    atomicAdd(p.probe, my_bytes);
    atomicAdd(p.probe + 1, (unsigned long long)max(0, kend - kbeg) * (unsigned long long)wcols * 2ull * NSUB);
  }
  if (p.mode == kModeSurvivors) return;  // (every CTA of the cluster returns here: no barrier is left half-arrived)

  stamp(3);
  // ---- CTA partial: warp partials -> shared memory (the list is dead) -> s_part ---------------------------------------
#pragma unroll
  for (int sub = 0; sub < NSUB; ++sub) {
    __syncthreads();  // list / previous partials consumed
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      float* d = s_red + warp * COLS + lane * 8 + 256 * v;
      *reinterpret_cast<float4*>(d) = make_float4(acc[sub][v][0].x, acc[sub][v][0].y, acc[sub][v][1].x, acc[sub][v][1].y);
      *reinterpret_cast<float4*>(d + 4) = make_float4(acc[sub][v][2].x, acc[sub][v][2].y, acc[sub][v][3].x, acc[sub][v][3].y);
    }
    __syncthreads();
    for (int c = tid * 4; c < COLS; c += NT * 4) {
      float4 a = *reinterpret_cast<const float4*>(s_red + c);
#pragma unroll
      for (int w = 1; w < WARPS; ++w) {
        const float4 b = *reinterpret_cast<const float4*>(s_red + w * COLS + c);
        a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
      }
      *reinterpret_cast<float4*>(&s_part[sub][c]) = a;
    }
  }
  // ---- cluster reduction: rank r owns columns [r * S, (r + 1) * S) of the tile ---------------------------------------
  stamp(4);
  cg::cluster_arrive_release();
  cg::cluster_wait_acquire();
  stamp(5);
  const int S = COLS / (int)C;  // C divides 256
  float ssq = 0.f;
  if (tid * 4 < S) {
    const int c = (int)rank * S + tid * 4;
    float4 a[NSUB];
#pragma unroll
    for (int sub = 0; sub < NSUB; ++sub) {
      const uint32_t local = smem_u32(&s_part[sub][c]);
      a[sub] = make_float4(0.f, 0.f, 0.f, 0.f);
      for (uint32_t r = 0; r < C; ++r) {  // rank order: deterministic
        const float4 b = cg::ld_dsmem_f4(cg::map_to_rank(local, r));
        a[sub].x += b.x; a[sub].y += b.y; a[sub].z += b.z; a[sub].w += b.w;
      }
    }
    if (c < wcols) {
      __half2 h0, h1;
      const CgSeg& sg = p.seg[PAIR ? 0 : sidx];
      const int y0 = sg.y_col0 + col0 + c;
      if constexpr (PAIR) {
        h0 = __halves2half2(silu_mul_f16(a[0].x, a[NSUB - 1].x), silu_mul_f16(a[0].y, a[NSUB - 1].y));
        h1 = __halves2half2(silu_mul_f16(a[0].z, a[NSUB - 1].z), silu_mul_f16(a[0].w, a[NSUB - 1].w));
      } else {
        h0 = __floats2half2_rn(a[0].x, a[0].y);
        h1 = __floats2half2_rn(a[0].z, a[0].w);
        if (p.out_mode == kOutAddResidual) {  // h = x + y in fp16 (gpt-fast/model.py:160-161)
          const uint2 rv = *reinterpret_cast<const uint2*>(p.resid + y0);
          h0 = __hadd2(*reinterpret_cast<const __half2*>(&rv.x), h0);
          h1 = __hadd2(*reinterpret_cast<const __half2*>(&rv.y), h1);
        }
      }
      uint2 o;
      o.x = *reinterpret_cast<const uint32_t*>(&h0);
      o.y = *reinterpret_cast<const uint32_t*>(&h1);
      *reinterpret_cast<uint2*>(p.y + y0) = o;
      const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
      ssq = fmaf(f0.x, f0.x, fmaf(f0.y, f0.y, fmaf(f1.x, f1.x, fmaf(f1.y, f1.y, 0.f))));
    }
  }
  if (p.ssq_out) {  // my slice's contribution to the next RMSNorm (fixed order: deterministic)
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) ssq += __shfl_xor_sync(0xffffffffu, ssq, d);
    if (lane == 0) s_ssq[warp] = ssq;
    __syncthreads();
    if (tid == 0) {
      float a = 0.f;
#pragma unroll
      for (int w = 0; w < WARPS; ++w) a += s_ssq[w];
      const int slot = tile * (int)C + (int)rank;
      p.ssq_out[slot] = a;
      // every slot of the ssq_slots-long buffer is (re)defined on every call: the tail beyond the last tile by tile 0's CTAs
      if (tile == 0)
        for (int sidx2 = p.ntiles * (int)C + (int)rank; sidx2 < p.ssq_slots; sidx2 += (int)C) p.ssq_out[sidx2] = 0.f;
    }
  }
  stamp(6);
  cg::cluster_arrive_release();  // nobody leaves while a peer may still read its partial
  cg::cluster_wait_acquire();
  stamp(7);
}

}  // namespace teal
