// step_program: one persistent data-flow kernel per decoded token (sm_100a).
//
// The launch-per-projection pipeline (rowchunk_gemv.cu + decode_kernels.cu, 163 launches per token for an
// 8B model) leaves HBM idle at every kernel boundary: the tail of kernel i (RED -> fence -> counter ->
// finaliser), the grid drain, and the head of kernel i+1 (x load, ballot, first DRAM access) are ~5 us on
// projections that stream for 3-18 us.  Three of the five boundaries of a layer are
// not real barriers:
//   gate/up -> down : a K chunk of the down projection needs only the h tile(s) covering its lanes
//   qkv -> attention: a head needs its q tile and the k / v tiles of its kv head
//   attention -> o  : a K chunk of the o projection needs only the heads covering its lanes
// Only the two RMSNorm points (sum of squares over the whole residual stream) are.
//
// So the step is planned once into a table of phases (step_program.cuh) and run by ONE kernel:
//   * grid = resident CTA slots (148 SMs x occupancy), 4 warps per CTA, same inner loop as rowchunk_gemv
//     (ballot compaction, LDG.128 evict-first streaming, FFMA2, smem + fp32 RED reduction, last-arriver
//     finaliser with fused residual / SiLU*mul / sum-of-squares epilogue);
//   * a CTA takes items from a global ticket counter (next ticket prefetched while the current item
//     streams).  Tickets are issued in dependency order, so every flag an item waits on belongs to an item
//     with a smaller ticket, i.e. to a CTA that is already running: the spin-waits cannot deadlock whatever
//     the number of resident CTAs (and a watchdog traps after ~2 s instead of hanging the GPU);
//   * a finished tile / head publishes `flags[i] = epoch` (release) and bumps the phase's `done` cell;
//     consumers acquire exactly the flags they read, RMSNorm consumers wait for the whole producer phase;
//   * flags hold the token serial number (epoch) and `done` cells are cumulative, so nothing is reset between
//     tokens; the last CTA to leave rewinds the ticket counter and bumps the epoch.
// Data written by other CTAs of the same launch (x, qkv, att, h, sums of squares, accumulators) is read
// with ld.global.cg (L2) after the acquire; weights stay on the non-coherent streaming path.
//
// Reference call stack this replaces: decode_one_token (gpt-fast/generate.py:73-77) -> Transformer.forward
// (model.py:133-143) -> TransformerBlock (model.py:159-161) -> _new_attn_forward (model.py:163-190) /
// _new_ffn_forward (model.py:258-259) -> torch.ops.teal.sparse_gemv / sparse_qkv_gemv
// (kernels/sparse_gemv.py:45-83,147-194).
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "common.cuh"
#include "step_program.cuh"

namespace teal {

namespace {

// ------------------------------------------------------------------------------------------------ device helpers
#ifdef TEAL_EMU  // CPU emulation build of tests/emu: C++ twins of the inline PTX below (acquire loads yield the fiber)
__device__ __forceinline__ void sp_ffma2(float2& acc, const float2 w, const float2 x) {
  acc.x = fmaf(w.x, x.x, acc.x);
  acc.y = fmaf(w.y, x.y, acc.y);
}
__device__ __forceinline__ uint4 sp_ld_stream_u4(const void* p, uint64_t) { uint4 r; memcpy(&r, p, sizeof r); return r; }
__device__ __forceinline__ unsigned sp_ld_acquire(const unsigned* p) { ::emu::yield(); return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
__device__ __forceinline__ void sp_st_release(unsigned* p, unsigned v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }
__device__ __forceinline__ void sp_red_release_inc(unsigned* p) { __atomic_fetch_add(p, 1u, __ATOMIC_RELEASE); }
#else
__device__ __forceinline__ void sp_ffma2(float2& acc, const float2 w, const float2 x) {
  asm("fma.rn.f32x2 %0, %1, %2, %0;"
      : "+l"(*reinterpret_cast<unsigned long long*>(&acc))
      : "l"(*reinterpret_cast<const unsigned long long*>(&w)), "l"(*reinterpret_cast<const unsigned long long*>(&x)));
}
__device__ __forceinline__ uint4 sp_ld_stream_u4(const void* p, uint64_t policy) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p), "l"(policy));
  return r;
}
__device__ __forceinline__ unsigned sp_ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void sp_st_release(unsigned* p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void sp_red_release_inc(unsigned* p) {
  asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p) : "memory");
}
#endif
// fp16 written by another CTA of this launch: read at L2
__device__ __forceinline__ __half sp_ldcg_half(const __half* p) {
  return __ushort_as_half(__ldcg(reinterpret_cast<const unsigned short*>(p)));
}
__device__ __forceinline__ float2 sp_ldcg_half2(const __half* p) {
  const unsigned v = __ldcg(reinterpret_cast<const unsigned*>(p));
  return __half22float2(*reinterpret_cast<const __half2*>(&v));
}

// Spin until *p has reached `want` (wrap-safe).  Forward progress is guaranteed by the ticket order; the
// watchdog turns a protocol bug into a trap (sticky CUDA error) instead of a hung GPU.
__device__ __forceinline__ void sp_spin_reach(const unsigned* p, unsigned want, unsigned poll_ns) {
  if ((int)(sp_ld_acquire(p) - want) >= 0) return;
  const long long t0 = clock64();
  while ((int)(sp_ld_acquire(p) - want) < 0) {
    if (poll_ns) __nanosleep(poll_ns);
    if (clock64() - t0 > 4000000000ll) __trap();
  }
}

// same, reporting whether the caller had to wait (i.e. it was released by the producer, not late to the party)
__device__ __forceinline__ bool sp_spin_reach_ext(const unsigned* p, unsigned want, unsigned poll_ns) {
  if ((int)(sp_ld_acquire(p) - want) >= 0) return false;
  const long long t0 = clock64();
  while ((int)(sp_ld_acquire(p) - want) < 0) {
    if (poll_ns) __nanosleep(poll_ns);
    if (clock64() - t0 > 4000000000ll) __trap();
  }
  return true;
}

// All threads of the CTA call this after the stores of publication `index`: barrier, then one thread makes the
// CTA's stores visible device-wide (cumulative release) and raises the flag / bumps the phase counter.
__device__ __forceinline__ void sp_publish(const SpPub& pub, int index, unsigned epoch, unsigned* flags, unsigned* done) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (pub.flag0 >= 0) sp_st_release(flags + pub.flag0 + index, epoch);
    sp_red_release_inc(done + pub.cell);
  }
}

#ifdef TEAL_EMU
__device__ __forceinline__ unsigned long long sp_ld_acquire_u64(const unsigned long long* p) { ::emu::yield(); return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
__device__ __forceinline__ void sp_st_release_u64(unsigned long long* p, unsigned long long v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }
#else
__device__ __forceinline__ unsigned long long sp_ld_acquire_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void sp_st_release_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
#endif

// Experimental publication (opts bit 2): like sp_publish, but the publication that completes the phase (it sees the
// cumulative counter reach epoch * count) sums the phase's sum-of-squares partials in the consumers' order, computes
// the scale of the next RMSNorm and writes {epoch, scale} to kSpBcast replicas, one L2 sector each.  Consumers poll
// one replica and get "phase complete" and the scale in ONE round trip; nobody hammers a single cell / sector.
__device__ __forceinline__ void sp_publish_bcast(const SpPub& pub, int index, unsigned epoch, unsigned* flags, unsigned* done,
                                                 unsigned long long* bcast) {
  __syncthreads();
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    unsigned last = 0;
    if (lane == 0) {
      __threadfence();
      if (pub.flag0 >= 0) sp_st_release(flags + pub.flag0 + index, epoch);
      const unsigned old = atomicAdd(done + pub.cell, 1u);  // fence + relaxed RMW = release
      last = (old + 1u == epoch * (unsigned)pub.count) ? 1u : 0u;
      if (last) __threadfence();  // relaxed RMW + fence = acquire: the other publications' partials are visible
    }
    last = __shfl_sync(0xffffffffu, last, 0);
    __syncwarp();  // memory ordering inside the warp: lane 0's acquire happens-before the other lanes' loads / stores
    if (last) {
      float v = 0.f;
      for (int i = lane; i < pub.n_ssq; i += 32) v += __ldcg(pub.ssq + i);
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
      const float rs = rsqrtf(v / (float)pub.rs_Z + pub.rs_eps);
      const unsigned long long word = ((unsigned long long)__float_as_uint(rs) << 32) | (unsigned long long)epoch;
      sp_st_release_u64(bcast + ((size_t)pub.bcast0 * kSpBcast + lane) * kSpBcastStride, word);
    }
  }
}

template <bool EXT>
__device__ __forceinline__ void sp_publish_any(const SpPub& pub, int index, unsigned epoch, unsigned* flags, unsigned* done,
                                               unsigned long long* bcast) {
  if constexpr (EXT) {
    if (pub.bcast0 >= 0 && bcast) {
      sp_publish_bcast(pub, index, epoch, flags, done, bcast);
      return;
    }
  }
  sp_publish(pub, index, epoch, flags, done);
}

// ------------------------------------------------------------------------------------------------ GEMV item
// One (column tile, K chunk) of a projection: [RMSNorm ->] threshold -> ballot compaction -> gather-GEMV ->
// smem + RED reduction -> arrival counter -> (last arriver) epilogue + publication.  Same arithmetic as
// rowchunk_gemv_kernel (kernels/sparse_gemv.py:45-83,147-194 of the reference).
template <int VEC, int U, int W, bool EXT>
__device__ __forceinline__ void sp_gemv_item(const SpGemv& g, const SpPub& pub, int item, unsigned epoch, unsigned* flags,
                                             unsigned* done, float* s_red, float* s_ssq, int* s_last, bool rs_known,
                                             float rs_value, unsigned long long* bcast) {
  constexpr int COLS = 256 * VEC;
  constexpr int kThreadsW = W * 32;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int ct, kc;
  sp_gemv_decode(g, item, &ct, &kc);
  int s = 0;
#pragma unroll
  for (int j = 1; j < 3; ++j) s += (j < g.nseg && ct >= g.seg[j].ctile0) ? 1 : 0;
  const SpSeg& sg = g.seg[s];
  const int col0 = (ct - sg.ctile0) * COLS;
  const int wcols = min(COLS, sg.ncols - col0);
  const float thr = sg.thr;
  bool vact[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) vact[v] = lane * 8 + 256 * v < wcols;
  const __half* wbase = sg.w + col0 + lane * 8;
  const long long ldw = sg.ldw;
  const uint64_t policy = make_evict_first_policy();

  // RMSNorm scale from the producer's per-tile sums of squares (gpt-fast/model.py:286-291)
  float rs = 1.f;
  const bool rms = g.in_mode == kInRmsNorm;
  bool load_ssq = rms;
  if constexpr (EXT) {
    if (rs_known) {  // computed once per (CTA, phase) by the caller: no 4 x CTAs hot-spot on the partials' L2 sector
      rs = rs_value;
      load_ssq = false;
    }
  }
  if (load_ssq) {
    float t = 0.f;
    for (int i = lane; i < g.n_ssq_in; i += 32) t += __ldcg(g.ssq_in + i);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) t += __shfl_xor_sync(0xffffffffu, t, d);
    rs = rsqrtf(t / (float)g.Z + g.eps);
  }

  float2 acc[VEC][4];
#pragma unroll
  for (int v = 0; v < VEC; ++v)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[v][c] = make_float2(0.f, 0.f);

  const int kbeg = kc * g.rows_per_item;
  const int kend = min(g.Z, kbeg + g.rows_per_item);
  const int lpw = g.lpw;
  for (int kb = kbeg + warp * lpw; kb < kend; kb += W * lpw) {
    const int k = kb + lane;
    const bool valid = lane < lpw && k < kend;
    float xf = 0.f;
    if (valid) {
      const __half xh = sp_ldcg_half(g.x + k);
      if (rms) {  // fp32 norm, cast to fp16, multiply by the fp16 weight in fp16
        xf = __half2float(__hmul(__float2half_rn(__half2float(xh) * rs), g.norm_w[k]));
      } else {
        xf = __half2float(xh);
      }
    }
    const bool keep = valid && !(fabsf(xf) <= thr);  // strict >, NaN kept (it must poison y like mask*x)
    unsigned mask = __ballot_sync(0xffffffffu, keep);
    while (mask) {
      int bit[U];
      uint4 wv[U][VEC];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        bit[u] = mask ? (__ffs(mask) - 1) : -1;
        mask &= mask - 1;  // no-op when mask == 0
        const __half* src = wbase + (long long)(kb + (bit[u] >= 0 ? bit[u] : 0)) * ldw;
#pragma unroll
        for (int v = 0; v < VEC; ++v)  // every element is defined on every pass: nothing is carried around the loop
          wv[u][v] = (bit[u] >= 0 && vact[v]) ? sp_ld_stream_u4(src + 256 * v, policy) : make_uint4(0u, 0u, 0u, 0u);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (bit[u] >= 0) {
          const float xs = __shfl_sync(0xffffffffu, xf, bit[u]);
          const float2 x2 = make_float2(xs, xs);
#pragma unroll
          for (int v = 0; v < VEC; ++v) {
            if (vact[v]) {
              sp_ffma2(acc[v][0], __half22float2(*reinterpret_cast<const __half2*>(&wv[u][v].x)), x2);
              sp_ffma2(acc[v][1], __half22float2(*reinterpret_cast<const __half2*>(&wv[u][v].y)), x2);
              sp_ffma2(acc[v][2], __half22float2(*reinterpret_cast<const __half2*>(&wv[u][v].z)), x2);
              sp_ffma2(acc[v][3], __half22float2(*reinterpret_cast<const __half2*>(&wv[u][v].w)), x2);
            }
          }
        }
      }
    }
  }

  // ---- CTA partial: 4 warps -> smem -> fp32 vector RED into the L2 accumulator ------------------------
#pragma unroll
  for (int v = 0; v < VEC; ++v) {
    float* d = s_red + warp * COLS + lane * 8 + 256 * v;
    *reinterpret_cast<float4*>(d) = make_float4(acc[v][0].x, acc[v][0].y, acc[v][1].x, acc[v][1].y);
    *reinterpret_cast<float4*>(d + 4) = make_float4(acc[v][2].x, acc[v][2].y, acc[v][3].x, acc[v][3].y);
  }
  __syncthreads();
  const int acc0 = sg.out_col0 + col0;
  for (int c = tid * 4; c < wcols; c += kThreadsW * 4) {
    float4 a = *reinterpret_cast<const float4*>(s_red + c);
#pragma unroll
    for (int w = 1; w < W; ++w) {
      const float4 b = *reinterpret_cast<const float4*>(s_red + w * COLS + c);
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    if (a.x != 0.f || a.y != 0.f || a.z != 0.f || a.w != 0.f) red_add_v4(g.acc + acc0 + c, a);
  }
  // ---- arrival counter: the REDs of all threads happen-before the barrier, thread 0's fence is cumulative over
  // them; the last arrival acquires and finalises the tile ---------------------------------------------------
  __syncthreads();
  const bool pair = g.mode == kModePairSilu;
  const int cidx = sp_gemv_cidx(g, ct);
  if (tid == 0) {
    fence_acq_rel_gpu();
    const unsigned old = atomicAdd(g.counters + cidx, 1u);
    const int last = ((int)old + 1 == g.target) ? 1 : 0;
    if (last) fence_acq_rel_gpu();
    *s_last = last;
  }
  __syncthreads();
  if (!*s_last) return;

  float ssq = 0.f;
  if (pair) {
    const int g0 = g.seg[0].out_col0 + col0, u0 = g.seg[1].out_col0 + col0;
    for (int c = tid * 4; c < wcols; c += kThreadsW * 4) {
      const float4 gg = ld_cg_f4(g.acc + g0 + c), uu = ld_cg_f4(g.acc + u0 + c);
      st_cg_f4(g.acc + g0 + c, make_float4(0.f, 0.f, 0.f, 0.f));
      st_cg_f4(g.acc + u0 + c, make_float4(0.f, 0.f, 0.f, 0.f));
      const __half2 h0 = __halves2half2(silu_mul_f16(gg.x, uu.x), silu_mul_f16(gg.y, uu.y));
      const __half2 h1 = __halves2half2(silu_mul_f16(gg.z, uu.z), silu_mul_f16(gg.w, uu.w));
      uint2 o;
      o.x = *reinterpret_cast<const uint32_t*>(&h0);
      o.y = *reinterpret_cast<const uint32_t*>(&h1);
      *reinterpret_cast<uint2*>(g.y + col0 + c) = o;
    }
  } else {
    for (int c = tid * 4; c < wcols; c += kThreadsW * 4) {
      const float4 a = ld_cg_f4(g.acc + acc0 + c);
      st_cg_f4(g.acc + acc0 + c, make_float4(0.f, 0.f, 0.f, 0.f));
      __half2 h0 = __floats2half2_rn(a.x, a.y), h1 = __floats2half2_rn(a.z, a.w);
      if (g.out_mode == kOutAddResidual) {  // h = x + y in fp16 (gpt-fast/model.py:160-161)
        const uint2 r = __ldcg(reinterpret_cast<const uint2*>(g.resid + acc0 + c));
        h0 = __hadd2(*reinterpret_cast<const __half2*>(&r.x), h0);
        h1 = __hadd2(*reinterpret_cast<const __half2*>(&r.y), h1);
      }
      uint2 o;
      o.x = *reinterpret_cast<const uint32_t*>(&h0);
      o.y = *reinterpret_cast<const uint32_t*>(&h1);
      *reinterpret_cast<uint2*>(g.y + acc0 + c) = o;
      const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
      ssq = fmaf(f0.x, f0.x, fmaf(f0.y, f0.y, fmaf(f1.x, f1.x, fmaf(f1.y, f1.y, ssq))));
    }
  }
  if (g.ssq_out) {  // this tile's contribution to the next RMSNorm (fixed reduction order: deterministic)
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) ssq += __shfl_xor_sync(0xffffffffu, ssq, d);
    if (lane == 0) s_ssq[warp] = ssq;
    __syncthreads();
    if (tid == 0) {
      float a = 0.f;
#pragma unroll
      for (int w = 0; w < W; ++w) a += s_ssq[w];
      g.ssq_out[cidx] = a;
    }
  }
  if (tid == 0) g.counters[cidx] = 0;
  sp_publish_any<EXT>(pub, cidx, epoch, flags, done, bcast);
}

// ------------------------------------------------------------------------------------------------ multi-tile GEMV item
// Experimental (teal_step_desc::tiles_per_item > 1): one item sweeps up to kSpMaxTpi column tiles of ONE K chunk.
// The x lanes are loaded once (one L2 round trip, one dependency check, one ticket for tpi tiles); the first rows of
// tile t+1 are issued BEFORE tile t is reduced, so the reduction chain (smem -> RED -> fence -> counter -> epilogue)
// overlaps DRAM latency instead of leaving the CTA's share of HBM idle.  Same arithmetic per tile as sp_gemv_item.
template <int VEC, int W, bool EXT>
__device__ __forceinline__ void sp_gemv_tail(const SpGemv& g, const SpPub& pub, int ct, const SpSeg& sg, int col0, int wcols,
                                             float2 (&acc)[VEC][4], unsigned epoch, unsigned* flags, unsigned* done,
                                             float* s_red, float* s_ssq, int* s_last, unsigned long long* bcast) {
  constexpr int COLS = 256 * VEC;
  constexpr int kThreadsW = W * 32;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
#pragma unroll
  for (int v = 0; v < VEC; ++v) {
    float* d = s_red + warp * COLS + lane * 8 + 256 * v;
    *reinterpret_cast<float4*>(d) = make_float4(acc[v][0].x, acc[v][0].y, acc[v][1].x, acc[v][1].y);
    *reinterpret_cast<float4*>(d + 4) = make_float4(acc[v][2].x, acc[v][2].y, acc[v][3].x, acc[v][3].y);
  }
  __syncthreads();
  const int acc0 = sg.out_col0 + col0;
  for (int c = tid * 4; c < wcols; c += kThreadsW * 4) {
    float4 a = *reinterpret_cast<const float4*>(s_red + c);
#pragma unroll
    for (int w = 1; w < W; ++w) {
      const float4 b = *reinterpret_cast<const float4*>(s_red + w * COLS + c);
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    if (a.x != 0.f || a.y != 0.f || a.z != 0.f || a.w != 0.f) red_add_v4(g.acc + acc0 + c, a);
  }
  __syncthreads();
  const bool pair = g.mode == kModePairSilu;
  const int cidx = sp_gemv_cidx(g, ct);
  if (tid == 0) {
    fence_acq_rel_gpu();
    const unsigned old = atomicAdd(g.counters + cidx, 1u);
    const int last = ((int)old + 1 == g.target) ? 1 : 0;
    if (last) fence_acq_rel_gpu();
    *s_last = last;
  }
  __syncthreads();
  if (*s_last) {  // CTA-uniform
    float ssq = 0.f;
    if (pair) {
      const int g0 = g.seg[0].out_col0 + col0, u0 = g.seg[1].out_col0 + col0;
      for (int c = tid * 4; c < wcols; c += kThreadsW * 4) {
        const float4 gg = ld_cg_f4(g.acc + g0 + c), uu = ld_cg_f4(g.acc + u0 + c);
        st_cg_f4(g.acc + g0 + c, make_float4(0.f, 0.f, 0.f, 0.f));
        st_cg_f4(g.acc + u0 + c, make_float4(0.f, 0.f, 0.f, 0.f));
        const __half2 h0 = __halves2half2(silu_mul_f16(gg.x, uu.x), silu_mul_f16(gg.y, uu.y));
        const __half2 h1 = __halves2half2(silu_mul_f16(gg.z, uu.z), silu_mul_f16(gg.w, uu.w));
        uint2 o;
        o.x = *reinterpret_cast<const uint32_t*>(&h0);
        o.y = *reinterpret_cast<const uint32_t*>(&h1);
        *reinterpret_cast<uint2*>(g.y + col0 + c) = o;
      }
    } else {
      for (int c = tid * 4; c < wcols; c += kThreadsW * 4) {
        const float4 a = ld_cg_f4(g.acc + acc0 + c);
        st_cg_f4(g.acc + acc0 + c, make_float4(0.f, 0.f, 0.f, 0.f));
        __half2 h0 = __floats2half2_rn(a.x, a.y), h1 = __floats2half2_rn(a.z, a.w);
        if (g.out_mode == kOutAddResidual) {
          const uint2 r = __ldcg(reinterpret_cast<const uint2*>(g.resid + acc0 + c));
          h0 = __hadd2(*reinterpret_cast<const __half2*>(&r.x), h0);
          h1 = __hadd2(*reinterpret_cast<const __half2*>(&r.y), h1);
        }
        uint2 o;
        o.x = *reinterpret_cast<const uint32_t*>(&h0);
        o.y = *reinterpret_cast<const uint32_t*>(&h1);
        *reinterpret_cast<uint2*>(g.y + acc0 + c) = o;
        const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
        ssq = fmaf(f0.x, f0.x, fmaf(f0.y, f0.y, fmaf(f1.x, f1.x, fmaf(f1.y, f1.y, ssq))));
      }
    }
    if (g.ssq_out) {
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) ssq += __shfl_xor_sync(0xffffffffu, ssq, d);
      if (lane == 0) s_ssq[warp] = ssq;
      __syncthreads();
      if (tid == 0) {
        float a = 0.f;
#pragma unroll
        for (int w = 0; w < W; ++w) a += s_ssq[w];
        g.ssq_out[cidx] = a;
      }
    }
    if (tid == 0) g.counters[cidx] = 0;
    sp_publish_any<EXT>(pub, cidx, epoch, flags, done, bcast);
  }
}

#ifdef TEAL_EMU
__device__ __forceinline__ void sp_prefetch_l2(const void* p) { (void)*reinterpret_cast<const volatile char*>(p); }
#else
__device__ __forceinline__ void sp_prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
#endif

template <int VEC, int U, int W>
__device__ __forceinline__ void sp_gemv_item_multi(const SpGemv& g, const SpPub& pub, int item, unsigned epoch, unsigned* flags,
                                                   unsigned* done, float* s_red, float* s_ssq, int* s_last, bool rs_known,
                                                   float rs_value, unsigned long long* bcast) {
  constexpr int COLS = 256 * VEC;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int p0, ntiles, kc;
  sp_gemv_decode_multi(g, item, &p0, &ntiles, &kc);
  const uint64_t policy = make_evict_first_policy();

  float rs = 1.f;
  const bool rms = g.in_mode == kInRmsNorm;
  if (rms) {
    if (rs_known) {
      rs = rs_value;
    } else {
      float t = 0.f;
      for (int i = lane; i < g.n_ssq_in; i += 32) t += __ldcg(g.ssq_in + i);
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) t += __shfl_xor_sync(0xffffffffu, t, d);
      rs = rsqrtf(t / (float)g.Z + g.eps);
    }
  }
  // ---- the CTA's K chunk: ONE pass (the planner guarantees rows_per_item == W * lpw for multi-tile items) ----------
  const int kbeg = kc * g.rows_per_item;
  const int kend = min(g.Z, kbeg + g.rows_per_item);
  const int kb = kbeg + warp * g.lpw;
  const int k = kb + lane;
  const bool valid = lane < g.lpw && k < kend;
  float xf = 0.f;
  if (valid) {
    const __half xh = sp_ldcg_half(g.x + k);
    xf = rms ? __half2float(__hmul(__float2half_rn(__half2float(xh) * rs), g.norm_w[k])) : __half2float(xh);
  }

  // ---- descriptor of the tile at order position p0 -------------------------------------------------------------
  int ct = sp_gemv_pos_tile(g, p0);
  int sidx = 0;
#pragma unroll
  for (int j = 1; j < 3; ++j) sidx += (j < g.nseg && ct >= g.seg[j].ctile0) ? 1 : 0;
  int col0 = (ct - g.seg[sidx].ctile0) * COLS;
  int wcols = min(COLS, g.seg[sidx].ncols - col0);
  unsigned mask = __ballot_sync(0xffffffffu, valid && !(fabsf(xf) <= g.seg[sidx].thr));

#pragma unroll 1
  for (int t = 0; t < ntiles; ++t) {
    const SpSeg& sg = g.seg[sidx];
    const __half* wbase = sg.w + col0 + lane * 8;
    const long long ldw = sg.ldw;
    bool vact[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) vact[v] = lane * 8 + 256 * v < wcols;
    float2 acc[VEC][4];
#pragma unroll
    for (int v = 0; v < VEC; ++v)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[v][c] = make_float2(0.f, 0.f);
    while (mask) {
      int bit[U];
      uint4 wv[U][VEC];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        bit[u] = mask ? (__ffs(mask) - 1) : -1;
        mask &= mask - 1;
        const __half* src = wbase + (long long)(kb + (bit[u] >= 0 ? bit[u] : 0)) * ldw;
#pragma unroll
        for (int v = 0; v < VEC; ++v)
          wv[u][v] = (bit[u] >= 0 && vact[v]) ? sp_ld_stream_u4(src + 256 * v, policy) : make_uint4(0u, 0u, 0u, 0u);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (bit[u] >= 0) {
          const float xs = __shfl_sync(0xffffffffu, xf, bit[u]);
          const float2 x2 = make_float2(xs, xs);
#pragma unroll
          for (int v = 0; v < VEC; ++v) {
            if (vact[v]) {
              sp_ffma2(acc[v][0], __half22float2(*reinterpret_cast<const __half2*>(&wv[u][v].x)), x2);
              sp_ffma2(acc[v][1], __half22float2(*reinterpret_cast<const __half2*>(&wv[u][v].y)), x2);
              sp_ffma2(acc[v][2], __half22float2(*reinterpret_cast<const __half2*>(&wv[u][v].z)), x2);
              sp_ffma2(acc[v][3], __half22float2(*reinterpret_cast<const __half2*>(&wv[u][v].w)), x2);
            }
          }
        }
      }
    }
    // ---- before reducing tile t: start the DRAM reads of tile t+1's first rows (into L2: no registers are tied up
    // across the reduction, the reduction chain overlaps DRAM latency and HBM keeps working for this CTA) -------------
    const int ct_cur = ct, col0_cur = col0, wcols_cur = wcols;
    if (t + 1 < ntiles) {
      ct = sp_gemv_pos_tile(g, p0 + t + 1);
      sidx = 0;
#pragma unroll
      for (int j = 1; j < 3; ++j) sidx += (j < g.nseg && ct >= g.seg[j].ctile0) ? 1 : 0;
      const SpSeg& sn = g.seg[sidx];
      col0 = (ct - sn.ctile0) * COLS;
      wcols = min(COLS, sn.ncols - col0);
      mask = __ballot_sync(0xffffffffu, valid && !(fabsf(xf) <= sn.thr));
      unsigned m = mask;
      const __half* nbase = sn.w + col0 + lane * 8;
#pragma unroll 1
      for (int u = 0; u < 2 * U && m; ++u) {
        const int b = __ffs(m) - 1;
        m &= m - 1;
        const __half* src = nbase + (long long)(kb + b) * sn.ldw;
#pragma unroll
        for (int v = 0; v < VEC; ++v)
          if (lane * 8 + 256 * v < wcols) sp_prefetch_l2(src + 256 * v);
      }
    }
    sp_gemv_tail<VEC, W, true>(g, pub, ct_cur, sg, col0_cur, wcols_cur, acc, epoch, flags, done, s_red, s_ssq, s_last, bcast);
  }
}

// ------------------------------------------------------------------------------------------------ attention item
// RoPE(q, k) + KV-cache append + attention of the one query over positions [0, pos] for ONE query head
// (gpt-fast/model.py:170-186).  Same arithmetic as attn_decode_kernel, 4 warps.
template <int HD, int W>
__device__ __forceinline__ void sp_attn_item(const SpAttn& a, const SpPub& pub, int h, unsigned epoch, unsigned* flags,
                                             unsigned* done, float* smem) {
  constexpr int EPL = HD / 32;
  float* s_q = smem;
  float* s_knew = s_q + HD;
  float* s_vnew = s_knew + HD;
  float* s_m = s_vnew + HD;
  float* s_l = s_m + W;
  float* s_acc = s_l + W;  // [W][HD]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int rep = a.n_head / a.n_kv;
  const int kvh = h / rep;
  const int pos = __ldcg(a.pos);
  const int max_seq = a.max_seq;

  const __half* qh = a.qkv + (size_t)h * HD;
  const __half* kh = a.qkv + (size_t)a.n_head * HD + (size_t)kvh * HD;
  const __half* vh = a.qkv + (size_t)(a.n_head + a.n_kv) * HD + (size_t)kvh * HD;
  const __half* fr = a.freqs + (size_t)pos * HD;  // HD/2 (cos, sin) pairs
  for (int i = tid; i < HD; i += (W * 32)) {
    if (i < HD / 2) {
      const float c = __half2float(fr[2 * i]), s = __half2float(fr[2 * i + 1]);
      const float2 q = sp_ldcg_half2(qh + 2 * i), k = sp_ldcg_half2(kh + 2 * i);
      // model.py:309-315, result cast back to fp16 (type_as)
      const __half qa = __float2half_rn(q.x * c - q.y * s), qb = __float2half_rn(q.y * c + q.x * s);
      const __half ka = __float2half_rn(k.x * c - k.y * s), kb = __float2half_rn(k.y * c + k.x * s);
      s_q[2 * i] = __half2float(qa);
      s_q[2 * i + 1] = __half2float(qb);
      s_knew[2 * i] = __half2float(ka);
      s_knew[2 * i + 1] = __half2float(kb);
      if (h % rep == 0) {  // one head per kv group appends to the cache (model.py:93-94)
        a.k_cache[((size_t)kvh * max_seq + pos) * HD + 2 * i] = ka;
        a.k_cache[((size_t)kvh * max_seq + pos) * HD + 2 * i + 1] = kb;
      }
    } else {
      const int j = i - HD / 2;
      const unsigned raw = __ldcg(reinterpret_cast<const unsigned*>(vh + 2 * j));
      const __half2 v2 = *reinterpret_cast<const __half2*>(&raw);
      s_vnew[2 * j] = __low2float(v2);
      s_vnew[2 * j + 1] = __high2float(v2);
      if (h % rep == 0) *reinterpret_cast<unsigned*>(a.v_cache + ((size_t)kvh * max_seq + pos) * HD + 2 * j) = raw;
    }
  }
  __syncthreads();

  float qreg[EPL];
#pragma unroll
  for (int e = 0; e < EPL; ++e) qreg[e] = s_q[lane * EPL + e];
  float m = -CUDART_INF_F, l = 0.f, acc[EPL];
#pragma unroll
  for (int e = 0; e < EPL; ++e) acc[e] = 0.f;
  const __half* kbase = a.k_cache + (size_t)kvh * max_seq * HD;
  const __half* vbase = a.v_cache + (size_t)kvh * max_seq * HD;
  constexpr int PB = 4;  // positions per iteration: independent loads (latency-bound at short context)
  for (int p0 = warp; p0 <= pos; p0 += W * PB) {
    float kk[PB][EPL], vv[PB][EPL];
#pragma unroll
    for (int b = 0; b < PB; ++b) {
      const int p = p0 + b * W;
      if (p < pos) {  // positions written by earlier launches
#pragma unroll
        for (int e = 0; e < EPL / 2; ++e) {
          const float2 kf = sp_ldcg_half2(kbase + (size_t)p * HD + lane * EPL + 2 * e);
          const float2 vf = sp_ldcg_half2(vbase + (size_t)p * HD + lane * EPL + 2 * e);
          kk[b][2 * e] = kf.x; kk[b][2 * e + 1] = kf.y;
          vv[b][2 * e] = vf.x; vv[b][2 * e + 1] = vf.y;
        }
      } else {  // the position appended in this step comes from shared memory (every head, not only the appender)
#pragma unroll
        for (int e = 0; e < EPL; ++e) { kk[b][e] = s_knew[lane * EPL + e]; vv[b][e] = s_vnew[lane * EPL + e]; }
      }
    }
#pragma unroll
    for (int b = 0; b < PB; ++b) {
      const int p = p0 + b * W;
      if (p <= pos) {
        float s = 0.f;
#pragma unroll
        for (int e = 0; e < EPL; ++e) s = fmaf(qreg[e], kk[b][e], s);
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
        s *= a.scale;
        const float mn = fmaxf(m, s);
        const float corr = __expf(m - mn), pe = __expf(s - mn);
        l = l * corr + pe;
#pragma unroll
        for (int e = 0; e < EPL; ++e) acc[e] = fmaf(pe, vv[b][e], acc[e] * corr);
        m = mn;
      }
    }
  }
  if (lane == 0) { s_m[warp] = m; s_l[warp] = l; }
#pragma unroll
  for (int e = 0; e < EPL; ++e) s_acc[warp * HD + lane * EPL + e] = acc[e];
  __syncthreads();
  for (int i = tid; i < HD; i += (W * 32)) {
    float M = -CUDART_INF_F;
#pragma unroll
    for (int w = 0; w < W; ++w) M = fmaxf(M, s_m[w]);
    float L = 0.f, o = 0.f;
#pragma unroll
    for (int w = 0; w < W; ++w) {
      const float f = (s_m[w] == -CUDART_INF_F) ? 0.f : __expf(s_m[w] - M);
      L = fmaf(s_l[w], f, L);
      o = fmaf(s_acc[w * HD + i], f, o);
    }
    a.out[(size_t)h * HD + i] = __float2half_rn(o / L);
  }
  sp_publish(pub, h, epoch, flags, done);
}

// ------------------------------------------------------------------------------------------------ lm_head item
__device__ __forceinline__ unsigned long long sp_pack_argmax(float v, int idx) {
  unsigned b = __float_as_uint(v);
  b = (b & 0x80000000u) ? ~b : (b | 0x80000000u);  // monotonic map float -> unsigned
  return ((unsigned long long)b << 32) | (unsigned long long)(0xffffffffu - (unsigned)idx);  // ties -> smallest index
}

// `rows_per_item` vocabulary rows of logits = W * rmsnorm(x) (gpt-fast/model.py:141-142) + packed arg-max.
// The normalised x (Z halves in shared memory) is rebuilt only when the CTA's previous item was not an lm_head item.
template <int UNROLL, int W>
__device__ __forceinline__ void sp_head_item(const SpHead& hp, const SpPub& pub, int item, bool build_xs, unsigned epoch,
                                             unsigned* flags, unsigned* done, __half* xs, float* s_part,
                                             unsigned long long* s_best) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int Z = hp.Z;
  if (build_xs) {
    float ss = 0.f;
    for (int m = tid * 8; m < Z; m += (W * 32) * 8) {
      const uint4 v = __ldcg(reinterpret_cast<const uint4*>(hp.x + m));
      *reinterpret_cast<uint4*>(xs + m) = v;
      const __half2* h2 = reinterpret_cast<const __half2*>(&v);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __half22float2(h2[j]);
        ss = fmaf(f.x, f.x, ss);
        ss = fmaf(f.y, f.y, ss);
      }
    }
    if (hp.norm_w) {
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, d);
      if (lane == 0) s_part[warp] = ss;
      __syncthreads();
      float tot = (lane < W) ? s_part[lane] : 0.f;
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, d);
      const float rs = rsqrtf(tot / (float)Z + hp.eps);
      for (int m = tid * 8; m < Z; m += (W * 32) * 8) {  // each thread renormalises the vectors it staged itself
        uint4 v = *reinterpret_cast<uint4*>(xs + m);
        const uint4 gw = ld_nc_u4(reinterpret_cast<const uint4*>(hp.norm_w + m));
        __half2* h2 = reinterpret_cast<__half2*>(&v);
        const __half2* g2 = reinterpret_cast<const __half2*>(&gw);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = __half22float2(h2[j]);
          h2[j] = __hmul2(__floats2half2_rn(f.x * rs, f.y * rs), g2[j]);
        }
        *reinterpret_cast<uint4*>(xs + m) = v;
      }
    }
    __syncthreads();
  }
  const int r0 = item * hp.rows_per_item;
  const int r1 = min(hp.V, r0 + hp.rows_per_item);
  const int nvec = Z >> 3;
  unsigned long long best = 0ull;
  for (int v = r0 + warp; v < r1; v += W) {
    const uint4* wr = reinterpret_cast<const uint4*>(hp.W + (size_t)v * hp.ldw);
    float acc = 0.f;
    for (int c0 = lane; c0 < nvec; c0 += 32 * UNROLL) {
      uint4 wv[UNROLL];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        const int c = c0 + 32 * u;
        wv[u] = (c < nvec) ? ld_nc_u4(wr + c) : make_uint4(0u, 0u, 0u, 0u);
      }
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        const int c = c0 + 32 * u;
        if (c < nvec) {
          const uint4 xv = *reinterpret_cast<const uint4*>(xs + c * 8);
          const __half2* w2 = reinterpret_cast<const __half2*>(&wv[u]);
          const __half2* x2 = reinterpret_cast<const __half2*>(&xv);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 wf = __half22float2(w2[j]), xf = __half22float2(x2[j]);
            acc = fmaf(wf.x, xf.x, acc);
            acc = fmaf(wf.y, xf.y, acc);
          }
        }
      }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
    if (lane == 0) {
      const __half r = __float2half_rn(acc);
      hp.y[v] = r;
      const unsigned long long pk = sp_pack_argmax(__half2float(r), v);
      best = pk > best ? pk : best;
    }
  }
  if (hp.argmax) {
    if (lane == 0) s_best[warp] = best;
    __syncthreads();
    if (tid == 0) {
      unsigned long long b = 0ull;
#pragma unroll
      for (int w = 0; w < W; ++w) b = s_best[w] > b ? s_best[w] : b;
      if (b) atomicMax(hp.argmax, b);
    }
  }
  sp_publish(pub, item, epoch, flags, done);
}

// ------------------------------------------------------------------------------------------------ embed / advance
template <int W, bool EXT>
__device__ __forceinline__ void sp_embed_item(const SpEmbed& e, const SpPub& pub, unsigned epoch, unsigned* flags,
                                              unsigned* done, float* s_part, unsigned long long* bcast) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int t = __ldcg(e.token);
  const uint4* src = reinterpret_cast<const uint4*>(e.emb + (size_t)t * e.dim);
  uint4* dst = reinterpret_cast<uint4*>(e.x);
  float ss = 0.f;
  for (int i = tid; i < e.dim / 8; i += (W * 32)) {
    const uint4 v = ld_nc_u4(src + i);
    dst[i] = v;
    const __half2* h2 = reinterpret_cast<const __half2*>(&v);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = __half22float2(h2[j]);
      ss = fmaf(f.x, f.x, fmaf(f.y, f.y, ss));
    }
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, d);
  if (lane == 0) s_part[warp] = ss;
  __syncthreads();
  if (tid == 0) {
    float a = 0.f;
#pragma unroll
    for (int w = 0; w < W; ++w) a += s_part[w];
    e.ssq_out[0] = a;
  }
  sp_publish_any<EXT>(pub, 0, epoch, flags, done, bcast);
}

// greedy step bookkeeping (generate.py:49-66 with temperature -> 0): token = arg-max, append, ++pos, reset cell
__device__ __forceinline__ void sp_advance_item(const SpAdvance& v) {
  if (threadIdx.x == 0) {
    const unsigned long long pk = __ldcg(v.argmax);
    const int t = (int)(0xffffffffu - (unsigned)(pk & 0xffffffffull));
    const int p = __ldcg(v.pos);
    v.token[0] = t;
    if (p < v.out_cap) v.out_tokens[p] = t;
    v.pos[0] = p + 1;
    *v.argmax = 0ull;
  }
}

#ifdef TEAL_EMU
#define sp_smem (::emu::dyn_smem())
#else
extern __shared__ __align__(16) unsigned char sp_smem[];
#endif

// ------------------------------------------------------------------------------------------------ the kernel
// W: warps per CTA (4: 7 CTAs per SM; 2: 14 CTAs per SM, the launch-per-projection engine's shape for the small
// projections).  EXT = false is the configuration validated on hardware in round 1 (profiles/r01_step_program_*);
// EXT = true adds the experimental paths selected at run time by SpHeader::opts / stagger_ns:
//   opts bit 1  the RMSNorm scale is computed once per (CTA, phase) by warp 0 right after the whole-phase wait instead
//               of by every warp of every item (the partials are one 64-byte L2 sector read by every warp of every CTA
//               at the same instant after a barrier)
//   stagger_ns  a CTA released by a whole-phase wait sleeps hash(blockIdx) in [0, 8) x stagger_ns before it starts, so
//               that the ~1000 released CTAs do not run their load / reduce phases in lock-step
template <int W, bool EXT>
__global__ void __launch_bounds__(W * 32, W == 4 ? 7 : 14) step_program_kernel(const unsigned char* __restrict__ blob) {
  constexpr int kThreadsW = W * 32;
  __shared__ __align__(16) SpPhase s_ph;
  __shared__ int s_ctl[2];
  __shared__ int s_deps[kSpMaxDeps];
  __shared__ int s_last;
  __shared__ float s_small[W];
  __shared__ unsigned long long s_best[W];
  __shared__ float s_rs;
  __shared__ int s_rs_valid;

  const SpHeader* hdr = reinterpret_cast<const SpHeader*>(blob);
  const SpPhase* phases = reinterpret_cast<const SpPhase*>(blob + sizeof(SpHeader));
  const int tid = threadIdx.x, lane = tid & 31;
  const int total = hdr->total_items;
  unsigned* const ticket = hdr->ticket;
  unsigned* const flags = hdr->flags;
  unsigned* const done = hdr->done;
  // serial number of this token: only the last CTA to LEAVE a launch bumps the cell, so every CTA of a launch reads the
  // same value no matter when it starts
  const unsigned epoch = __ldcg(hdr->epoch) + 1u;
  const unsigned poll_ns = (unsigned)hdr->poll_ns;
  const bool cache_all = (hdr->opts & 1) != 0;
  int dep_ok = -1;  // phase whose whole-phase dependency this CTA has already seen satisfied
  unsigned long long* const dbg = hdr->dbg;
  long long t_ready = 0;
  // EXT state
  const bool rs_once = EXT && (hdr->opts & 2) != 0;
  const unsigned stagger = EXT ? ((blockIdx.x * 0x9E3779B1u) >> 29) * (unsigned)hdr->stagger_ns : 0u;
  int rs_phase = -1;
  float rs_value = 1.f;
  unsigned long long* const bcast = EXT ? hdr->bcast : nullptr;

  unsigned next_t = 0;
  if (tid == 0) next_t = atomicAdd(ticket, 1u);
  if constexpr (EXT) {
    if (tid == 0) s_rs_valid = 0;
  }
  int ph_idx = 0;   // thread 0 only: phase cursor (tickets of one CTA only grow)
  int loaded = -1;  // phase currently in s_ph
  int xs_phase = -1;  // lm_head phase whose normalised x sits in shared memory

  for (;;) {
    if (tid == 0) {
      const unsigned t = next_t;
      if (t < (unsigned)total) {
        next_t = atomicAdd(ticket, 1u);  // prefetch: the round trip overlaps this item
        while ((int)t >= __ldg(&phases[ph_idx].item0) + __ldg(&phases[ph_idx].nitems)) ++ph_idx;
      }
      s_ctl[0] = (t < (unsigned)total) ? (int)t : -1;
      s_ctl[1] = ph_idx;
    }
    __syncthreads();
    const int t = s_ctl[0];
    if (t < 0) break;
    const int pi = s_ctl[1];
    if (pi != loaded) {
      const uint4* src = reinterpret_cast<const uint4*>(phases + pi);
      uint4* dst = reinterpret_cast<uint4*>(&s_ph);
      for (int i = tid; i < (int)(sizeof(SpPhase) / 16); i += kThreadsW) dst[i] = ld_nc_u4(src + i);
      loaded = pi;
      __syncthreads();
    }
    const int item = t - s_ph.item0;
    if (dbg && tid == 0) atomicMin(dbg + 4 * pi, (unsigned long long)globaltimer_ns());

    // ---- wait for exactly what this item reads --------------------------------------------------------------
    if (s_ph.dep.kind != kDepNone && !(cache_all && s_ph.dep.kind == kDepAll && dep_ok == pi)) {
      if (tid < 32) {
        if (s_ph.dep.kind == kDepAll) {
          if constexpr (EXT) {
            if (lane == 0) {
              bool waited;
              if (bcast && s_ph.dep.bcast0 >= 0) {  // one replica of the producer's {epoch, RMSNorm scale} word
                const unsigned long long* cell =
                    bcast + ((size_t)s_ph.dep.bcast0 * kSpBcast + (blockIdx.x & (kSpBcast - 1))) * kSpBcastStride;
                unsigned long long v = sp_ld_acquire_u64(cell);
                waited = (int)((unsigned)v - epoch) < 0;
                if (waited) {
                  const long long t0 = clock64();
                  while ((int)((unsigned)(v = sp_ld_acquire_u64(cell)) - epoch) < 0) {
                    if (poll_ns) __nanosleep(poll_ns);
                    if (clock64() - t0 > 4000000000ll) __trap();
                  }
                }
                s_rs = __uint_as_float((unsigned)(v >> 32));
                s_rs_valid = 1;  // s_rs holds this phase's scale
              } else {
                waited = sp_spin_reach_ext(done + s_ph.dep.cell, epoch * (unsigned)s_ph.dep.per_epoch, poll_ns);
              }
              if (waited && stagger) __nanosleep(stagger);
            }
          } else {
            if (lane == 0) sp_spin_reach(done + s_ph.dep.cell, epoch * (unsigned)s_ph.dep.per_epoch, poll_ns);
          }
        } else {
          int n = 0;
          if (lane == 0) n = sp_item_deps<EXT>(s_ph, item, s_deps);
          n = __shfl_sync(0xffffffffu, n, 0);  // also orders lane 0's s_deps stores before the other lanes' loads
          __syncwarp();
          if (lane < n) sp_spin_reach(flags + s_deps[lane], epoch, poll_ns);
        }
      }
      __syncthreads();
      if (s_ph.dep.kind == kDepAll) dep_ok = pi;
      if constexpr (EXT) {
        if (s_rs_valid) {  // scale delivered with the broadcast word (reset below, after everybody has read it)
          rs_value = s_rs;
          rs_phase = pi;
        }
      }
    }
    if (dbg && tid == 0) {
      t_ready = globaltimer_ns();
      atomicMin(dbg + 4 * pi + 1, (unsigned long long)t_ready);
    }
    if constexpr (EXT) {
      // RMSNorm scale once per (CTA, phase): warp 0 sums the producer's partials (fixed order), everybody reads one float
      if (rs_once && s_ph.kind == kPhGemv && s_ph.u.g.in_mode == kInRmsNorm && rs_phase != pi) {
        if (tid < 32) {
          float v = 0.f;
          for (int i = lane; i < s_ph.u.g.n_ssq_in; i += 32) v += __ldcg(s_ph.u.g.ssq_in + i);
#pragma unroll
          for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
          if (lane == 0) s_rs = rsqrtf(v / (float)s_ph.u.g.Z + s_ph.u.g.eps);
        }
        __syncthreads();
        rs_value = s_rs;
        rs_phase = pi;
      }
    }
    const bool rs_known = EXT && rs_phase == pi;

    switch (s_ph.kind) {
      case kPhGemv:
        if (EXT && s_ph.u.g.tpi > 1) {
          if constexpr (EXT) {
            if (s_ph.u.g.variant == 1)
              sp_gemv_item_multi<2, 4, W>(s_ph.u.g, s_ph.pub, item, epoch, flags, done, reinterpret_cast<float*>(sp_smem), s_small,
                                          &s_last, rs_known, rs_value, bcast);
            else
              sp_gemv_item_multi<1, 8, W>(s_ph.u.g, s_ph.pub, item, epoch, flags, done, reinterpret_cast<float*>(sp_smem), s_small,
                                          &s_last, rs_known, rs_value, bcast);
          }
        } else if (s_ph.u.g.variant == 1)
          sp_gemv_item<2, 4, W, EXT>(s_ph.u.g, s_ph.pub, item, epoch, flags, done, reinterpret_cast<float*>(sp_smem), s_small,
                                     &s_last, rs_known, rs_value, bcast);
        else
          sp_gemv_item<1, 8, W, EXT>(s_ph.u.g, s_ph.pub, item, epoch, flags, done, reinterpret_cast<float*>(sp_smem), s_small,
                                     &s_last, rs_known, rs_value, bcast);
        xs_phase = -1;
        break;
      case kPhAttn:
        if (s_ph.u.a.hd == 128)
          sp_attn_item<128, W>(s_ph.u.a, s_ph.pub, item, epoch, flags, done, reinterpret_cast<float*>(sp_smem));
        else
          sp_attn_item<64, W>(s_ph.u.a, s_ph.pub, item, epoch, flags, done, reinterpret_cast<float*>(sp_smem));
        xs_phase = -1;
        break;
      case kPhHead:
        sp_head_item<8, W>(s_ph.u.h, s_ph.pub, item, xs_phase != pi, epoch, flags, done, reinterpret_cast<__half*>(sp_smem),
                           s_small, s_best);
        xs_phase = pi;
        break;
      case kPhEmbed:
        sp_embed_item<W, EXT>(s_ph.u.e, s_ph.pub, epoch, flags, done, s_small, bcast);
        break;
      case kPhAdvance:
        sp_advance_item(s_ph.u.v);
        break;
      default:
        break;
    }
    __syncthreads();  // s_ctl / s_ph / dynamic shared memory are reused by the next item
    if constexpr (EXT) {
      if (tid == 0) s_rs_valid = 0;  // ordered before the next item's reads by the barrier after the ticket hand-out
    }
    if (dbg && tid == 0) {
      const long long t_end = globaltimer_ns();
      atomicMax(dbg + 4 * pi + 2, (unsigned long long)t_end);
      atomicAdd(dbg + 4 * pi + 3, (unsigned long long)(t_end - t_ready));
    }
  }

  // ---- leave: the last CTA out rewinds the ticket counter and bumps the epoch for the next launch -----------
  if (tid == 0) {
    __threadfence();
    const unsigned old = atomicAdd(hdr->exited, 1u);
    if (old + 1u == gridDim.x) {
      *hdr->ticket = 0u;
      *hdr->exited = 0u;
      *hdr->epoch = epoch;
      __threadfence();
    }
  }
}

// ------------------------------------------------------------------------------------------------ planner (host)
struct KindCfg {  // one per projection kind: qkv, o, gate/up, down
  int variant, cols, rows, lpw, warps;
};

int desc_warps(const teal_step_desc* d) { return d->warps > 0 ? d->warps : kSpWarpsDefault; }

struct Plan {
  KindCfg kind[4];
  StepScratchLayout L;
  int nphases;
  int total_tiles[4];   // column tiles per call (pair: both matrices)
  int pub_tiles[4];     // publications per call
  size_t smem;
  int warps;
};

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

int check_desc(const teal_step_desc* d) {
  if (!d) return set_error(TEAL_ERR_BAD_ARG, "step program: desc is null");
  if (d->dim <= 0 || d->n_head <= 0 || d->n_kv_head <= 0 || d->head_dim <= 0 || d->n_layer < 0 || d->max_seq <= 0)
    return set_error(TEAL_ERR_BAD_ARG, "step program: bad model shape");
  if (d->n_head % d->n_kv_head) return set_error(TEAL_ERR_BAD_ARG, "step program: n_head must be a multiple of n_kv_head");
  if (d->head_dim != 64 && d->head_dim != 128) return set_error(TEAL_ERR_UNSUPPORTED, "step program: head_dim must be 64 or 128");
  if ((d->dim & 7) || d->dim > (1 << 20)) return set_error(TEAL_ERR_ALIGNMENT, "step program: dim must be a multiple of 8");
  if (d->n_layer > 0 && !d->layers) return set_error(TEAL_ERR_BAD_ARG, "step program: layers is null");
  if (!d->x || !d->scratch) return set_error(TEAL_ERR_BAD_ARG, "step program: x / scratch is null");
  if (d->warps != 0 && d->warps != 2 && d->warps != 4) return set_error(TEAL_ERR_UNSUPPORTED, "step program: warps must be 2 or 4");
  if (d->stagger_ns < 0 || d->stagger_ns > 100000) return set_error(TEAL_ERR_BAD_ARG, "step program: stagger_ns out of range");
  if (d->tiles_per_item < 0 || d->tiles_per_item > kSpMaxTpi) return set_error(TEAL_ERR_BAD_ARG, "step program: tiles_per_item must be 0..4");
  return TEAL_OK;
}

int make_kind(const teal_step_desc* d, long long elems, int rows_override, KindCfg* k) {
  const int W = desc_warps(d);
  const long long wide = (long long)(d->wide_melems > 0 ? d->wide_melems : 40) * 1000 * 1000;
  k->variant = elems >= wide ? 1 : 0;
  k->cols = k->variant ? 512 : 256;
  int rows = k->variant ? (d->rows_large > 0 ? d->rows_large : 128) : (d->rows_small > 0 ? d->rows_small : 64);
  if (rows_override > 0) rows = rows_override;
  if (rows < W || rows % W) return set_error(TEAL_ERR_BAD_ARG, "step program: rows per item must be a multiple of the warps per CTA");
  k->lpw = rows / W < 32 ? rows / W : 32;
  k->warps = W;
  if (rows % (W * k->lpw)) return set_error(TEAL_ERR_BAD_ARG, "step program: rows per item must be a multiple of 32 x warps once it exceeds it");
  k->rows = rows;
  return TEAL_OK;
}

bool layer_has_attn(const teal_layer_desc& l) { return l.wqkv != nullptr; }
bool layer_has_mlp(const teal_layer_desc& l) { return l.w1 != nullptr; }

int make_plan(const teal_step_desc* d, Plan* P) {
  if (int rc = check_desc(d)) return rc;
  memset(P, 0, sizeof(*P));
  const int nqkv = (d->n_head + 2 * d->n_kv_head) * d->head_dim;
  const int adim = d->n_head * d->head_dim;
  bool any_attn = false, any_mlp = false;
  int nphases = 0, nflags = 0;
  for (int i = 0; i < d->n_layer; ++i) {
    any_attn = any_attn || layer_has_attn(d->layers[i]);
    any_mlp = any_mlp || layer_has_mlp(d->layers[i]);
  }
  if (any_mlp && (d->inter <= 0 || (d->inter & 7))) return set_error(TEAL_ERR_ALIGNMENT, "step program: inter must be a positive multiple of 8");
  const long long elems[4] = {(long long)d->dim * nqkv, (long long)adim * d->dim, 2ll * d->dim * d->inter,
                              (long long)d->inter * d->dim};
  const int out_cols[4] = {nqkv, d->dim, 2 * d->inter, d->dim};
  for (int k = 0; k < 4; ++k) {
    if (int rc = make_kind(d, elems[k], d->rows_kind[k], &P->kind[k])) return rc;
    const int cols = P->kind[k].cols;
    if (k == 2) {
      P->pub_tiles[k] = (d->inter + cols - 1) / cols;
      P->total_tiles[k] = 2 * P->pub_tiles[k];
    } else if (k == 0) {
      // upper bound: q / k / v regions tiled separately (non-uniform thresholds)
      P->total_tiles[k] = (adim + cols - 1) / cols + 2 * ((d->n_kv_head * d->head_dim + cols - 1) / cols);
      P->pub_tiles[k] = P->total_tiles[k];
    } else {
      P->total_tiles[k] = P->pub_tiles[k] = (out_cols[k] + cols - 1) / cols;
    }
  }
  if (d->tok_emb) { ++nphases; nflags += 1; }
  for (int i = 0; i < d->n_layer; ++i) {
    if (layer_has_attn(d->layers[i])) { nphases += 3; nflags += P->pub_tiles[0] + d->n_head + P->pub_tiles[1]; }
    if (layer_has_mlp(d->layers[i])) { nphases += 2; nflags += P->pub_tiles[2] + P->pub_tiles[3]; }
  }
  if (d->lm_head) nphases += 2;
  if (nphases == 0) return set_error(TEAL_ERR_BAD_ARG, "step program: nothing to run");
  P->nphases = nphases;

  StepScratchLayout& L = P->L;
  size_t off = 0;
  L.control = off; off += 64;
  L.ndone = nphases;
  L.done = off; off = align_up(off + (size_t)nphases * 4, 64);
  L.nflags = nflags;
  L.flags = off; off = align_up(off + (size_t)nflags * 4, 64);
  for (int k = 0; k < 4; ++k) {
    L.ncounters[k] = P->pub_tiles[k];
    L.counters[k] = off; off = align_up(off + (size_t)P->pub_tiles[k] * 4, 64);
  }
  for (int k = 0; k < 4; ++k) {
    L.acc[k] = off; off = align_up(off + (size_t)(out_cols[k] > 0 ? out_cols[k] : 8) * 4, 256);
  }
  int max_tiles = 64;
  for (int k = 0; k < 4; ++k) max_tiles = P->total_tiles[k] > max_tiles ? P->total_tiles[k] : max_tiles;
  for (int b = 0; b < 2; ++b) { L.ssq[b] = off; off = align_up(off + (size_t)max_tiles * 4, 64); }
  L.bcast = 0;
  if (d->opts & 4) {  // phase-complete broadcast words: one 128-byte line per replica
    off = align_up(off, 128);
    L.bcast = off;
    off += (size_t)nphases * kSpBcast * kSpBcastStride * 8;
  }
  L.total = off;

  // dynamic shared memory: max over the phase kinds (GEMV partials, attention staging, normalised x of the lm_head)
  const int W = desc_warps(d);
  P->warps = W;
  size_t smem = (size_t)W * 512 * 4;
  const size_t attn = (size_t)(3 * d->head_dim + 2 * W + W * d->head_dim) * 4;
  smem = attn > smem ? attn : smem;
  if (d->lm_head) smem = (size_t)d->dim * 2 > smem ? (size_t)d->dim * 2 : smem;
  P->smem = align_up(smem, 16);
  return TEAL_OK;
}

struct Producer {  // who produced the residual stream the next RMSNorm consumes
  const float* ssq;
  int n_ssq, cell, per_epoch;
  bool valid;
  int bcast0;  // its broadcast block (opts bit 2) or -1
};

int fill_seg(SpSeg* s, const void* w, long long ldw, int ncols, int out_col0, double thr, int cols, int* ctile0) {
  if (!w || ncols <= 0 || (ncols & 7) || (ldw & 7) || ldw < ncols || (out_col0 & 7) || (reinterpret_cast<uintptr_t>(w) & 15))
    return set_error(TEAL_ERR_ALIGNMENT, "step program: weights: N, ldw, region offsets must be multiples of 8 and 16-byte aligned");
  s->w = reinterpret_cast<const __half*>(w);
  s->ldw = ldw;
  s->ncols = ncols;
  s->out_col0 = out_col0;
  s->thr = round_threshold_fp16(thr);
  s->ctile0 = *ctile0;
  s->nctiles = (ncols + cols - 1) / cols;
  *ctile0 += s->nctiles;
  return TEAL_OK;
}

void finish_gemv(SpPhase& ph, const KindCfg& kc, int total_ctiles, int group, int rot, int tpi) {
  SpGemv& g = ph.u.g;
  g.variant = kc.variant;
  g.rows_per_item = kc.rows;
  g.lpw = kc.lpw;
  g.nkchunks = (g.Z + kc.rows - 1) / kc.rows;
  g.total_ctiles = total_ctiles;
  const bool pair = g.mode == kModePairSilu;
  g.target = pair ? 2 * g.nkchunks : g.nkchunks;
  if (group <= 0 || group > total_ctiles) group = total_ctiles;
  if (pair && (group & 1)) group += 1;
  // multi-tile items: one pass over the K chunk only, whole items per group (pair mode: gate j and up j share an item)
  if (tpi < 1 || kc.rows != kc.warps * kc.lpw) tpi = 1;  // K chunks that need several passes keep one tile per item
  if (tpi > kSpMaxTpi) tpi = kSpMaxTpi;
  if (pair && tpi > 1 && (tpi & 1)) tpi += 1;
  if (tpi > 1) {
    group = (group + tpi - 1) / tpi * tpi;
    if (group > total_ctiles) group = total_ctiles;  // a single (possibly ragged) group
  }
  g.tpi = tpi;
  g.group = group;
  g.rot = rot;
  ph.nitems = sp_gemv_items(g);
  ph.pub.count = pair ? total_ctiles / 2 : total_ctiles;
}

int build_program(const teal_step_desc* d, const Plan& P, unsigned char* blob) {
  SpHeader* H = reinterpret_cast<SpHeader*>(blob);
  SpPhase* ph = reinterpret_cast<SpPhase*>(blob + sizeof(SpHeader));
  memset(blob, 0, sizeof(SpHeader) + (size_t)P.nphases * sizeof(SpPhase));
  unsigned char* sc = reinterpret_cast<unsigned char*>(d->scratch);
  const StepScratchLayout& L = P.L;
  unsigned* counters[4];
  float* accs[4];
  float* ssqb[2];
  for (int k = 0; k < 4; ++k) {
    counters[k] = reinterpret_cast<unsigned*>(sc + L.counters[k]);
    accs[k] = reinterpret_cast<float*>(sc + L.acc[k]);
  }
  for (int b = 0; b < 2; ++b) ssqb[b] = reinterpret_cast<float*>(sc + L.ssq[b]);
  const int hd = d->head_dim, adim = d->n_head * hd, kvd = d->n_kv_head * hd, nqkv = adim + 2 * kvd;

  int n = 0, item0 = 0, flag0 = 0, toggle = 0;
  Producer cur = {nullptr, 0, 0, 0, false, -1};
  const bool use_bcast = (d->opts & 4) != 0;
  // a phase whose completion opens an RMSNorm (or the lm_head): broadcast {epoch, scale of the next norm}
  auto producer_of_x = [&](SpPhase& p, float* ssq, int n_ssq) {
    cur = {ssq, n_ssq, n, p.pub.count, true, -1};
    if (use_bcast) {
      p.pub.bcast0 = n;
      p.pub.ssq = ssq;
      p.pub.n_ssq = n_ssq;
      p.pub.rs_Z = d->dim;
      p.pub.rs_eps = (float)d->eps;
      cur.bcast0 = n;
    }
  };
  auto open = [&](int kind, int layer) -> SpPhase& {
    SpPhase& p = ph[n];
    p.kind = kind;
    p.layer = layer;
    p.item0 = item0;
    p.pub.cell = n;
    p.pub.flag0 = -1;
    p.pub.bcast0 = -1;
    p.dep.bcast0 = -1;
    return p;
  };
  auto close = [&](SpPhase& p, bool with_flags) {
    if (with_flags) { p.pub.flag0 = flag0; flag0 += p.pub.count; }
    item0 += p.nitems;
    ++n;
  };
  auto norm_input = [&](SpPhase& p, const void* norm_w) -> int {
    SpGemv& g = p.u.g;
    if (!norm_w) return set_error(TEAL_ERR_BAD_ARG, "step program: norm weight is null");
    g.in_mode = kInRmsNorm;
    g.norm_w = reinterpret_cast<const __half*>(norm_w);
    g.eps = (float)d->eps;
    if (cur.valid) {
      g.ssq_in = cur.ssq;
      g.n_ssq_in = cur.n_ssq;
      p.dep.kind = kDepAll;
      p.dep.cell = cur.cell;
      p.dep.per_epoch = cur.per_epoch;
      p.dep.bcast0 = cur.bcast0;
    } else {  // x preset by the caller
      if (!d->ssq_init) return set_error(TEAL_ERR_BAD_ARG, "step program: ssq_init is required when tok_emb is null");
      g.ssq_in = d->ssq_init;
      g.n_ssq_in = 1;
      p.dep.kind = kDepNone;
    }
    return TEAL_OK;
  };

  if (d->tok_emb) {
    if (!d->token) return set_error(TEAL_ERR_BAD_ARG, "step program: token is null");
    SpPhase& p = open(kPhEmbed, -1);
    p.nitems = 1;
    p.pub.count = 1;
    p.u.e.emb = reinterpret_cast<const __half*>(d->tok_emb);
    p.u.e.token = d->token;
    p.u.e.x = reinterpret_cast<__half*>(d->x);
    p.u.e.ssq_out = ssqb[toggle];
    p.u.e.dim = d->dim;
    producer_of_x(p, ssqb[toggle], 1);
    toggle ^= 1;
    close(p, true);
  }

  for (int li = 0; li < d->n_layer; ++li) {
    const teal_layer_desc& l = d->layers[li];
    if (layer_has_attn(l)) {
      if (!l.wo || !l.k_cache || !l.v_cache || !d->qkv || !d->att || !d->freqs_cis || !d->pos)
        return set_error(TEAL_ERR_BAD_ARG, "step program: attention block needs wo, caches, qkv, att, freqs_cis, pos");
      // ---- RMSNorm + QKV ---------------------------------------------------------------------------------
      const KindCfg& kq = P.kind[0];
      SpPhase& q = open(kPhGemv, li);
      SpGemv& g = q.u.g;
      g.x = reinterpret_cast<const __half*>(d->x);
      g.y = reinterpret_cast<__half*>(d->qkv);
      g.Z = d->dim;
      g.mode = kModePlain;
      g.out_mode = kOutPlain;
      g.acc = accs[0];
      g.counters = counters[0];
      if (int rc = norm_input(q, l.attn_norm)) return rc;
      const float tq = round_threshold_fp16(l.thr[0]), tk = round_threshold_fp16(l.thr[1]), tv = round_threshold_fp16(l.thr[2]);
      int ct0 = 0, rot;
      const __half* w = reinterpret_cast<const __half*>(l.wqkv);
      if (tq == tk && tk == tv) {  // uniform sparsity: one region
        g.nseg = 1;
        if (int rc = fill_seg(&g.seg[0], w, l.ld_qkv, nqkv, 0, l.thr[0], kq.cols, &ct0)) return rc;
        rot = adim / kq.cols;
      } else {  // thresholds selected per column region exactly (the reference selects per BLOCK_N tile, sparse_gemv.py:167-168)
        g.nseg = 3;
        if (l.ld_qkv < nqkv) return set_error(TEAL_ERR_BAD_ARG, "step program: ld_qkv < N");
        if (int rc = fill_seg(&g.seg[0], w, l.ld_qkv, adim, 0, l.thr[0], kq.cols, &ct0)) return rc;
        if (int rc = fill_seg(&g.seg[1], w + adim, l.ld_qkv, kvd, adim, l.thr[1], kq.cols, &ct0)) return rc;
        if (int rc = fill_seg(&g.seg[2], w + adim + kvd, l.ld_qkv, kvd, adim + kvd, l.thr[2], kq.cols, &ct0)) return rc;
        rot = g.seg[1].ctile0;
      }
      if (ct0 > L.ncounters[0]) return set_error(TEAL_ERR_UNSUPPORTED, "step program: QKV tile count exceeds the scratch layout");
      finish_gemv(q, kq, ct0, d->group_qkv > 0 ? d->group_qkv : 4, rot % ct0, d->tiles_per_item);
      const int q_flag0 = flag0;
      const SpGemv qg = g;
      close(q, true);
      // ---- RoPE + KV append + attention, one item per query head -----------------------------------------
      SpPhase& at = open(kPhAttn, li);
      SpAttn& a = at.u.a;
      a.qkv = reinterpret_cast<const __half*>(d->qkv);
      a.freqs = reinterpret_cast<const __half*>(d->freqs_cis);
      a.k_cache = reinterpret_cast<__half*>(l.k_cache);
      a.v_cache = reinterpret_cast<__half*>(l.v_cache);
      a.pos = d->pos;
      a.out = reinterpret_cast<__half*>(d->att);
      a.n_head = d->n_head; a.n_kv = d->n_kv_head; a.hd = hd; a.max_seq = d->max_seq;
      a.scale = 1.0f / sqrtf((float)hd);
      a.prod_cols = kq.cols;
      a.prod_nseg = qg.nseg;
      for (int s = 0; s < qg.nseg; ++s) {
        a.prod_col0[s] = qg.seg[s].out_col0;
        a.prod_ncols[s] = qg.seg[s].ncols;
        a.prod_ctile0[s] = qg.seg[s].ctile0;
      }
      at.nitems = d->n_head;
      at.pub.count = d->n_head;
      at.dep.kind = kDepQkvHead;
      at.dep.flag0 = q_flag0;
      const int a_flag0 = flag0;
      close(at, true);
      // ---- o projection + residual ---------------------------------------------------------------------
      const KindCfg& ko = P.kind[1];
      SpPhase& o = open(kPhGemv, li);
      SpGemv& og = o.u.g;
      og.x = reinterpret_cast<const __half*>(d->att);
      og.y = reinterpret_cast<__half*>(d->x);
      og.resid = reinterpret_cast<const __half*>(d->x);
      og.Z = adim;
      og.mode = kModePlain;
      og.in_mode = kInPlain;
      og.out_mode = kOutAddResidual;
      og.acc = accs[1];
      og.counters = counters[1];
      og.ssq_out = ssqb[toggle];
      og.nseg = 1;
      int ot0 = 0;
      if (int rc = fill_seg(&og.seg[0], l.wo, l.ld_o, d->dim, 0, l.thr[3], ko.cols, &ot0)) return rc;
      finish_gemv(o, ko, ot0, 0, 0, d->tiles_per_item);
      o.dep.kind = kDepLanes;
      o.dep.flag0 = a_flag0;
      o.dep.div = hd;
      if (ko.rows / hd + 2 > kSpMaxDeps) return set_error(TEAL_ERR_UNSUPPORTED, "step program: o-projection K chunk spans too many heads");
      producer_of_x(o, ssqb[toggle], ot0);
      toggle ^= 1;
      close(o, true);
    }
    if (layer_has_mlp(l)) {
      if (!l.w3 || !l.w2 || !d->h) return set_error(TEAL_ERR_BAD_ARG, "step program: MLP block needs w1, w3, w2, h");
      // ---- RMSNorm + gate/up + SiLU*mul ------------------------------------------------------------------
      const KindCfg& kg = P.kind[2];
      SpPhase& gu = open(kPhGemv, li);
      SpGemv& g = gu.u.g;
      g.x = reinterpret_cast<const __half*>(d->x);
      g.y = reinterpret_cast<__half*>(d->h);
      g.Z = d->dim;
      g.mode = kModePairSilu;
      g.out_mode = kOutPlain;
      g.acc = accs[2];
      g.counters = counters[2];
      if (int rc = norm_input(gu, l.ffn_norm)) return rc;
      g.nseg = 2;
      int gt0 = 0;
      if (int rc = fill_seg(&g.seg[0], l.w1, l.ld_1, d->inter, 0, l.thr[4], kg.cols, &gt0)) return rc;
      if (int rc = fill_seg(&g.seg[1], l.w3, l.ld_3, d->inter, d->inter, l.thr[5], kg.cols, &gt0)) return rc;
      finish_gemv(gu, kg, gt0, d->group_gate_up > 0 ? d->group_gate_up : 8, 0, d->tiles_per_item);
      const int g_flag0 = flag0;
      close(gu, true);
      // ---- down projection + residual -------------------------------------------------------------------
      const KindCfg& kd = P.kind[3];
      SpPhase& dn = open(kPhGemv, li);
      SpGemv& dg = dn.u.g;
      dg.x = reinterpret_cast<const __half*>(d->h);
      dg.y = reinterpret_cast<__half*>(d->x);
      dg.resid = reinterpret_cast<const __half*>(d->x);
      dg.Z = d->inter;
      dg.mode = kModePlain;
      dg.in_mode = kInPlain;
      dg.out_mode = kOutAddResidual;
      dg.acc = accs[3];
      dg.counters = counters[3];
      dg.ssq_out = ssqb[toggle];
      dg.nseg = 1;
      int dt0 = 0;
      if (int rc = fill_seg(&dg.seg[0], l.w2, l.ld_2, d->dim, 0, l.thr[6], kd.cols, &dt0)) return rc;
      finish_gemv(dn, kd, dt0, 0, 0, d->tiles_per_item);
      dn.dep.kind = kDepLanes;
      dn.dep.flag0 = g_flag0;
      dn.dep.div = kg.cols;
      if (kd.rows / kg.cols + 2 > kSpMaxDeps) return set_error(TEAL_ERR_UNSUPPORTED, "step program: down-projection K chunk spans too many tiles");
      producer_of_x(dn, ssqb[toggle], dt0);
      toggle ^= 1;
      close(dn, true);
    }
  }

  if (d->lm_head) {
    if (!d->logits || d->vocab <= 0 || d->ld_head < d->dim || (d->ld_head & 7) || (reinterpret_cast<uintptr_t>(d->lm_head) & 15) ||
        (reinterpret_cast<uintptr_t>(d->x) & 15))
      return set_error(TEAL_ERR_BAD_ARG, "step program: lm_head needs logits, vocab, ld_head >= dim (multiple of 8), 16-byte aligned pointers");
    if (!d->argmax_packed || !d->token || !d->pos || !d->out_tokens)
      return set_error(TEAL_ERR_BAD_ARG, "step program: lm_head needs argmax_packed, token, pos, out_tokens");
    SpPhase& hp = open(kPhHead, -1);
    SpHead& h = hp.u.h;
    h.x = reinterpret_cast<const __half*>(d->x);
    h.norm_w = reinterpret_cast<const __half*>(d->final_norm);
    h.W = reinterpret_cast<const __half*>(d->lm_head);
    h.y = reinterpret_cast<__half*>(d->logits);
    h.argmax = reinterpret_cast<unsigned long long*>(d->argmax_packed);
    h.ldw = d->ld_head;
    h.Z = d->dim;
    h.V = d->vocab;
    h.rows_per_item = d->head_rows > 0 ? d->head_rows : 32;
    h.eps = (float)d->eps;
    hp.nitems = (d->vocab + h.rows_per_item - 1) / h.rows_per_item;
    hp.pub.count = hp.nitems;
    if (cur.valid) {
      hp.dep.kind = kDepAll;
      hp.dep.cell = cur.cell;
      hp.dep.per_epoch = cur.per_epoch;
      hp.dep.bcast0 = cur.bcast0;
    }
    const int head_cell = n, head_items = hp.nitems;
    close(hp, false);
    SpPhase& av = open(kPhAdvance, -1);
    av.nitems = 1;
    av.pub.count = 0;
    av.u.v.argmax = reinterpret_cast<unsigned long long*>(d->argmax_packed);
    av.u.v.token = const_cast<int*>(d->token);
    av.u.v.pos = d->pos;
    av.u.v.out_tokens = d->out_tokens;
    av.u.v.out_cap = d->out_cap;
    av.dep.kind = kDepAll;
    av.dep.cell = head_cell;
    av.dep.per_epoch = head_items;
    close(av, false);
  }
  if (n != P.nphases || flag0 > L.nflags) return set_error(TEAL_ERR_UNSUPPORTED, "step program: internal planning mismatch");

  H->magic = kSpMagic;
  H->nphases = n;
  H->total_items = item0;
  H->smem_bytes = (int)P.smem;
  H->ticket = reinterpret_cast<unsigned*>(sc + L.control);
  H->exited = H->ticket + 1;
  H->epoch = H->ticket + 2;
  H->flags = reinterpret_cast<unsigned*>(sc + L.flags);
  H->done = reinterpret_cast<unsigned*>(sc + L.done);
  H->nflags = L.nflags;
  H->ndone = L.ndone;
  H->scratch_bytes = L.total;
  H->ctas_per_sm = d->ctas_per_sm;
  H->poll_ns = d->poll_ns == 0 ? 32 : (d->poll_ns < 0 ? 0 : d->poll_ns);
  H->opts = d->opts;
  H->dbg = reinterpret_cast<unsigned long long*>(d->debug_times);
  H->warps = P.warps;
  H->stagger_ns = d->stagger_ns;
  H->ext = ((d->opts & ~1) != 0 || d->stagger_ns > 0 || d->tiles_per_item > 1) ? 1 : 0;
  H->bcast = use_bcast ? reinterpret_cast<unsigned long long*>(sc + L.bcast) : nullptr;
  return TEAL_OK;
}

const SpHeader* checked_header(const void* host_program) {
  if (!host_program) { set_error(TEAL_ERR_BAD_ARG, "step program: program is null"); return nullptr; }
  const SpHeader* H = reinterpret_cast<const SpHeader*>(host_program);
  if (H->magic != kSpMagic || H->nphases <= 0) { set_error(TEAL_ERR_BAD_ARG, "step program: not a planned program"); return nullptr; }
  return H;
}

}  // namespace

}  // namespace teal

using namespace teal;

extern "C" {

int teal_step_program_sizes(const teal_step_desc* desc, size_t* program_bytes, size_t* scratch_bytes) {
  Plan P;
  if (int rc = make_plan(desc, &P)) return rc;
  if (program_bytes) *program_bytes = sizeof(SpHeader) + (size_t)P.nphases * sizeof(SpPhase);
  if (scratch_bytes) *scratch_bytes = P.L.total;
  return TEAL_OK;
}

int teal_step_program_plan(const teal_step_desc* desc, void* host_program, size_t program_bytes) {
  Plan P;
  if (int rc = make_plan(desc, &P)) return rc;
  if (!host_program || program_bytes < sizeof(SpHeader) + (size_t)P.nphases * sizeof(SpPhase))
    return set_error(TEAL_ERR_BAD_ARG, "step program: program buffer too small");
  return build_program(desc, P, reinterpret_cast<unsigned char*>(host_program));
}

int teal_step_program_items(const void* host_program) {
  const SpHeader* H = checked_header(host_program);
  return H ? H->total_items : TEAL_ERR_BAD_ARG;
}

int teal_step_program_item(const void* host_program, int ticket, teal_item_info* out) {
  const SpHeader* H = checked_header(host_program);
  if (!H) return TEAL_ERR_BAD_ARG;
  if (!out || ticket < 0 || ticket >= H->total_items) return set_error(TEAL_ERR_BAD_ARG, "step program: ticket out of range");
  const SpPhase* ph = reinterpret_cast<const SpPhase*>(reinterpret_cast<const unsigned char*>(host_program) + sizeof(SpHeader));
  int pi = 0;
  while (ticket >= ph[pi].item0 + ph[pi].nitems) ++pi;
  const SpPhase& p = ph[pi];
  memset(out, 0, sizeof(*out));
  out->phase = pi;
  out->kind = p.kind;
  out->layer = p.layer;
  out->item = ticket - p.item0;
  out->ctile = out->kchunk = -1;
  out->pub_index = -1;
  out->pub_flag = -1;
  out->pub_cell = p.pub.cell;
  out->arrivals = 1;
  out->dep_kind = p.dep.kind;
  out->dep_cell = p.dep.cell;
  out->dep_per_epoch = p.dep.per_epoch;
  out->n_pubs = 0;
  if (p.kind == kPhGemv) {
    int ct, kc, nt = 1;
    if (p.u.g.tpi > 1) {
      int p0;
      sp_gemv_decode_multi(p.u.g, out->item, &p0, &nt, &kc);
      for (int t = 0; t < nt; ++t) out->ctiles[t] = sp_gemv_pos_tile(p.u.g, p0 + t);
      ct = out->ctiles[0];
    } else {
      sp_gemv_decode(p.u.g, out->item, &ct, &kc);
      out->ctiles[0] = ct;
    }
    out->ctile = ct;
    out->kchunk = kc;
    out->k_begin = kc * p.u.g.rows_per_item;
    out->k_end = out->k_begin + p.u.g.rows_per_item < p.u.g.Z ? out->k_begin + p.u.g.rows_per_item : p.u.g.Z;
    out->pub_index = sp_gemv_cidx(p.u.g, ct);
    out->arrivals = p.u.g.target;
    out->n_pubs = nt;
    for (int t = 0; t < nt; ++t) {
      out->pub_indices[t] = sp_gemv_cidx(p.u.g, out->ctiles[t]);
      out->pub_flags[t] = p.pub.flag0 >= 0 ? p.pub.flag0 + out->pub_indices[t] : -1;
    }
  } else if (p.kind == kPhAdvance) {
    out->pub_cell = -1;
  } else {
    out->pub_index = p.kind == kPhEmbed ? 0 : out->item;
  }
  if (p.pub.flag0 >= 0 && out->pub_index >= 0) out->pub_flag = p.pub.flag0 + out->pub_index;
  if (p.kind != kPhGemv && out->pub_index >= 0) {
    out->n_pubs = 1;
    out->pub_indices[0] = out->pub_index;
    out->pub_flags[0] = out->pub_flag;
    out->ctiles[0] = -1;
  }
  out->n_deps = sp_item_deps<true>(p, out->item, out->deps);
  return TEAL_OK;
}

int teal_step_program_run(const void* host_program, const void* device_program, void* stream) {
  const SpHeader* H = checked_header(host_program);
  if (!H) return TEAL_ERR_BAD_ARG;
  if (!device_program || (reinterpret_cast<uintptr_t>(device_program) & 15))
    return set_error(TEAL_ERR_ALIGNMENT, "step program: device program must be 16-byte aligned");
  const DeviceInfo& di = device_info();
  if (di.err != cudaSuccess) return set_cuda_error(di.err, "device query");
  // kernel instantiation: warps per CTA x {validated, experimental paths}
  using Kern = void (*)(const unsigned char*);
  static const Kern kKernels[2][2] = {{step_program_kernel<4, false>, step_program_kernel<4, true>},
                                      {step_program_kernel<2, false>, step_program_kernel<2, true>}};
  if (H->warps != 4 && H->warps != 2) return set_error(TEAL_ERR_BAD_ARG, "step program: bad warps per CTA in the plan");
  const int wi = H->warps == 4 ? 0 : 1, ei = H->ext ? 1 : 0;
  const Kern kern = kKernels[wi][ei];
  const int threads = H->warps * 32;
  static int s_smem_set[2][2] = {{0, 0}, {0, 0}};
  if (H->smem_bytes > 48 * 1024 && s_smem_set[wi][ei] < H->smem_bytes) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, H->smem_bytes);
    if (e != cudaSuccess) return set_cuda_error(e, "cudaFuncSetAttribute");
    s_smem_set[wi][ei] = H->smem_bytes;
  }
  static int s_occ[2][2] = {{0, 0}, {0, 0}}, s_occ_smem[2][2] = {{-1, -1}, {-1, -1}};
  if (s_occ_smem[wi][ei] != H->smem_bytes) {
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, (size_t)H->smem_bytes) != cudaSuccess || occ < 1) occ = 1;
    s_occ[wi][ei] = occ;
    s_occ_smem[wi][ei] = H->smem_bytes;
  }
  int per_sm = s_occ[wi][ei];
  if (H->ctas_per_sm > 0 && H->ctas_per_sm < per_sm) per_sm = H->ctas_per_sm;
  long long grid = (long long)di.sm_count * per_sm;
  if (grid > H->total_items) grid = H->total_items;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = (size_t)H->smem_bytes;
  cfg.stream = reinterpret_cast<cudaStream_t>(stream);
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, reinterpret_cast<const unsigned char*>(device_program));
  if (e != cudaSuccess) return set_cuda_error(e, "step_program launch");
  return TEAL_OK;
}

}  // extern "C"
