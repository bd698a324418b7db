// gather_gemv: the production fused kernel of the sparse decode path (sm_100a).
//
//   [optional RMSNorm of the residual stream] -> magnitude threshold -> survivor compaction in
//   shared memory -> gather-GEMV over only the surviving weight rows -> [fp16 | +residual |
//   SiLU(gate)*up] epilogue.
//
// One launch replaces, per projection, the reference's Y.zero_() + Triton split-K kernel
// (kernels/sparse_gemv.py:8-12,45-83 / :147-194) and, in the decode pipeline, the RMSNorm that
// precedes it and the residual add / SiLU*mul that follows it (gpt-fast/model.py:159-161,258-259).
//
// Data path (chosen by measurement, tools/micro/tma_rate.cu): 1-2 KB row
// segments are too small for TMA bulk copies (one UBLKCP per ~320 cycles per issuing warp), so
// every one of the 16 warps runs its own cp.async (LDGSTS, 16 B per lane, coalesced 512 B per
// instruction) ring and consumes what it fetched itself: no mbarriers, no producer/consumer
// hand-off, ~160 KB of loads in flight per SM.
//
// Work split: the host cuts the output columns into a small table of tiles (4 per matrix, 2+2 for
// the gate/up pair, so that 148 = 4 * 37 CTAs divide evenly) and assigns each tile a contiguous
// range of CTAs; the CTAs of a tile split that tile's compacted survivor list evenly, so every
// CTA owns exactly one (tile, survivor-range) piece.  Inside a CTA the 16 warps form CG column
// groups x RG row groups.  Row-group partials are reduced through shared memory, CTA partials
// with fp32 vector RED into an L2-resident accumulator; the last CTA to arrive on a tile's
// counter converts it, applies the epilogue and re-zeroes accumulator and counter.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "common.cuh"

namespace teal {

namespace {

constexpr int kGThreads = 512;
constexpr int kGWarps = kGThreads / 32;
constexpr int kMaxJ = 4;  // 16-byte vectors per lane per row (sub-tile <= 1024 columns)

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global.L2::128B [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
// Blackwell packed fp32 FMA (FFMA2): acc.xy += w.xy * x.xy -- halves the FMA issue slots of the
// instruction-bound inner loop (fp16 x fp16 products are exact in fp32, accumulation is fp32).
__device__ __forceinline__ void ffma2(float2& acc, const float2 w, const float2 x) {
  asm("fma.rn.f32x2 %0, %1, %2, %0;"
      : "+l"(*reinterpret_cast<unsigned long long*>(&acc))
      : "l"(*reinterpret_cast<const unsigned long long*>(&w)), "l"(*reinterpret_cast<const unsigned long long*>(&x)));
}

}  // namespace

// DEPTH: rows in flight per warp.
template <int DEPTH>
__global__ void __launch_bounds__(kGThreads, 1) gather_gemv_kernel(const __grid_constant__ GatherParams p) {
  extern __shared__ __align__(128) unsigned char smem[];
  // layout: [rings (16 warps * DEPTH * slot) aliased by red (RG * tile_w floats)][list_idx][list_x][xs]
  float* red = reinterpret_cast<float*>(smem);
  int* list_idx = reinterpret_cast<int*>(smem + p.region_bytes);
  float* list_x = reinterpret_cast<float*>(smem + p.region_bytes + (size_t)p.list_cap * 4);
  uint4* xs_vec = reinterpret_cast<uint4*>(smem + p.region_bytes + (size_t)p.list_cap * 8);

  __shared__ int s_warp_cnt[kGWarps];
  __shared__ float s_part[kGWarps];
  __shared__ int s_last;

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  long long* const dbg = p.dbg_time ? p.dbg_time + (size_t)blockIdx.x * 16 : nullptr;
#define TEAL_STAMP(i) do { if (dbg && tid == 0) dbg[i] = clock64(); } while (0)
  if (dbg && tid == 0) dbg[13] = globaltimer_ns();
  TEAL_STAMP(0);

  // ---- 0. which tile is mine (host-built table, <= kMaxTiles entries) ---------------------------
  int ti = 0;
#pragma unroll
  for (int u = 1; u < kMaxTiles; ++u) ti += (u < p.ntiles && (int)blockIdx.x >= p.tile[u].cta0) ? 1 : 0;
  const GatherTile& tl = p.tile[ti];
  const int jct = (int)blockIdx.x - tl.cta0;  // my index among the tile's CTAs
  const float thr = p.thr[tl.thr_slot];

  // Programmatic dependent launch: nothing above depends on the previous kernel.
  griddep_wait();

  // ---- 1. x: load this thread's chunk, optional RMSNorm, stage in smem, count survivors --------
  // thread `tid` owns lanes [c0, c1) of x; p.chunk is a multiple of 8.
  const int c0 = min(tid * p.chunk, p.Z);
  const int c1 = min(c0 + p.chunk, p.Z);
  if (p.in_mode == kInRmsNorm) {
    // gpt-fast/model.py:286-291: fp32 norm, cast to fp16, then multiply by the fp16 weight
    float ss = 0.f;
    for (int m = c0; m < c1; m += 8) {
      uint4 v = ld_nc_u4(reinterpret_cast<const uint4*>(p.x + m));
      xs_vec[m >> 3] = v;
      const __half2* h2 = reinterpret_cast<const __half2*>(&v);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float2 f = __half22float2(h2[j]);
        ss = fmaf(f.x, f.x, ss);
        ss = fmaf(f.y, f.y, ss);
      }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, d);
    if (lane == 0) s_part[warp] = ss;
    __syncthreads();
    float tot = (lane < kGWarps) ? s_part[lane] : 0.f;
#pragma unroll
    for (int d = 8; d > 0; d >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, d);
    tot = __shfl_sync(0xffffffffu, tot, 0);
    const float rs = rsqrtf(tot / (float)p.Z + p.eps);
    for (int m = c0; m < c1; m += 8) {
      uint4 v = xs_vec[m >> 3];
      uint4 g = ld_nc_u4(reinterpret_cast<const uint4*>(p.norm_w + m));
      __half2* h2 = reinterpret_cast<__half2*>(&v);
      const __half2* g2 = reinterpret_cast<const __half2*>(&g);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float2 f = __half22float2(h2[j]);
        h2[j] = __hmul2(__floats2half2_rn(f.x * rs, f.y * rs), g2[j]);
      }
      xs_vec[m >> 3] = v;
    }
  } else {
    for (int m = c0; m < c1; m += 8) xs_vec[m >> 3] = ld_nc_u4(reinterpret_cast<const uint4*>(p.x + m));
  }
  if (p.x_out && blockIdx.x == 0)  // optional side output: the (normalised) x this op consumed
    for (int m = c0; m < c1; m += 8) *reinterpret_cast<uint4*>(p.x_out + m) = xs_vec[m >> 3];
  TEAL_STAMP(1);

  // keep-mask of my chunk (bit i <-> lane c0 + i; chunk <= 128 lanes because Z <= 65536)
  uint32_t keep[4] = {0u, 0u, 0u, 0u};
  int cnt = 0;
  for (int m = c0; m < c1; m += 8) {
    uint4 v = xs_vec[m >> 3];
    const __half2* h2 = reinterpret_cast<const __half2*>(&v);
    uint32_t bits = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float2 f = __half22float2(h2[j]);
      bits |= (uint32_t)!(fabsf(f.x) <= thr) << (2 * j);
      bits |= (uint32_t)!(fabsf(f.y) <= thr) << (2 * j + 1);
    }
    const int o = m - c0;
    const uint32_t sh = bits << (o & 31);
    keep[0] |= (o < 32) ? sh : 0u;
    keep[1] |= (o >= 32 && o < 64) ? sh : 0u;
    keep[2] |= (o >= 64 && o < 96) ? sh : 0u;
    keep[3] |= (o >= 96) ? sh : 0u;
    cnt += __popc(bits);
  }
  int incl = cnt;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int o = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += o;
  }
  if (lane == 31) s_warp_cnt[warp] = incl;
  __syncthreads();
  int pre, nnz;
  {
    const int c = (lane < kGWarps) ? s_warp_cnt[lane] : 0;
    pre = incl - cnt + __reduce_add_sync(0xffffffffu, (lane < warp) ? c : 0);
    nnz = __reduce_add_sync(0xffffffffu, c);
  }
  TEAL_STAMP(2);

  // measurement probe: gathered vs dense weight bytes of this launch (one CTA per tile reports)
  if (p.probe && jct == 0 && tid == 0 && p.mode != kModeSurvivors) {
    atomicAdd(p.probe, (unsigned long long)nnz * (unsigned long long)tl.width * 2ull);
    atomicAdd(p.probe + 1, (unsigned long long)p.Z * (unsigned long long)tl.width * 2ull);
  }
  if (p.mode == kModeSurvivors) {  // parity hook: dump the survivor list
    if (blockIdx.x == 0) {
      int rank = pre;
#pragma unroll
      for (int w = 0; w < 4; ++w) {
        uint32_t b = keep[w];
        while (b) {
          const int i = __ffs(b) - 1;
          b &= b - 1;
          p.dbg_idx[rank++] = c0 + 32 * w + i;
        }
      }
      if (tid == 0) p.dbg_cnt[0] = nnz;
    }
    return;
  }

  // ---- 2. my survivor range: the tile's nctas CTAs split [0, nnz) evenly -------------------------
  // floor(nnz * j / n) via the host's magic = ceil(2^32 / n): exact for nnz * j < 2^32 / n (< 2^24 here)
  const int ra = tl.magic ? (int)__umulhi((unsigned)(nnz * jct), tl.magic) : nnz * jct;
  const int rb = tl.magic ? (int)__umulhi((unsigned)(nnz * (jct + 1)), tl.magic) : nnz * (jct + 1);

  // warp roles inside the CTA tile: cgi = column group (sub-tile), rgi = row group
  const int cgi = warp % p.cg;
  const int rgi = warp / p.cg;
  const int sub_col0 = cgi * p.sub_n;
  const int scols = max(0, min(p.sub_n, tl.width - sub_col0));
  const int nvec = scols >> 3;  // 16-byte vectors per row of my sub-tile
  const int slot_bytes = p.sub_n * 2;
  const uint32_t ring_u32 = smem_u32(smem) + (uint32_t)warp * DEPTH * slot_bytes + lane * 16;
  const unsigned char* ring_ptr = smem + (size_t)warp * DEPTH * slot_bytes + lane * 16;
  const __half* wbase = tl.w + sub_col0 + lane * 8;
  bool vj[kMaxJ];
#pragma unroll
  for (int j = 0; j < kMaxJ; ++j) vj[j] = lane + 32 * j < nvec;

  float2 acc[kMaxJ][4];
#pragma unroll
  for (int j = 0; j < kMaxJ; ++j)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[j][c] = make_float2(0.f, 0.f);

  for (int r0 = ra; r0 < rb; r0 += p.list_cap) {  // one round unless the range exceeds the list capacity
    const int r1 = min(rb, r0 + p.list_cap);
    // ---- 2a. windowed compaction: list[rank - r0] = (m, x[m]) for rank in [r0, r1) ---------------
    if (pre < r1 && pre + cnt > r0) {
      int rank = pre;
#pragma unroll
      for (int w = 0; w < 4; ++w) {
        uint32_t b = keep[w];
        while (b) {
          const int i = __ffs(b) - 1;
          b &= b - 1;
          if (rank >= r0 && rank < r1) {
            const int m = c0 + 32 * w + i;
            list_idx[rank - r0] = m;
            list_x[rank - r0] = __half2float(reinterpret_cast<const __half*>(xs_vec)[m]);
          }
          ++rank;
        }
      }
    }
    __syncthreads();
    TEAL_STAMP(3);

    // ---- 3. per-warp cp.async pipeline: my rows are r = rgi, rgi + RG, ... of the window ----------
    const int nwin = r1 - r0;
    const int nrows = (nwin > rgi) ? (nwin - rgi + p.rg - 1) / p.rg : 0;
    const int* lp = list_idx + rgi;
    const float* xp = list_x + rgi;
    const long long ldw = tl.ldw;
    int islot = 0;  // ring slot the next issued row goes to
    auto issue = [&](int i) {
      if (i < nrows) {
        const __half* src = wbase + (long long)lp[i * p.rg] * ldw;
        const uint32_t dst = ring_u32 + (uint32_t)islot * slot_bytes;
#pragma unroll
        for (int j = 0; j < kMaxJ; ++j)
          if (vj[j]) cp_async16(dst + j * 512, src + j * 256);
      }
      cp_async_commit();
      islot = (islot + 1 == DEPTH) ? 0 : islot + 1;
    };
#pragma unroll
    for (int i = 0; i < DEPTH - 1; ++i) issue(i);
    int cslot = 0;  // ring slot of the row being consumed
    for (int i = 0; i < nrows; ++i) {
      issue(i + DEPTH - 1);        // its slot was released by the __syncwarp at the end of row i - 1
      cp_async_wait<DEPTH - 1>();  // row i has landed (for this lane) ...
      __syncwarp();                // ... and for every lane of the warp
      const float xv = xp[i * p.rg];
      const float2 x2 = make_float2(xv, xv);
      const unsigned char* sb = ring_ptr + (size_t)cslot * slot_bytes;
#pragma unroll
      for (int j = 0; j < kMaxJ; ++j) {
        if (vj[j]) {
          const uint4 v = *reinterpret_cast<const uint4*>(sb + j * 512);
          ffma2(acc[j][0], __half22float2(*reinterpret_cast<const __half2*>(&v.x)), x2);
          ffma2(acc[j][1], __half22float2(*reinterpret_cast<const __half2*>(&v.y)), x2);
          ffma2(acc[j][2], __half22float2(*reinterpret_cast<const __half2*>(&v.z)), x2);
          ffma2(acc[j][3], __half22float2(*reinterpret_cast<const __half2*>(&v.w)), x2);
        }
      }
      __syncwarp();                // all lanes done with the slot before it is refilled
      cslot = (cslot + 1 == DEPTH) ? 0 : cslot + 1;
      if (p.dbg_time && lane == 0 && (warp == 0 || warp == 15) && i < 32)
        p.dbg_time[(size_t)gridDim.x * 16 + (size_t)blockIdx.x * 64 + (warp ? 32 : 0) + i] = clock64();
    }
    cp_async_wait<0>();
    __syncthreads();  // rings drained (red aliases them); lists may be rewritten by the next round
  }
  // let the next kernel in the stream start launching as our CTAs retire
  griddep_launch_dependents();
  TEAL_STAMP(6);

  // ---- 4. combine: row groups -> CTA partial (smem) -> fp32 RED into the L2 accumulator ------------
  if (rb > ra) {
    if (p.rg > 1) {
#pragma unroll
      for (int j = 0; j < kMaxJ; ++j) {
        if (vj[j]) {
          float* dst = red + (size_t)rgi * p.tile_n + sub_col0 + (lane + 32 * j) * 8;
          *reinterpret_cast<float4*>(dst) = make_float4(acc[j][0].x, acc[j][0].y, acc[j][1].x, acc[j][1].y);
          *reinterpret_cast<float4*>(dst + 4) = make_float4(acc[j][2].x, acc[j][2].y, acc[j][3].x, acc[j][3].y);
        }
      }
      __syncthreads();
      const int rg_used = min(p.rg, rb - ra);
      for (int c = tid * 4; c < tl.width; c += kGThreads * 4) {
        float4 s = *reinterpret_cast<const float4*>(&red[c]);
        for (int g = 1; g < rg_used; ++g) {
          float4 t = *reinterpret_cast<const float4*>(&red[(size_t)g * p.tile_n + c]);
          s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
        }
        red_add_v4(p.acc + tl.acc_col0 + c, s);
      }
    } else {
#pragma unroll
      for (int j = 0; j < kMaxJ; ++j) {
        if (vj[j]) {
          float* dst = p.acc + tl.acc_col0 + sub_col0 + (lane + 32 * j) * 8;
          red_add_v4(dst, make_float4(acc[j][0].x, acc[j][0].y, acc[j][1].x, acc[j][1].y));
          red_add_v4(dst + 4, make_float4(acc[j][2].x, acc[j][2].y, acc[j][3].x, acc[j][3].y));
        }
      }
    }
  }
  // ---- 5. completion counter (counts CTA arrivals; target is host-known): last arrival finalises ----
  // REDs of all threads happen-before the barrier; thread 0's release fence is cumulative over them,
  // its acquire fence orders the finaliser's loads after the counter read.
  __syncthreads();
  TEAL_STAMP(7);
  if (tid == 0) {
    fence_acq_rel_gpu();
    const unsigned old = atomicAdd(&p.counters[tl.cidx], 1u);
    const int last = ((int)old + 1 == tl.target) ? 1 : 0;
    if (last) fence_acq_rel_gpu();
    s_last = last;
  }
  __syncthreads();
  TEAL_STAMP(8);
  if (s_last) {
    if (p.mode == kModePairSilu) {
      for (int c = tid * 4; c < tl.width; c += kGThreads * 4) {
        float4 g = ld_cg_f4(p.acc + tl.gate_acc0 + c);
        float4 u = ld_cg_f4(p.acc + tl.up_acc0 + c);
        st_cg_f4(p.acc + tl.gate_acc0 + c, make_float4(0.f, 0.f, 0.f, 0.f));
        st_cg_f4(p.acc + tl.up_acc0 + c, make_float4(0.f, 0.f, 0.f, 0.f));
        __half2 h0 = __halves2half2(silu_mul_f16(g.x, u.x), silu_mul_f16(g.y, u.y));
        __half2 h1 = __halves2half2(silu_mul_f16(g.z, u.z), silu_mul_f16(g.w, u.w));
        uint2 o;
        o.x = *reinterpret_cast<uint32_t*>(&h0);
        o.y = *reinterpret_cast<uint32_t*>(&h1);
        *reinterpret_cast<uint2*>(p.y + tl.y_col0 + c) = o;
      }
    } else {
      for (int c = tid * 4; c < tl.width; c += kGThreads * 4) {
        float4 v = ld_cg_f4(p.acc + tl.acc_col0 + c);
        st_cg_f4(p.acc + tl.acc_col0 + c, make_float4(0.f, 0.f, 0.f, 0.f));
        __half2 h0 = __floats2half2_rn(v.x, v.y);
        __half2 h1 = __floats2half2_rn(v.z, v.w);
        if (p.out_mode == kOutAddResidual) {  // h = x + y in fp16 (gpt-fast/model.py:160-161)
          const uint2 r = *reinterpret_cast<const uint2*>(p.resid + tl.y_col0 + c);
          h0 = __hadd2(*reinterpret_cast<const __half2*>(&r.x), h0);
          h1 = __hadd2(*reinterpret_cast<const __half2*>(&r.y), h1);
        }
        uint2 o;
        o.x = *reinterpret_cast<uint32_t*>(&h0);
        o.y = *reinterpret_cast<uint32_t*>(&h1);
        *reinterpret_cast<uint2*>(p.y + tl.y_col0 + c) = o;
      }
    }
    if (tid == 0) p.counters[tl.cidx] = 0;
  }
  TEAL_STAMP(9);
  TEAL_STAMP(10);
  if (dbg && tid == 0) {
    dbg[14] = globaltimer_ns();
    dbg[15] = 1ll | ((long long)(rb - ra) << 8);
  }
#undef TEAL_STAMP
}

// ------------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------------
static void (*pick_kernel(int depth))(const GatherParams) {
  switch (depth) {
    case 4: return gather_gemv_kernel<4>;
    case 6: return gather_gemv_kernel<6>;
    case 8: return gather_gemv_kernel<8>;
    case 10: return gather_gemv_kernel<10>;
    default: return gather_gemv_kernel<12>;
  }
}

// Builds the tile table from the segment description shared with the TMA kernel (GemvParams) and launches.
int launch_gather_gemv(GemvParams& gp, void* workspace, size_t ws_bytes, int out_cols_total, cudaStream_t stream,
                       int tiles_per_matrix) {
  const DeviceInfo& di = device_info();
  if (di.err != cudaSuccess) return set_cuda_error(di.err, "device query");
  if (gp.Z <= 0 || gp.Z > 65536 || (gp.Z & 7)) return set_error(TEAL_ERR_ALIGNMENT, "Z must be in (0, 65536] and a multiple of 8");
  if ((reinterpret_cast<uintptr_t>(gp.x) & 15)) return set_error(TEAL_ERR_ALIGNMENT, "x must be 16-byte aligned");
  GatherParams p;
  memset(&p, 0, sizeof(p));
  p.x = gp.x; p.y = gp.y; p.norm_w = gp.norm_w; p.resid = gp.resid; p.x_out = gp.x_out;
  p.dbg_idx = gp.dbg_idx; p.dbg_cnt = gp.dbg_cnt; p.dbg_time = debug_timestamps(); p.probe = bytes_probe();
  p.Z = gp.Z; p.mode = gp.mode; p.in_mode = gp.in_mode; p.out_mode = gp.out_mode; p.eps = gp.eps;
  for (int k = 0; k < kMaxThr; ++k) p.thr[k] = gp.thr[k];
  if (gp.mode != kModeSurvivors) {
    if (!workspace || ws_bytes < workspace_bytes(out_cols_total)) return set_error(TEAL_ERR_WORKSPACE, "workspace too small");
    if ((reinterpret_cast<uintptr_t>(gp.y) & 15)) return set_error(TEAL_ERR_ALIGNMENT, "y must be 16-byte aligned");
    p.counters = reinterpret_cast<unsigned*>(workspace);
    p.acc = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(workspace) + (size_t)kMaxCounters * 4);
  }
  if (gp.in_mode == kInRmsNorm && (!gp.norm_w || (reinterpret_cast<uintptr_t>(gp.norm_w) & 15)))
    return set_error(TEAL_ERR_ALIGNMENT, "norm weight must be non-null and 16-byte aligned");
  if (gp.out_mode == kOutAddResidual && (!gp.resid || (reinterpret_cast<uintptr_t>(gp.resid) & 7)))
    return set_error(TEAL_ERR_ALIGNMENT, "residual must be non-null and 8-byte aligned");

  const int G = (gp.mode == kModeSurvivors) ? 1 : di.sm_count;
  // ---- tile table: `tiles_per_matrix` equal tiles per segment (fewer if the segment is narrow) ------
  const bool pair = gp.mode == kModePairSilu;
  int total_cols = 0;
  for (int s = 0; s < gp.nseg; ++s) {
    const GemvSeg& sg = gp.seg[s];
    if (sg.ncols <= 0 || (sg.ncols & 7) || (sg.ldw & 7) || (sg.out_col0 & 7) || (reinterpret_cast<uintptr_t>(sg.w) & 15))
      return set_error(TEAL_ERR_ALIGNMENT, "weights: N, ldw, region offsets must be multiples of 8 and 16-byte aligned");
    total_cols += sg.ncols;
  }
  int nt = 0, max_w = 8;
  if (gp.mode == kModeSurvivors) {
    p.tile[0] = GatherTile{};
    p.tile[0].nctas = 1;
    p.tile[0].magic = 0;
    nt = 1;
  } else {
    // tiles per segment proportional to its width, at least one, `tiles_per_matrix * nseg` in total at most
    const int want_total = pair ? 2 * tiles_per_matrix : (gp.nseg == 1 ? tiles_per_matrix : tiles_per_matrix + gp.nseg - 1);
    for (int s = 0; s < gp.nseg; ++s) {
      const GemvSeg& sg = gp.seg[s];
      int k = pair ? tiles_per_matrix : (int)(((long long)want_total * sg.ncols + total_cols / 2) / total_cols);
      if (k < 1) k = 1;
      while (k > 1 && sg.ncols / k < 64) --k;
      if (nt + k > kMaxTiles) k = kMaxTiles - nt;
      if (k < 1) return set_error(TEAL_ERR_UNSUPPORTED, "too many column regions");
      const int w = ((sg.ncols + k - 1) / k + 7) & ~7;
      for (int t = 0; t < k; ++t) {
        const int col0 = t * w;
        if (col0 >= sg.ncols) break;
        GatherTile& tl = p.tile[nt];
        tl = GatherTile{};
        tl.w = sg.w + col0;
        tl.ldw = sg.ldw;
        tl.width = (sg.ncols - col0 < w) ? sg.ncols - col0 : w;
        tl.acc_col0 = sg.out_col0 + col0;
        tl.y_col0 = pair ? col0 : sg.out_col0 + col0;
        tl.thr_slot = sg.thr_slot;
        tl.cidx = pair ? t : nt;
        if (pair) { tl.gate_acc0 = gp.seg[0].out_col0 + col0; tl.up_acc0 = gp.seg[1].out_col0 + col0; }
        if (tl.width > max_w) max_w = tl.width;
        ++nt;
      }
    }
    // CTAs per tile proportional to width (largest remainder), every tile gets at least one
    long long wsum = 0;
    for (int u = 0; u < nt; ++u) wsum += p.tile[u].width;
    int given = 0;
    for (int u = 0; u < nt; ++u) {
      int n = (int)((long long)G * p.tile[u].width / wsum);
      if (n < 1) n = 1;
      p.tile[u].nctas = n;
      given += n;
    }
    for (int u = 0; given < G; u = (u + 1) % nt) { ++p.tile[u].nctas; ++given; }
    for (int u = 0; given > G; u = (u + 1) % nt)
      if (p.tile[u].nctas > 1) { --p.tile[u].nctas; --given; }
    int cta0 = 0;
    for (int u = 0; u < nt; ++u) {
      GatherTile& tl = p.tile[u];
      tl.cta0 = cta0;
      cta0 += tl.nctas;
      tl.magic = (tl.nctas == 1) ? 0u : (unsigned)((0x100000000ull + tl.nctas - 1) / tl.nctas);
    }
    for (int u = 0; u < nt; ++u) {
      GatherTile& tl = p.tile[u];
      if (pair) {  // one counter per (gate tile t, up tile t) pair
        int tgt = 0;
        for (int v = 0; v < nt; ++v)
          if (p.tile[v].cidx == tl.cidx) tgt += p.tile[v].nctas;
        tl.target = tgt;
      } else {
        tl.target = tl.nctas;
      }
    }
  }
  p.ntiles = nt;
  // sub-tiles: <= kSubCols columns per warp (kSubCols / 256 x 16 B per lane per row)
  const int kSubCols = sub_cols_override() ? sub_cols_override() : 1024;
  int tile_n = (max_w + 7) & ~7;
  int cg = 1;
  while (cg < kGWarps && (tile_n + cg - 1) / cg > kSubCols) cg <<= 1;
  if ((tile_n + cg - 1) / cg > kSubCols) return set_error(TEAL_ERR_UNSUPPORTED, "column tile too wide");
  p.tile_n = tile_n;
  p.cg = cg;
  p.rg = kGWarps / cg;
  p.sub_n = ((tile_n + cg - 1) / cg + 7) & ~7;
  p.chunk = (((p.Z + kGThreads - 1) / kGThreads) + 7) & ~7;
  p.list_cap = 2048;
  const size_t budget = (size_t)di.smem_optin - 1024;
  const size_t fixed = (size_t)p.list_cap * 8 + (size_t)((p.Z + 7) & ~7) * 2 + 128;
  const size_t red_bytes = (size_t)p.rg * tile_n * 4;
  int depth = 12;
  while (depth > 4 && fixed + (size_t)kGWarps * depth * p.sub_n * 2 > budget) depth -= 2;
  size_t region = (size_t)kGWarps * depth * p.sub_n * 2;
  if (region < red_bytes) region = red_bytes;
  region = (region + 127) & ~(size_t)127;
  if (fixed + region > budget) return set_error(TEAL_ERR_UNSUPPORTED, "shape needs too much shared memory");
  p.region_bytes = (int)region;
  const size_t smem = fixed + region;

  auto kern = pick_kernel(depth);
  static bool attr_set[16] = {};
  if (!attr_set[depth]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget);
    if (e != cudaSuccess) return set_cuda_error(e, "cudaFuncSetAttribute");
    attr_set[depth] = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(G);
  cfg.blockDim = dim3(kGThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, p);
  if (e != cudaSuccess) return set_cuda_error(e, "gather_gemv launch");
  return TEAL_OK;
}

}  // namespace teal
