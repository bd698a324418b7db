// Fused threshold -> survivor compaction -> gather-GEMV for sm_100a (B200).
//
// Replaces the reference's Triton split-K kernels (kernels/sparse_gemv.py:45-83, :147-194 in
// FasterDecoding/TEAL).  Not a translation: the reference issues masked tile loads over a fixed
// (N-tile x K-tile) grid and accumulates with fp16 global atomics; this kernel
//   1. stages x in shared memory, thresholds it and block-scans the keep flags, so that every
//      CTA knows nnz and the rank of every surviving lane (compaction),
//   2. flattens the work (column tile x survivor rank) and gives each of the #SM CTAs an equal
//      contiguous slice of it -- perfect balance for any shape / sparsity,
//   3. a producer warp gathers only the surviving weight-row segments from HBM with 1-D TMA bulk
//      copies (cp.async.bulk + mbarrier complete_tx, L2 evict-first) into a deep smem ring,
//   4. 16 consumer warps FMA them in fp32 against the compacted x values,
//   5. CTA partials are combined with fp32 vector RED into an L2-resident accumulator; the CTA
//      whose contribution completes a tile converts it to fp16 (optionally SiLU(gate)*up),
//      and re-zeroes the accumulator, so no memset / zero-fill launch is ever needed.
// No tensor cores: batch-1 GEMV is 1 FMA per 2 bytes, HBM-bound.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace teal {

// ------------------------------------------------------------------------------------------
// Kernel
// ------------------------------------------------------------------------------------------
// register-array element select without dynamic indexing (keeps the arrays out of local memory)
template <typename T, int N>
__device__ __forceinline__ T sel(const T (&arr)[N], int i) {
  T v = arr[0];
#pragma unroll
  for (int j = 1; j < N; ++j) v = (i == j) ? arr[j] : v;
  return v;
}

// One "piece" = the part of one column tile's survivor list that this CTA processes.  Lane j of
// every warp holds piece j of the current round in registers; fields are broadcast with shuffles.
struct Piece {
  int seg;        // segment index
  int tile;       // tile index inside the segment
  int take;       // flat work units contributed to the tile's completion counter
  int real_rows;  // rows actually gathered (0 for the dummy unit of an all-dropped segment)
  int list_off;   // offset of this piece's survivors in list_idx / list_x
  int a;          // first survivor rank of the window
  int cidx;       // completion counter index
  int target;     // counter value that completes the tile
};
__device__ __forceinline__ Piece piece_bcast(const Piece& mine, int src_lane) {
  Piece r;
  r.seg = __shfl_sync(0xffffffffu, mine.seg, src_lane);
  r.tile = __shfl_sync(0xffffffffu, mine.tile, src_lane);
  r.take = __shfl_sync(0xffffffffu, mine.take, src_lane);
  r.real_rows = __shfl_sync(0xffffffffu, mine.real_rows, src_lane);
  r.list_off = __shfl_sync(0xffffffffu, mine.list_off, src_lane);
  r.a = __shfl_sync(0xffffffffu, mine.a, src_lane);
  r.cidx = __shfl_sync(0xffffffffu, mine.cidx, src_lane);
  r.target = __shfl_sync(0xffffffffu, mine.target, src_lane);
  return r;
}

template <int CPT, int GROUPS, int R, int NT>
__global__ void __launch_bounds__(kThreads, 1)
sparse_gemv_kernel(const __grid_constant__ GemvParams p) {
  constexpr int TPG = kConsumerThreads / GROUPS;  // threads per row-group
  constexpr int TILE_N = TPG * CPT;               // columns per tile
  constexpr int ROW_BYTES = TILE_N * 2;
  constexpr int STAGE_BYTES = R * ROW_BYTES;
  constexpr int RPG = (R + GROUPS - 1) / GROUPS;  // rows per group per stage
  constexpr int kWarps = kThreads / 32;
  constexpr int kConsumerWarps = kConsumerThreads / 32;
  static_assert(R <= 32, "one producer lane per row");
  static_assert(CPT == 4 || CPT == 8, "CPT");
  static_assert(NT == 1 || NT == kMaxThr, "NT");

  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* ring = smem;
  size_t off = (size_t)p.nstages * STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + off);
  off += kMaxStages * 8;
  uint64_t* empty_bar = reinterpret_cast<uint64_t*>(smem + off);
  off += kMaxStages * 8;
  float* red = reinterpret_cast<float*>(smem + off);
  off += (GROUPS > 1) ? (size_t)GROUPS * TILE_N * 4 : 0;
  int* list_idx = reinterpret_cast<int*>(smem + off);
  off += (size_t)p.list_cap * 4;
  float* list_x = reinterpret_cast<float*>(smem + off);
  off += (size_t)p.list_cap * 4;
  uint4* xs_vec = reinterpret_cast<uint4*>(smem + off);  // fp16 x, p.xs_elems halves

  __shared__ int s_warp_cnt[NT][kWarps];
  __shared__ int s_last;

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  long long* const dbg = p.dbg_time ? p.dbg_time + (size_t)blockIdx.x * 16 : nullptr;
#define TEAL_STAMP(i) do { if (dbg && tid == 0) dbg[i] = clock64(); } while (0)
  if (dbg && tid == 0) { dbg[13] = globaltimer_ns(); }
  TEAL_STAMP(0);

  if (tid < p.nstages) {
    mbar_init(smem_u32(&full_bar[tid]), 1);
    mbar_init(smem_u32(&empty_bar[tid]), kConsumerWarps);
    fence_mbar_init();
  }
  // Programmatic dependent launch: everything above overlaps the previous kernel's tail.
  griddep_wait();

  // ---- 1. stage x (fp16) in shared memory ------------------------------------------------
  {
    const uint4* xg = reinterpret_cast<const uint4*>(p.x);
    const int nvec = p.Z >> 3;
    for (int v = tid; v < nvec; v += kThreads) xs_vec[v] = ld_nc_u4(xg + v);
  }
  __syncthreads();
  TEAL_STAMP(1);

  // ---- 2. threshold + block scan: rank of every surviving lane ----------------------------
  // thread `tid` (< 512) owns lanes [c0, c1) of x; p.chunk is a multiple of 8.
  const int c0 = min(tid * p.chunk, p.Z);
  const int c1 = (tid < kConsumerThreads) ? min(c0 + p.chunk, p.Z) : c0;
  float thr[NT];
  int cnt[NT], pre[NT], nnz[NT];
#pragma unroll
  for (int k = 0; k < NT; ++k) { thr[k] = p.thr[k]; cnt[k] = 0; }
  for (int m = c0; m < c1; m += 8) {
    uint4 v = xs_vec[m >> 3];
    const __half2* h2 = reinterpret_cast<const __half2*>(&v);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float2 f = __half22float2(h2[j]);
      float a0 = fabsf(f.x), a1 = fabsf(f.y);
#pragma unroll
      for (int k = 0; k < NT; ++k) cnt[k] += (int)!(a0 <= thr[k]) + (int)!(a1 <= thr[k]);
    }
  }
#pragma unroll
  for (int k = 0; k < NT; ++k) {
    int incl = cnt[k];
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      int o = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += o;
    }
    pre[k] = incl - cnt[k];
    if (lane == 31) s_warp_cnt[k][warp] = incl;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < NT; ++k) {
    // consumer warps 0..15 hold the counts: one lane-parallel read + butterfly reductions
    int c = (lane < kConsumerWarps) ? s_warp_cnt[k][lane] : 0;
    int below = (lane < warp) ? c : 0;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      c += __shfl_xor_sync(0xffffffffu, c, d);
      below += __shfl_xor_sync(0xffffffffu, below, d);
    }
    pre[k] += below;
    nnz[k] = c;
  }
  TEAL_STAMP(2);

  if (p.mode == kModeSurvivors) {  // parity hook: dump the survivor list of threshold slot 0
    if (blockIdx.x == 0) {
      int rank = pre[0];
      for (int m = c0; m < c1; ++m) {
        float a = fabsf(__half2float(reinterpret_cast<const __half*>(xs_vec)[m]));
        if (!(a <= thr[0])) p.dbg_idx[rank++] = m;
      }
      if (tid == 0) p.dbg_cnt[0] = nnz[0];
    }
    return;
  }

  // ---- 3. flat work split ------------------------------------------------------------------
  // unit = (tile, survivor rank); a tile of an all-dropped segment still has one dummy unit so
  // that somebody finalises (zero-fills) its outputs.  W <= 4096 tiles * 65536 rows < 2^31.
  const bool pair = (p.mode == kModePairSilu);
  int rows[kMaxSeg], snnz[kMaxSeg], segpre[kMaxSeg + 1], tile0[kMaxSeg + 1];
  segpre[0] = 0;
  tile0[0] = 0;
#pragma unroll
  for (int s = 0; s < kMaxSeg; ++s) {
    const bool on = s < p.nseg;
    snnz[s] = on ? sel(nnz, p.seg[s].thr_slot) : 0;
    rows[s] = on ? max(snnz[s], 1) : 0;
    const int nt = on ? p.seg[s].ntiles : 0;
    segpre[s + 1] = segpre[s] + nt * rows[s];
    tile0[s + 1] = tile0[s] + nt;
  }
  const int pair_rows = rows[0] + rows[1];
  const int U = tile0[kMaxSeg];  // total tiles (pair mode: 2 * ntiles)
  const int W = pair ? p.seg[0].ntiles * pair_rows : segpre[kMaxSeg];
  const int f0 = (int)((unsigned long long)W * blockIdx.x / gridDim.x);
  const int f1 = (int)((unsigned long long)W * (blockIdx.x + 1) / gridDim.x);

  const uint32_t ring_u32 = smem_u32(ring);
  const uint32_t full_u32 = smem_u32(full_bar);
  const uint32_t empty_u32 = smem_u32(empty_bar);
  int stage = 0;       // ring slot of the next stage in sequence (same sequence in every warp)
  uint32_t phase = 0;  // consumer parity; producers wait on phase ^ 1
  int seq = 0;         // stage sequence number, selects the owning producer warp
  const uint64_t policy = make_evict_first_policy();

  const int group = (tid < kConsumerThreads) ? tid / TPG : 0;
  const int tcol = (tid % TPG) * CPT;
  const int pw = warp - kConsumerWarps;  // producer warp index (>= 0 for producers)
  bool first_stamp = true;

  int f = f0;
  while (f < f1) {
    // ---- 3a. piece table: lane j computes the j-th tile intersecting [f, f1) -----------------
    Piece mine;
    int np, fnext;
    {
      int s, t, u_first;
      if (pair) {
        t = f / pair_rows;
        s = (f - t * pair_rows >= rows[0]) ? 1 : 0;
        u_first = 2 * t + s;
      } else {
        s = 0;
#pragma unroll
        for (int j = 1; j < kMaxSeg; ++j) s += (j < p.nseg && f >= segpre[j]) ? 1 : 0;
        t = (f - sel(segpre, s)) / sel(rows, s);
        u_first = sel(tile0, s) + t;
      }
      const int u = u_first + lane;
      int rs, start;
      if (pair) {
        s = u & 1;
        t = u >> 1;
        rs = s ? rows[1] : rows[0];
        start = t * pair_rows + (s ? rows[0] : 0);
      } else {
        s = 0;
#pragma unroll
        for (int j = 1; j < kMaxSeg; ++j) s += (j < p.nseg && u >= tile0[j]) ? 1 : 0;
        t = u - sel(tile0, s);
        rs = sel(rows, s);
        start = sel(segpre, s) + t * rs;
      }
      const int lo = max(f, start), hi = min(f1, start + rs);
      int take = (lane < kMaxPieces && u < U) ? max(hi - lo, 0) : 0;
      const bool real = sel(snnz, s) > 0;
      int real_rows = real ? take : 0;
      int incl = real_rows;
#pragma unroll
      for (int d = 1; d < kMaxPieces; d <<= 1) {
        int o = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += o;
      }
      const int excl = incl - real_rows;
      if (excl >= p.list_cap) { take = 0; real_rows = 0; }
      else if (excl + real_rows > p.list_cap) { real_rows = p.list_cap - excl; take = real_rows; }
      mine.seg = s; mine.tile = t; mine.take = take; mine.real_rows = real_rows;
      mine.list_off = excl; mine.a = lo - start;
      mine.cidx = pair ? t : u;
      mine.target = pair ? pair_rows : rs;
      np = __popc(__ballot_sync(0xffffffffu, take > 0));
      fnext = f + __reduce_add_sync(0xffffffffu, take);
    }
    if (first_stamp) TEAL_STAMP(3);

    // ---- 3b. windowed compaction: list[off + rank - a] = (m, x[m]) for rank in [a, a+rows) ----
    for (int pi = 0; pi < np; ++pi) {
      const Piece pc = piece_bcast(mine, pi);
      if (pc.real_rows == 0) continue;
      const int k = (NT == 1) ? 0 : p.seg[pc.seg].thr_slot;
      const int a = pc.a, b = pc.a + pc.real_rows;
      const int my_pre = sel(pre, k), my_cnt = sel(cnt, k);
      if (my_pre + my_cnt <= a || my_pre >= b) continue;
      const float th = sel(thr, k);
      int rank = my_pre;
      for (int m = c0; m < c1 && rank < b; ++m) {
        float xv = __half2float(reinterpret_cast<const __half*>(xs_vec)[m]);
        if (!(fabsf(xv) <= th)) {
          if (rank >= a) {
            list_idx[pc.list_off + rank - a] = m;
            list_x[pc.list_off + rank - a] = xv;
          }
          ++rank;
        }
      }
    }
    __syncthreads();
    if (first_stamp) TEAL_STAMP(4);

    if (pw >= 0) {
      // ================= producer warps: TMA bulk gather of surviving row segments ============
      // Stage `seq` is owned by producer warp seq % kProducerWarps.  Lanes compute their row's
      // source address in parallel; one elected lane issues the copies (UBLKCP takes uniform
      // operands, so a lane-divergent issue would be serialised by a waterfall loop anyway).
      for (int pi = 0; pi < np; ++pi) {
        const Piece pc = piece_bcast(mine, pi);
        if (pc.real_rows == 0) continue;
        const GemvSeg& sg = p.seg[pc.seg];
        const int wcols = min(TILE_N, sg.ncols - pc.tile * TILE_N);
        const uint32_t row_bytes = (uint32_t)wcols * 2u;
        const __half* wbase = sg.w + (size_t)pc.tile * TILE_N;
        for (int r0 = 0; r0 < pc.real_rows; r0 += R) {
          if (seq % kProducerWarps == pw) {
            const int nr = min(R, pc.real_rows - r0);
            const int m = (lane < nr) ? list_idx[pc.list_off + r0 + lane] : 0;
            const unsigned long long my_src =
                reinterpret_cast<unsigned long long>(wbase + (size_t)m * (size_t)sg.ldw);
            const uint32_t bar = full_u32 + stage * 8;
            const uint32_t dst0 = ring_u32 + (uint32_t)stage * STAGE_BYTES;
            mbar_wait(empty_u32 + stage * 8, phase ^ 1u);
            const bool leader = elect_one();
            if (leader) mbar_arrive_expect_tx(bar, (uint32_t)nr * row_bytes);
#pragma unroll 4
            for (int r = 0; r < nr; ++r) {
              const unsigned long long src = __shfl_sync(0xffffffffu, my_src, r);
              if (leader) bulk_g2s(dst0 + (uint32_t)r * ROW_BYTES, reinterpret_cast<const void*>(src), row_bytes, bar, policy);
            }
            if (dbg && pw == 0 && lane == 0 && seq == 0) dbg[11] = clock64();
          }
          ++seq;
          if (++stage == p.nstages) { stage = 0; phase ^= 1u; }
        }
      }
      if (dbg && pw == 0 && lane == 0) dbg[12] = clock64();
    } else {
      // ================= consumer warps =======================================================
      for (int pi = 0; pi < np; ++pi) {
        const Piece pc = piece_bcast(mine, pi);
        const GemvSeg& sg = p.seg[pc.seg];
        const int wcols = min(TILE_N, sg.ncols - pc.tile * TILE_N);
        const bool active = tcol < wcols;
        float acc[CPT];
#pragma unroll
        for (int c = 0; c < CPT; ++c) acc[c] = 0.f;

        for (int r0 = 0; r0 < pc.real_rows; r0 += R) {
          const int nr = min(R, pc.real_rows - r0);
          mbar_wait(full_u32 + stage * 8, phase);
          if (first_stamp && r0 == 0 && pi == 0) TEAL_STAMP(5);
          if (active) {
            const unsigned char* sb = ring + (size_t)stage * STAGE_BYTES + tcol * 2;
            const float* lx = list_x + pc.list_off + r0;
#pragma unroll
            for (int i = 0; i < RPG; ++i) {
              const int r = group + i * GROUPS;
              if (r < nr) {
                const float xv = lx[r];
                if constexpr (CPT == 4) {
                  uint2 v = *reinterpret_cast<const uint2*>(sb + (size_t)r * ROW_BYTES);
                  float2 w0 = __half22float2(*reinterpret_cast<const __half2*>(&v.x));
                  float2 w1 = __half22float2(*reinterpret_cast<const __half2*>(&v.y));
                  acc[0] = fmaf(w0.x, xv, acc[0]); acc[1] = fmaf(w0.y, xv, acc[1]);
                  acc[2] = fmaf(w1.x, xv, acc[2]); acc[3] = fmaf(w1.y, xv, acc[3]);
                } else {
                  uint4 v = *reinterpret_cast<const uint4*>(sb + (size_t)r * ROW_BYTES);
                  float2 w0 = __half22float2(*reinterpret_cast<const __half2*>(&v.x));
                  float2 w1 = __half22float2(*reinterpret_cast<const __half2*>(&v.y));
                  float2 w2 = __half22float2(*reinterpret_cast<const __half2*>(&v.z));
                  float2 w3 = __half22float2(*reinterpret_cast<const __half2*>(&v.w));
                  acc[0] = fmaf(w0.x, xv, acc[0]); acc[1] = fmaf(w0.y, xv, acc[1]);
                  acc[2] = fmaf(w1.x, xv, acc[2]); acc[3] = fmaf(w1.y, xv, acc[3]);
                  acc[4] = fmaf(w2.x, xv, acc[4]); acc[5] = fmaf(w2.y, xv, acc[5]);
                  acc[6] = fmaf(w3.x, xv, acc[6]); acc[7] = fmaf(w3.y, xv, acc[7]);
                }
              }
            }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(empty_u32 + stage * 8);
          if (++stage == p.nstages) { stage = 0; phase ^= 1u; }
        }
        TEAL_STAMP(6);

        // ---- combine: groups -> CTA partial -> fp32 RED into the L2 accumulator --------------
        const int colbase = sg.out_col0 + pc.tile * TILE_N;  // column in acc / y space
        if (pc.real_rows > 0) {
          if constexpr (GROUPS > 1) {
            if (active) {
#pragma unroll
              for (int c = 0; c < CPT; c += 4)
                *reinterpret_cast<float4*>(&red[group * TILE_N + tcol + c]) =
                    make_float4(acc[c], acc[c + 1], acc[c + 2], acc[c + 3]);
            }
            consumer_bar_sync();
            for (int c = tid * 4; c < wcols; c += kConsumerThreads * 4) {
              float4 s = *reinterpret_cast<const float4*>(&red[c]);
#pragma unroll
              for (int g = 1; g < GROUPS; ++g) {
                float4 t = *reinterpret_cast<const float4*>(&red[g * TILE_N + c]);
                s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
              }
              red_add_v4(p.acc + colbase + c, s);
            }
          } else {
            if (active) {
#pragma unroll
              for (int c = 0; c < CPT; c += 4)
                red_add_v4(p.acc + colbase + tcol + c,
                           make_float4(acc[c], acc[c + 1], acc[c + 2], acc[c + 3]));
            }
          }
        }
        // ---- completion counter: who finishes the tile converts it ---------------------------
        // REDs of all consumer threads happen-before the barrier; thread 0's release fence is
        // cumulative over them, its acquire fence orders the finaliser's loads after the counter.
        consumer_bar_sync();
        TEAL_STAMP(7);
        if (tid == 0) {
          fence_acq_rel_gpu();
          unsigned old = atomicAdd(&p.counters[pc.cidx], (unsigned)pc.take);
          const int last = ((int)old + pc.take == pc.target) ? 1 : 0;
          if (last) fence_acq_rel_gpu();
          s_last = last;
        }
        consumer_bar_sync();
        TEAL_STAMP(8);
        if (s_last) {
          if (pair) {
            const int I = p.seg[0].ncols;
            const int cb = pc.tile * TILE_N;
            for (int c = tid * 4; c < wcols; c += kConsumerThreads * 4) {
              float4 g = ld_cg_f4(p.acc + cb + c);
              float4 u = ld_cg_f4(p.acc + I + cb + c);
              st_cg_f4(p.acc + cb + c, make_float4(0.f, 0.f, 0.f, 0.f));
              st_cg_f4(p.acc + I + cb + c, make_float4(0.f, 0.f, 0.f, 0.f));
              __half2 h0 = __halves2half2(silu_mul_f16(g.x, u.x), silu_mul_f16(g.y, u.y));
              __half2 h1 = __halves2half2(silu_mul_f16(g.z, u.z), silu_mul_f16(g.w, u.w));
              uint2 o;
              o.x = *reinterpret_cast<uint32_t*>(&h0);
              o.y = *reinterpret_cast<uint32_t*>(&h1);
              *reinterpret_cast<uint2*>(p.y + cb + c) = o;
            }
          } else {
            for (int c = tid * 4; c < wcols; c += kConsumerThreads * 4) {
              float4 v = ld_cg_f4(p.acc + colbase + c);
              st_cg_f4(p.acc + colbase + c, make_float4(0.f, 0.f, 0.f, 0.f));
              __half2 h0 = __floats2half2_rn(v.x, v.y);
              __half2 h1 = __floats2half2_rn(v.z, v.w);
              uint2 o;
              o.x = *reinterpret_cast<uint32_t*>(&h0);
              o.y = *reinterpret_cast<uint32_t*>(&h1);
              *reinterpret_cast<uint2*>(p.y + colbase + c) = o;
            }
          }
          if (tid == 0) p.counters[pc.cidx] = 0;
        }
        TEAL_STAMP(9);
      }
    }
    if (dbg && tid == 0 && first_stamp) dbg[15] = (long long)np | ((long long)(f1 - f0) << 8);
    first_stamp = false;
    f = fnext;
    __syncthreads();  // lists are rewritten next round
  }
  griddep_launch_dependents();
  TEAL_STAMP(10);
  if (dbg && tid == 0) { dbg[14] = globaltimer_ns(); }
#undef TEAL_STAMP
}

// ------------------------------------------------------------------------------------------
// Slow independent cross-check kernel: one warp per 256 output columns, walks every lane of x.
// ------------------------------------------------------------------------------------------
__global__ void sparse_gemv_simple_kernel(const __half* __restrict__ x, const __half* __restrict__ wt,
                                          __half* __restrict__ y, int Z, int N, long long ldw,
                                          float thr) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int col = warp * 256 + lane * 8;
  if (col >= N) return;
  float acc[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) acc[c] = 0.f;
  for (int m = 0; m < Z; ++m) {
    float xv = __half2float(x[m]);
    if (!(fabsf(xv) <= thr)) {
      uint4 v = *reinterpret_cast<const uint4*>(wt + (size_t)m * ldw + col);
      const __half2* h2 = reinterpret_cast<const __half2*>(&v);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float2 f = __half22float2(h2[j]);
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
struct TileConfig {
  int cpt, groups, rows;
  void (*kernel[2])(const GemvParams);  // [0]: one threshold, [1]: up to kMaxThr thresholds
};
#define TEAL_CFG(C, G, RR) {C, G, RR, {sparse_gemv_kernel<C, G, RR, 1>, sparse_gemv_kernel<C, G, RR, kMaxThr>}}
static const TileConfig kConfigs[] = {
    TEAL_CFG(4, 4, 16),  // 1:  512-col tiles (1 KB row segments), 16 KB stages
    TEAL_CFG(8, 4, 8),   // 2: 1024-col tiles (2 KB), 16 KB stages
    TEAL_CFG(8, 2, 8),   // 3: 2048-col tiles (4 KB), 32 KB stages
    TEAL_CFG(8, 2, 4),   // 4: 2048-col tiles (4 KB), 16 KB stages
    TEAL_CFG(8, 1, 4),   // 5: 4096-col tiles (8 KB), 32 KB stages, no in-CTA reduction
    TEAL_CFG(8, 4, 16),  // 6: 1024-col tiles (2 KB), 32 KB stages
};
#undef TEAL_CFG
static constexpr int kNumConfigs = sizeof(kConfigs) / sizeof(kConfigs[0]);
static_assert(kNumConfigs == kNumTmaConfigs, "common.cuh: kNumTmaConfigs");

static int pick_config(const GemvParams&) {
  const int id = forced_config();
  return (id >= 1 && id <= kNumConfigs) ? id - 1 : 0;
}

int launch_gemv(GemvParams& p, void* workspace, size_t ws_bytes, int out_cols_total, cudaStream_t stream) {
  const DeviceInfo& di = device_info();
  if (di.err != cudaSuccess) return set_cuda_error(di.err, "device query");
  if (p.Z <= 0 || p.Z > 65536 || (p.Z & 7)) return set_error(TEAL_ERR_ALIGNMENT, "Z must be in (0, 65536] and a multiple of 8");
  if ((reinterpret_cast<uintptr_t>(p.x) & 15)) return set_error(TEAL_ERR_ALIGNMENT, "x must be 16-byte aligned");
  if (p.mode != kModeSurvivors) {
    if (!workspace || ws_bytes < workspace_bytes(out_cols_total)) return set_error(TEAL_ERR_WORKSPACE, "workspace too small");
    if ((reinterpret_cast<uintptr_t>(p.y) & 15)) return set_error(TEAL_ERR_ALIGNMENT, "y must be 16-byte aligned");
    p.counters = reinterpret_cast<unsigned*>(workspace);
    p.acc = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(workspace) + (size_t)kMaxCounters * 4);
  }
  p.dbg_time = debug_timestamps();
  const int ci = pick_config(p);
  const TileConfig& cfg = kConfigs[ci];
  const int tpg = kConsumerThreads / cfg.groups;
  const int tile_n = tpg * cfg.cpt;
  int tile0 = 0;
  for (int s = 0; s < p.nseg; ++s) {
    GemvSeg& sg = p.seg[s];
    if (sg.ncols <= 0 || (sg.ncols & 7) || (sg.ldw & 7) || (sg.out_col0 & 7) ||
        (reinterpret_cast<uintptr_t>(sg.w) & 15))
      return set_error(TEAL_ERR_ALIGNMENT, "weights: N, ldw, region offsets must be multiples of 8 and 16-byte aligned");
    sg.ntiles = (sg.ncols + tile_n - 1) / tile_n;
    sg.tile0 = tile0;
    tile0 += sg.ntiles;
  }
  if (tile0 > kMaxCounters) return set_error(TEAL_ERR_UNSUPPORTED, "too many column tiles");
  p.chunk = (((p.Z + kConsumerThreads - 1) / kConsumerThreads) + 7) & ~7;
  p.xs_elems = (p.Z + 7) & ~7;
  p.list_cap = 2048;
  const size_t stage_bytes = (size_t)cfg.rows * tile_n * 2;
  const size_t fixed = (size_t)kMaxStages * 16 + (cfg.groups > 1 ? (size_t)cfg.groups * tile_n * 4 : 0) +
                       (size_t)p.list_cap * 8 + (size_t)p.xs_elems * 2 + 128;
  const size_t budget = (size_t)di.smem_optin - 2048;  // static __shared__ + slack
  if (fixed + 2 * stage_bytes > budget) return set_error(TEAL_ERR_UNSUPPORTED, "Z too large for shared memory staging");
  int nstages = (int)((budget - fixed) / stage_bytes);
  if (nstages > kMaxStages) nstages = kMaxStages;
  p.nstages = nstages;
  const size_t smem = fixed + (size_t)nstages * stage_bytes;

  const int ki = (p.nthr > 1) ? 1 : 0;
  static size_t s_attr_set[kNumConfigs][2] = {{0}};
  if (s_attr_set[ci][ki] < smem) {
    cudaError_t e = cudaFuncSetAttribute(cfg.kernel[ki], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget);
    if (e != cudaSuccess) return set_cuda_error(e, "cudaFuncSetAttribute");
    s_attr_set[ci][ki] = budget;
  }
  const int grid = (p.mode == kModeSurvivors) ? 1 : di.sm_count;
  cfg.kernel[ki]<<<grid, kThreads, smem, stream>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error(e, "sparse_gemv launch");
  return TEAL_OK;
}

int launch_simple(const __half* x, const __half* wt, __half* y, int Z, int N, long long ldw, float thr,
                  cudaStream_t stream) {
  if (N & 7) return set_error(TEAL_ERR_ALIGNMENT, "N must be a multiple of 8");
  const int warps = (N + 255) / 256;
  const int threads = 128;
  const int blocks = (warps * 32 + threads - 1) / threads;
  sparse_gemv_simple_kernel<<<blocks, threads, 0, stream>>>(x, wt, y, Z, N, ldw, thr);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error(e, "simple launch");
  return TEAL_OK;
}

}  // namespace teal
