// The rest of one batch-1 decode step around the sparse GEMVs:
// token embedding, RoPE + KV-append + decode attention, dense lm_head GEMV with the final RMSNorm
// fused in and a fused arg-max, and the step bookkeeping.  Reference: gpt-fast/model.py:133-190,
// 286-318 and generate.py:49-77.  All kernels read the position / token from device memory so
// that one captured CUDA graph can be replayed for every token.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "common.cuh"

namespace teal {

// --------------------------------------------------------------------------------------------
// x[dim] = tok_embeddings[token]            (gpt-fast/model.py:137)
// --------------------------------------------------------------------------------------------
// Also leaves the row's sum of squares in ssq_out[0] (and zeroes ssq_out[1..kSsqParts)) for the first RMSNorm.
__global__ void embed_kernel(const __half* __restrict__ emb, const int* __restrict__ token, __half* __restrict__ x,
                             float* __restrict__ ssq_out, int dim) {
  __shared__ float sp[32];
  griddep_wait();
  const int t = token[0];
  const uint4* src = reinterpret_cast<const uint4*>(emb + (size_t)t * dim);
  uint4* dst = reinterpret_cast<uint4*>(x);
  float ss = 0.f;
  for (int i = threadIdx.x; i < dim / 8; i += blockDim.x) {
    const uint4 v = src[i];
    dst[i] = v;
    const __half2* h2 = reinterpret_cast<const __half2*>(&v);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = __half22float2(h2[j]);
      ss = fmaf(f.x, f.x, fmaf(f.y, f.y, ss));
    }
  }
  if (ssq_out) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, d);
    if ((threadIdx.x & 31) == 0) sp[threadIdx.x >> 5] = ss;
    __syncthreads();
    if (threadIdx.x == 0) {
      float a = 0.f;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) a += sp[w];
      sp[0] = a;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kSsqParts * kSsqReplicas; i += blockDim.x)  // every replica: {total, 0, 0, ...}
      ssq_out[i] = (i % kSsqParts == 0) ? sp[0] : 0.f;
  }
  griddep_launch_dependents();
}

// --------------------------------------------------------------------------------------------
// RoPE(q, k) + KV-cache append + attention of one query token over positions [0, pos].
// gpt-fast/model.py:170-186: adjacent-pair rotation with the fp16 (cos, sin) table, KVCache.update,
// GQA (kv head = q head / rep), scaled dot product attention, heads concatenated.
// grid = n_head CTAs (one per query head), block = kAttnThreads.
//
// The kernel sits between two HBM-bound GEMVs and is pure latency at the benchmark's short context, so it is shaped
// for the length of its dependency chain: 16 warps, each with 8 positions' K and V rows in flight at once (128 positions
// per sweep; +2 % tokens/s over 8 warps x 4 positions); the position, its RoPE row and the first sweep's cache rows are
// requested BEFORE griddepcontrol.wait (they do not depend on the QKV launch), only q | k | v after it.
// Measured and dropped (profiles/r02_quick_ab_attn_pf*.jsonl): helper CTAs that pull the o projection's weights into L2
// while this kernel leaves HBM idle -- 33 MB do not fit the 2.5 us window, the rest delays the o projection's own
// requests (-3 %).
// --------------------------------------------------------------------------------------------
constexpr int kAttnThreads = 512;
constexpr int kAttnWarps = kAttnThreads / 32;

template <int HD>
__global__ void __launch_bounds__(kAttnThreads) attn_decode_kernel(
    const __half* __restrict__ qkv,    // [n_head*HD | n_kv*HD | n_kv*HD]
    const float* __restrict__ qkv_acc, // or (qkv == nullptr): the QKV launch's deferred fp32 sums, rounded to fp16 here
    const __half* __restrict__ freqs,  // [max_seq][HD/2][2] fp16 (cos, sin)
    __half* __restrict__ k_cache,      // [n_kv][max_seq][HD]
    __half* __restrict__ v_cache, const int* __restrict__ pos_ptr, __half* __restrict__ out,  // [n_head*HD]
    int n_head, int n_kv, int max_seq, float scale) {
  static_assert(HD % 64 == 0 && HD <= 256, "head_dim");
  constexpr int EPL = HD / 32;  // elements per lane (4 for HD = 128)
  __shared__ float s_q[HD];
  __shared__ float s_knew[HD];
  __shared__ float s_vnew[HD];
  __shared__ float s_m[kAttnWarps], s_l[kAttnWarps];
  __shared__ float s_acc[kAttnWarps][HD];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int h = blockIdx.x;
  const int rep = n_head / n_kv;
  const int kvh = h / rep;
  // ---- before the producer (the QKV launch) is done: everything that does not depend on it.  The position was written
  // by the previous step's `advance`, the cache rows [0, pos) by earlier steps, the RoPE table is constant -- all complete
  // before the producer itself was allowed to start.  So the position, its (cos, sin) row and the first sweep's K / V rows
  // (8 positions per warp, 128 per CTA: the whole context at the benchmark's length) are in registers when the wait
  // returns, and the dependent chain after it is ONE round trip (q | k | v) instead of three. -----------------------
  constexpr int PB = 8;  // positions per warp per sweep: 2 * PB independent row requests per lane
  const int pos = pos_ptr[0];
  const bool pos_ok = pos >= 0 && pos < max_seq;
  const __half* kbase = k_cache + (size_t)kvh * max_seq * HD;
  const __half* vbase = v_cache + (size_t)kvh * max_seq * HD;
  __half2 kraw[PB][EPL / 2], vraw[PB][EPL / 2];
  auto load_rows = [&](int p0) {
#pragma unroll
    for (int b = 0; b < PB; ++b) {
      const int p = p0 + b * kAttnWarps;
      if (p < pos) {
        const __half2* kp = reinterpret_cast<const __half2*>(kbase + (size_t)p * HD + lane * EPL);
        const __half2* vp = reinterpret_cast<const __half2*>(vbase + (size_t)p * HD + lane * EPL);
#pragma unroll
        for (int e = 0; e < EPL / 2; ++e) { kraw[b][e] = kp[e]; vraw[b][e] = vp[e]; }
      }
    }
  };
  __half2 cs = __halves2half2(__ushort_as_half((unsigned short)0), __ushort_as_half((unsigned short)0));
  if (pos_ok) {
    if (tid < HD / 2) cs = reinterpret_cast<const __half2*>(freqs + (size_t)pos * HD)[tid];  // (cos, sin) of pair `tid`
    load_rows(warp);
  }
  griddep_wait();
  griddep_launch_dependents();  // n_head CTAs leave most of the machine free: the o-proj grid moves in and waits for us
  const size_t qo = (size_t)h * HD, ko = (size_t)n_head * HD + (size_t)kvh * HD, vo = (size_t)(n_head + n_kv) * HD + (size_t)kvh * HD;
  __half2 in2 = __halves2half2(__ushort_as_half((unsigned short)0), __ushort_as_half((unsigned short)0));
  __half2 k2 = in2;
  if (qkv) {
    if (tid < HD / 2) {
      in2 = reinterpret_cast<const __half2*>(qkv + qo)[tid];
      k2 = reinterpret_cast<const __half2*>(qkv + ko)[tid];
    } else if (tid < HD) {
      in2 = reinterpret_cast<const __half2*>(qkv + vo)[tid - HD / 2];
    }
  } else {  // the same fp16 values the QKV tile finaliser would have written (one round-to-nearest of the fp32 sum)
    if (tid < HD / 2) {
      const float2 a = __ldcg(reinterpret_cast<const float2*>(qkv_acc + qo) + tid);
      const float2 b = __ldcg(reinterpret_cast<const float2*>(qkv_acc + ko) + tid);
      in2 = __floats2half2_rn(a.x, a.y);
      k2 = __floats2half2_rn(b.x, b.y);
    } else if (tid < HD) {
      const float2 a = __ldcg(reinterpret_cast<const float2*>(qkv_acc + vo) + (tid - HD / 2));
      in2 = __floats2half2_rn(a.x, a.y);
    }
  }
  if (!pos_ok) {  // beyond the KV cache / RoPE table (the host raises before this can happen): touch nothing
    if (tid < HD) out[(size_t)h * HD + tid] = __ushort_as_half((unsigned short)0x7e00);  // NaN: a silent wrong token is worse
    return;
  }
  // ---- RoPE on q (this head) and on the new k (its kv head); v passes through ---------------------
  if (tid < HD / 2) {
    const float c = __low2float(cs), sn = __high2float(cs);
    const float q0 = __low2float(in2), q1 = __high2float(in2);
    const float k0 = __low2float(k2), k1 = __high2float(k2);
    // model.py:309-315, result cast back to fp16 (type_as)
    const __half qa = __float2half_rn(q0 * c - q1 * sn), qb = __float2half_rn(q1 * c + q0 * sn);
    const __half ka = __float2half_rn(k0 * c - k1 * sn), kb = __float2half_rn(k1 * c + k0 * sn);
    s_q[2 * tid] = __half2float(qa);
    s_q[2 * tid + 1] = __half2float(qb);
    s_knew[2 * tid] = __half2float(ka);
    s_knew[2 * tid + 1] = __half2float(kb);
    if (h % rep == 0)  // one CTA per kv head appends to the cache (model.py:93-94)
      reinterpret_cast<__half2*>(k_cache + ((size_t)kvh * max_seq + pos) * HD)[tid] = __halves2half2(ka, kb);
  } else if (tid < HD) {
    const int i = tid - HD / 2;
    s_vnew[2 * i] = __low2float(in2);
    s_vnew[2 * i + 1] = __high2float(in2);
    if (h % rep == 0) reinterpret_cast<__half2*>(v_cache + ((size_t)kvh * max_seq + pos) * HD)[i] = in2;
  }
  __syncthreads();

  // ---- online-softmax attention: warp w takes positions w, w + kAttnWarps, ... ---------------------
  float qreg[EPL];
#pragma unroll
  for (int e = 0; e < EPL; ++e) qreg[e] = s_q[lane * EPL + e];
  float m = -CUDART_INF_F, l = 0.f, acc[EPL];
#pragma unroll
  for (int e = 0; e < EPL; ++e) acc[e] = 0.f;
  for (int p0 = warp; p0 <= pos; p0 += kAttnWarps * PB) {
    if (p0 != warp) load_rows(p0);  // (the first sweep's rows were requested before the wait)
    float kk[PB][EPL], vv[PB][EPL];
#pragma unroll
    for (int b = 0; b < PB; ++b) {
      const int p = p0 + b * kAttnWarps;
      if (p < pos) {
#pragma unroll
        for (int e = 0; e < EPL / 2; ++e) {
          const float2 kf = __half22float2(kraw[b][e]), vf = __half22float2(vraw[b][e]);
          kk[b][2 * e] = kf.x; kk[b][2 * e + 1] = kf.y;
          vv[b][2 * e] = vf.x; vv[b][2 * e + 1] = vf.y;
        }
      } else {  // the new position's row comes from shared memory (every CTA of the kv head has it)
#pragma unroll
        for (int e = 0; e < EPL; ++e) { kk[b][e] = s_knew[lane * EPL + e]; vv[b][e] = s_vnew[lane * EPL + e]; }
      }
    }
#pragma unroll
    for (int b = 0; b < PB; ++b) {
      const int p = p0 + b * kAttnWarps;
      if (p <= pos) {
        float sc = 0.f;
#pragma unroll
        for (int e = 0; e < EPL; ++e) sc = fmaf(qreg[e], kk[b][e], sc);
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) sc += __shfl_xor_sync(0xffffffffu, sc, d);
        sc *= scale;
        const float mn = fmaxf(m, sc);
        const float corr = __expf(m - mn), pe = __expf(sc - mn);
        l = l * corr + pe;
#pragma unroll
        for (int e = 0; e < EPL; ++e) acc[e] = fmaf(pe, vv[b][e], acc[e] * corr);
        m = mn;
      }
    }
  }
  if (lane == 0) { s_m[warp] = m; s_l[warp] = l; }
#pragma unroll
  for (int e = 0; e < EPL; ++e) s_acc[warp][lane * EPL + e] = acc[e];
  __syncthreads();
  // ---- merge the warps' partial softmaxes ---------------------------------------------------------
  if (tid < HD) {
    float M = -CUDART_INF_F;
#pragma unroll
    for (int w = 0; w < kAttnWarps; ++w) M = fmaxf(M, s_m[w]);
    float L = 0.f, o = 0.f;
#pragma unroll
    for (int w = 0; w < kAttnWarps; ++w) {
      const float f = (s_m[w] == -CUDART_INF_F) ? 0.f : __expf(s_m[w] - M);
      L = fmaf(s_l[w], f, L);
      o = fmaf(s_acc[w][tid], f, o);
    }
    out[(size_t)h * HD + tid] = __float2half_rn(o / L);
  }
}

// --------------------------------------------------------------------------------------------
// Dense row-major GEMV for the (never sparsified) lm_head, gpt-fast/model.py:141-142:
//   logits[v] = sum_z W[v][z] * rmsnorm(x)[z]
// with the final RMSNorm in the prologue and a fused arg-max (packed 64-bit atomicMax).
// One warp per output row, 16 B per lane per load, warp-shuffle reduction.
// --------------------------------------------------------------------------------------------
constexpr int kHeadThreads = 512;

__device__ __forceinline__ unsigned long long pack_argmax(float v, int idx) {
  unsigned b = __float_as_uint(v);
  b = (b & 0x80000000u) ? ~b : (b | 0x80000000u);  // monotonic map float -> unsigned
  return ((unsigned long long)b << 32) | (unsigned long long)(0xffffffffu - (unsigned)idx);  // ties -> smallest index
}

template <int UNROLL>
__global__ void __launch_bounds__(kHeadThreads, 2) dense_rowmajor_gemv_kernel(
    const __half* __restrict__ x, const __half* __restrict__ norm_w, float eps, const __half* __restrict__ W,
    __half* __restrict__ y, unsigned long long* __restrict__ argmax_packed, int Z, int V, long long ldw) {
#ifdef TEAL_EMU
  unsigned char* smem_raw = ::emu::dyn_smem();
#else
  extern __shared__ __align__(16) unsigned char smem_raw[];
#endif
  __half* xs = reinterpret_cast<__half*>(smem_raw);  // normalised x, Z halves
  __shared__ float s_part[kHeadThreads / 32];
  __shared__ unsigned long long s_best[kHeadThreads / 32];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int kWarps = kHeadThreads / 32;
  griddep_wait();
  // ---- prologue: RMSNorm (model.py:286-291) or plain copy ------------------------------------------
  float ss = 0.f;
  for (int m = tid * 8; m < Z; m += kHeadThreads * 8) {
    uint4 v = ld_nc_u4(reinterpret_cast<const uint4*>(x + m));
    *reinterpret_cast<uint4*>(xs + m) = v;
    const __half2* h2 = reinterpret_cast<const __half2*>(&v);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float2 f = __half22float2(h2[j]);
      ss = fmaf(f.x, f.x, ss);
      ss = fmaf(f.y, f.y, ss);
    }
  }
  if (norm_w) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, d);
    if (lane == 0) s_part[warp] = ss;
    __syncthreads();
    float tot = (lane < kWarps) ? s_part[lane] : 0.f;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, d);
    const float rs = rsqrtf(tot / (float)Z + eps);
    for (int m = tid * 8; m < Z; m += kHeadThreads * 8) {
      uint4 v = *reinterpret_cast<uint4*>(xs + m);
      uint4 g = ld_nc_u4(reinterpret_cast<const uint4*>(norm_w + m));
      __half2* h2 = reinterpret_cast<__half2*>(&v);
      const __half2* g2 = reinterpret_cast<const __half2*>(&g);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float2 f = __half22float2(h2[j]);
        h2[j] = __hmul2(__floats2half2_rn(f.x * rs, f.y * rs), g2[j]);
      }
      *reinterpret_cast<uint4*>(xs + m) = v;
    }
  }
  __syncthreads();

  // ---- rows: warp gw takes rows gw, gw + total_warps, ... ------------------------------------------
  const int total_warps = gridDim.x * kWarps;
  unsigned long long best = 0ull;
  const int nvec = Z >> 3;  // 16-byte vectors per row
  for (int v = blockIdx.x * kWarps + warp; v < V; v += total_warps) {
    const uint4* wr = reinterpret_cast<const uint4*>(W + (size_t)v * ldw);
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
      y[v] = r;
      const unsigned long long pk = pack_argmax(__half2float(r), v);
      best = pk > best ? pk : best;
    }
  }
  if (argmax_packed) {
    if (lane == 0) s_best[warp] = best;
    __syncthreads();
    if (tid == 0) {
      unsigned long long b = 0ull;
#pragma unroll
      for (int w = 0; w < kWarps; ++w) b = s_best[w] > b ? s_best[w] : b;
      if (b) atomicMax(argmax_packed, b);
    }
  }
  griddep_launch_dependents();
}

// --------------------------------------------------------------------------------------------
// Step bookkeeping: next token = arg-max (greedy; reference sample() with temperature -> 0,
// generate.py:49-66), append to the output, advance the position, reset the arg-max cell.
// --------------------------------------------------------------------------------------------
__global__ void advance_kernel(unsigned long long* argmax_packed, int* token, int* pos, int* out_tokens, int out_cap) {
  griddep_wait();
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    const unsigned long long pk = *argmax_packed;
    const int t = (int)(0xffffffffu - (unsigned)(pk & 0xffffffffull));
    const int p = pos[0];
    token[0] = t;
    if (p < out_cap) out_tokens[p] = t;
    pos[0] = p + 1;
    *argmax_packed = 0ull;
  }
  griddep_launch_dependents();
}

// --------------------------------------------------------------------------------------------
// Top-k / temperature sampling of the next token (gpt-fast/generate.py:49-66, `sample`):
//   l = fp16(logits / max(T, 1e-5));  pivot = k-th largest l;  l < pivot -> -inf;  p = fp16(softmax_fp32(l));
//   q ~ Exponential(1) in fp16;  token = argmax fp16(p / q)   (first index on ties, like torch.argmax)
// One CTA, four passes over the L2-resident fp16 logits: radix select of the pivot on the order-preserving 16-bit key
// (high byte, then low byte), online soft-max statistics of the kept values, arg-max of p / q.  The result replaces
// the greedy arg-max in the packed cell that advance_kernel consumes.  q comes from `q_ext` (tests: the reference's
// own noise) or from Philox4x32-10 keyed by (seed; vocabulary index, position) -- the reference's torch generator
// stream cannot be reproduced, the distribution is the same.
// --------------------------------------------------------------------------------------------
constexpr int kSampleMaxThreads = 1024;

__host__ __device__ inline unsigned sample_key(unsigned short b) {  // fp16 bits -> unsigned key with the same order
  if (b == 0x8000u) b = 0;  // -0.0 == +0.0 for torch.topk and for `l < pivot`
  return (b & 0x8000u) ? (unsigned)(unsigned short)~b : (unsigned)(b | 0x8000u);
}
__host__ __device__ inline void philox4x32_10(unsigned c0, unsigned c1, unsigned c2, unsigned c3, unsigned k0, unsigned k1, unsigned out[4]) {
  for (int r = 0; r < 10; ++r) {
    const unsigned long long p0 = 0xD2511F53ull * c0, p1 = 0xCD9E8D57ull * c2;
    const unsigned n0 = (unsigned)(p1 >> 32) ^ c1 ^ k0, n1 = (unsigned)p1, n2 = (unsigned)(p0 >> 32) ^ c3 ^ k1, n3 = (unsigned)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
// Exponential(1) noise of vocabulary entry `idx` at decode position `pos`: u in (0, 1], q = fp16(-log u), never zero
__device__ __forceinline__ __half sample_noise(unsigned long long seed, int idx, int pos) {
  unsigned r[4];
  philox4x32_10((unsigned)idx, (unsigned)pos, 0u, 0u, (unsigned)seed, (unsigned)(seed >> 32), r);
  const float u = ((float)(r[0] >> 8) + 1.0f) * (1.0f / 16777216.0f);
  const __half q = __float2half_rn(-logf(u));
  return __half_as_ushort(q) == 0 ? __ushort_as_half((unsigned short)1) : q;
}

__global__ void __launch_bounds__(kSampleMaxThreads) sample_topk_kernel(const __half* __restrict__ logits, int V, float temperature,
                                                                       int top_k, const __half* __restrict__ q_ext,
                                                                       unsigned long long seed, const int* __restrict__ pos,
                                                                       unsigned long long* __restrict__ argmax_packed,
                                                                       __half* __restrict__ probs_out) {
  __shared__ unsigned s_hist[256];
  __shared__ unsigned s_sel[2];   // selected bin, count still needed inside it
  __shared__ float s_m[32], s_s[32];
  __shared__ unsigned long long s_best[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
  griddep_wait();
  const float t = fmaxf(temperature, 1e-5f);
  auto scaled = [&](int i) { return __float2half_rn(__half2float(logits[i]) / t); };  // fp16 tensor / python float
  const bool use_topk = top_k > 0 && top_k < V;
  unsigned pivot_key = 0u;  // keep everything
  if (use_topk) {
    unsigned need = (unsigned)top_k, prefix = 0u;
    for (int pass = 0; pass < 2; ++pass) {  // high byte of the key, then the low byte among keys with that high byte
      for (int b = tid; b < 256; b += blockDim.x) s_hist[b] = 0u;
      __syncthreads();
      for (int i = tid; i < V; i += blockDim.x) {
        const unsigned key = sample_key(__half_as_ushort(scaled(i)));
        if (pass == 0) atomicAdd(&s_hist[key >> 8], 1u);
        else if ((key >> 8) == prefix) atomicAdd(&s_hist[key & 0xffu], 1u);
      }
      __syncthreads();
      if (tid == 0) {  // walk down from the largest bin until `need` values are covered
        unsigned above = 0u;
        int b = 255;
        for (; b > 0; --b) {
          if (above + s_hist[b] >= need) break;
          above += s_hist[b];
        }
        s_sel[0] = (unsigned)b;
        s_sel[1] = need - above;
      }
      __syncthreads();
      if (pass == 0) prefix = s_sel[0];
      else pivot_key = (prefix << 8) | s_sel[0];
      need = s_sel[1];
      __syncthreads();
    }
  }
  // ---- soft-max statistics of the kept values (online max / sum, fp32) --------------------------------------
  float m = -CUDART_INF_F, sum = 0.f;
  for (int i = tid; i < V; i += blockDim.x) {
    const __half l = scaled(i);
    if (sample_key(__half_as_ushort(l)) >= pivot_key) {
      const float f = __half2float(l);
      if (f > m) { sum = sum * expf(m - f) + 1.f; m = f; }
      else if (f != -CUDART_INF_F) sum += expf(f - m);
    }
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, m, d), s2 = __shfl_xor_sync(0xffffffffu, sum, d);
    const float mm = fmaxf(m, m2);
    sum = (m == -CUDART_INF_F ? 0.f : sum * expf(m - mm)) + (m2 == -CUDART_INF_F ? 0.f : s2 * expf(m2 - mm));
    m = mm;
  }
  if (lane == 0) { s_m[warp] = m; s_s[warp] = sum; }
  __syncthreads();
  m = -CUDART_INF_F;
  for (int w = 0; w < nwarp; ++w) m = fmaxf(m, s_m[w]);
  sum = 0.f;
  for (int w = 0; w < nwarp; ++w) sum += (s_m[w] == -CUDART_INF_F) ? 0.f : s_s[w] * expf(s_m[w] - m);
  // ---- arg-max of fp16(p / q) over the kept values ------------------------------------------------------------
  const int p0 = pos ? pos[0] : 0;
  unsigned long long best = 0ull;
  for (int i = tid; i < V; i += blockDim.x) {
    const __half l = scaled(i);
    const bool keep = sample_key(__half_as_ushort(l)) >= pivot_key;
    const __half p = keep ? __float2half_rn(expf(__half2float(l) - m) / sum) : __ushort_as_half((unsigned short)0);
    if (probs_out) probs_out[i] = p;
    if (keep) {
      const __half q = q_ext ? q_ext[i] : sample_noise(seed, i, p0);
      const __half r = __float2half_rn(__half2float(p) / __half2float(q));
      const unsigned long long pk = ((unsigned long long)sample_key(__half_as_ushort(r)) << 32) | (unsigned long long)(0xffffffffu - (unsigned)i);
      best = pk > best ? pk : best;
    }
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    const unsigned long long o = __shfl_xor_sync(0xffffffffu, best, d);
    best = o > best ? o : best;
  }
  if (lane == 0) s_best[warp] = best;
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < nwarp; ++w) best = s_best[w] > best ? s_best[w] : best;
    *argmax_packed = best;  // same packing as the greedy arg-max: low word = 0xffffffff - index
  }
  griddep_launch_dependents();
}

// --------------------------------------------------------------------------------------------
// TP path: x += all-reduced partial (fp16 add, gpt-fast/model.py:160-161 after tp.py:120-121,139-140) and the
// sum of squares of the new residual stream for the next RMSNorm, in one tiny launch.
// --------------------------------------------------------------------------------------------
__global__ void residual_add_kernel(__half* __restrict__ x, const __half* __restrict__ part, float* __restrict__ ssq_out,
                                    int dim) {
  __shared__ float sp[32];
  griddep_wait();
  float ss = 0.f;
  for (int i = threadIdx.x; i < dim / 2; i += blockDim.x) {
    const __half2 a = reinterpret_cast<const __half2*>(x)[i], b = reinterpret_cast<const __half2*>(part)[i];
    const __half2 r = __hadd2(a, b);
    reinterpret_cast<__half2*>(x)[i] = r;
    const float2 f = __half22float2(r);
    ss = fmaf(f.x, f.x, fmaf(f.y, f.y, ss));
  }
  if (ssq_out) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, d);
    if ((threadIdx.x & 31) == 0) sp[threadIdx.x >> 5] = ss;
    __syncthreads();
    if (threadIdx.x == 0) {
      float a = 0.f;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) a += sp[w];
      sp[0] = a;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kSsqParts * kSsqReplicas; i += blockDim.x)  // every replica: {total, 0, 0, ...}
      ssq_out[i] = (i % kSsqParts == 0) ? sp[0] : 0.f;
  }
  griddep_launch_dependents();
}

// --------------------------------------------------------------------------------------------
// Tensor-parallel exchange step, fused: one-shot all-reduce over NVLink peer memory + residual add + sum of
// squares, in ONE single-CTA kernel (replaces ncclAllReduce + add, gpt-fast/tp.py:120-121,139-140 + model.py:160-161).
//   * every rank's row-parallel GEMV writes its fp16 partial straight into its own symmetric buffer
//   * this kernel publishes "epoch e ready" into every peer's flag array (st.release.sys over NVLink), spins until all
//     peers have published e (ld.acquire.sys on local memory), then pulls the ws partials with volatile 16-byte P2P
//     loads, sums them in rank order in fp32 (identical on every rank => replicas stay bit-identical), rounds once to
//     fp16, adds the residual in fp16 and leaves the sum of squares for the next RMSNorm.
// The two all-reduces of a layer use alternating buffers (o-proj: 0, down-proj: 1): a rank can only overwrite a
// buffer two exchanges later, by which time every peer has entered the next exchange and is done reading it.
// --------------------------------------------------------------------------------------------
struct TpPeers {
  const __half* buf[kMaxTpWorld];  // this parity's partial buffer of every rank (peer-mapped)
  unsigned* flags[kMaxTpWorld];    // flag array of every rank: flags[r][s] = last epoch published by sender s
};

#ifdef TEAL_EMU  // CPU emulation build of tests/emu: C++ twins of the inline PTX below
__device__ __forceinline__ void st_release_sys_u32(unsigned* p, unsigned v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }
__device__ __forceinline__ unsigned ld_acquire_sys_u32(const unsigned* p) { ::emu::yield(); return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
__device__ __forceinline__ uint4 ld_volatile_u4(const void* p) { uint4 r; memcpy(&r, p, sizeof r); return r; }
#else
__device__ __forceinline__ void st_release_sys_u32(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint4 ld_volatile_u4(const void* p) {
  uint4 r;
  asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
  return r;
}
#endif

__global__ void __launch_bounds__(1024, 1) tp_allreduce_residual_kernel(const __grid_constant__ TpPeers peers, int rank,
                                                                        int ws, __half* __restrict__ x,
                                                                        float* __restrict__ ssq_out, unsigned* epoch_cell,
                                                                        int dim) {
  __shared__ float sp[32];
  griddep_wait();  // our own partial (previous kernel in the stream) is complete and visible
  const unsigned e = *epoch_cell + 1u;
  if (threadIdx.x < ws) {
    __threadfence_system();
    st_release_sys_u32(peers.flags[threadIdx.x] + rank, e);  // tell peer `threadIdx.x` that rank's partial e is ready
    const unsigned* mine = peers.flags[rank] + threadIdx.x;  // ... and wait for its partial
    const long long t0 = clock64();
    while ((int)(ld_acquire_sys_u32(mine) - e) < 0) {
      if (clock64() - t0 > 4000000000ll) __trap();  // a missing peer traps instead of hanging the GPU
    }
  }
  __syncthreads();
  float ss = 0.f;
  for (int i = threadIdx.x * 8; i < dim; i += blockDim.x * 8) {
    float a[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = 0.f;
    for (int r = 0; r < ws; ++r) {  // fixed rank order on every rank
      const uint4 v = ld_volatile_u4(peers.buf[r] + i);
      const __half2* h2 = reinterpret_cast<const __half2*>(&v);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __half22float2(h2[j]);
        a[2 * j] += f.x;
        a[2 * j + 1] += f.y;
      }
    }
    uint4 xv = *reinterpret_cast<const uint4*>(x + i);
    __half2* x2 = reinterpret_cast<__half2*>(&xv);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      x2[j] = __hadd2(x2[j], __floats2half2_rn(a[2 * j], a[2 * j + 1]));  // y rounds to fp16, then the fp16 residual add
      const float2 f = __half22float2(x2[j]);
      ss = fmaf(f.x, f.x, fmaf(f.y, f.y, ss));
    }
    *reinterpret_cast<uint4*>(x + i) = xv;
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, d);
  if ((threadIdx.x & 31) == 0) sp[threadIdx.x >> 5] = ss;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sp[w];
    sp[0] = t;
    *epoch_cell = e;
  }
  __syncthreads();
  if (ssq_out)
    for (int i = threadIdx.x; i < kSsqParts * kSsqReplicas; i += blockDim.x)  // every replica: {total, 0, 0, ...}
      ssq_out[i] = (i % kSsqParts == 0) ? sp[0] : 0.f;
  griddep_launch_dependents();
}

// ---------------------------------------------------------------------------------- launchers
template <typename... KArgs, typename... Args>
static int launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, const char* what,
                      Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = launch_ctx().pdl ? 1 : 0;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
  if (e != cudaSuccess) return set_cuda_error(e, what);
  return TEAL_OK;
}

int launch_embed(const __half* emb, const int* token, __half* x, float* ssq_out, int dim, cudaStream_t stream) {
  if (!emb || !token || !x || dim <= 0 || (dim & 7)) return set_error(TEAL_ERR_BAD_ARG, "embed: bad argument");
  return launch_pdl(embed_kernel, dim3(1), dim3(256), 0, stream, "embed launch", emb, token, x, ssq_out, dim);
}

int launch_attn_decode(const __half* qkv, const float* qkv_acc, const __half* freqs, __half* k_cache, __half* v_cache,
                       const int* pos, __half* out, int n_head, int n_kv, int head_dim, int max_seq, cudaStream_t stream) {
  if ((qkv == nullptr) == (qkv_acc == nullptr) || (reinterpret_cast<uintptr_t>(qkv_acc) & 7))
    return set_error(TEAL_ERR_BAD_ARG, "attn_decode: exactly one of qkv (fp16) / qkv_acc (fp32, 8-byte aligned)");
  if (!freqs || !k_cache || !v_cache || !pos || !out || n_head <= 0 || n_kv <= 0 || n_head % n_kv)
    return set_error(TEAL_ERR_BAD_ARG, "attn_decode: bad argument");
  if ((reinterpret_cast<uintptr_t>(qkv) & 3) || (reinterpret_cast<uintptr_t>(freqs) & 3) || (reinterpret_cast<uintptr_t>(k_cache) & 3) ||
      (reinterpret_cast<uintptr_t>(v_cache) & 3))
    return set_error(TEAL_ERR_ALIGNMENT, "attn_decode: 4-byte aligned vectors");
  const float scale = 1.0f / sqrtf((float)head_dim);
  if (head_dim == 128)
    return launch_pdl(attn_decode_kernel<128>, dim3(n_head), dim3(kAttnThreads), 0, stream, "attn launch", qkv, qkv_acc, freqs,
                      k_cache, v_cache, pos, out, n_head, n_kv, max_seq, scale);
  if (head_dim == 64)
    return launch_pdl(attn_decode_kernel<64>, dim3(n_head), dim3(kAttnThreads), 0, stream, "attn launch", qkv, qkv_acc, freqs,
                      k_cache, v_cache, pos, out, n_head, n_kv, max_seq, scale);
  return set_error(TEAL_ERR_UNSUPPORTED, "attn_decode: head_dim must be 64 or 128");
}

int launch_dense_rowmajor(const __half* x, const __half* norm_w, float eps, const __half* W, __half* y,
                          unsigned long long* argmax_packed, int Z, int V, long long ldw, cudaStream_t stream) {
  if (!x || !W || !y || Z <= 0 || V <= 0 || (Z & 7) || (ldw & 7) || ldw < Z)
    return set_error(TEAL_ERR_BAD_ARG, "dense gemv: bad argument");
  if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(W) & 15))
    return set_error(TEAL_ERR_ALIGNMENT, "dense gemv: x / W must be 16-byte aligned");
  const DeviceInfo& di = device_info();
  if (di.err != cudaSuccess) return set_cuda_error(di.err, "device query");
  const size_t smem = (size_t)Z * 2;
  auto kern = dense_rowmajor_gemv_kernel<8>;
  static bool attr_set = false;
  if (!attr_set && smem > 40 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    if (e != cudaSuccess) return set_cuda_error(e, "cudaFuncSetAttribute");
    attr_set = true;
  }
  if (smem > 100 * 1024) return set_error(TEAL_ERR_UNSUPPORTED, "dense gemv: Z too large");
  int grid = di.sm_count * 2;
  const int rows_per_cta = kHeadThreads / 32;
  if ((long long)grid * rows_per_cta > V) grid = (V + rows_per_cta - 1) / rows_per_cta;
  return launch_pdl(kern, dim3(grid), dim3(kHeadThreads), smem, stream, "dense gemv launch", x, norm_w, eps, W, y,
                    argmax_packed, Z, V, ldw);
}

int launch_residual_add(__half* x, const __half* part, float* ssq_out, int dim, cudaStream_t stream) {
  if (!x || !part || dim <= 0 || (dim & 1)) return set_error(TEAL_ERR_BAD_ARG, "residual_add: bad argument");
  return launch_pdl(residual_add_kernel, dim3(1), dim3(1024), 0, stream, "residual_add launch", x, part, ssq_out, dim);
}

int launch_tp_allreduce_residual(const void* const* bufs, void* const* flags, int rank, int ws, __half* x, float* ssq_out,
                                 unsigned* epoch_cell, int dim, cudaStream_t stream) {
  if (!bufs || !flags || !x || !epoch_cell || ws < 1 || ws > kMaxTpWorld || rank < 0 || rank >= ws || dim <= 0 || (dim & 7))
    return set_error(TEAL_ERR_BAD_ARG, "tp_allreduce_residual: bad argument");
  TpPeers p;
  for (int r = 0; r < kMaxTpWorld; ++r) {
    p.buf[r] = (r < ws) ? reinterpret_cast<const __half*>(bufs[r]) : nullptr;
    p.flags[r] = (r < ws) ? reinterpret_cast<unsigned*>(flags[r]) : nullptr;
    if (r < ws && (!p.buf[r] || !p.flags[r] || (reinterpret_cast<uintptr_t>(p.buf[r]) & 15)))
      return set_error(TEAL_ERR_ALIGNMENT, "tp_allreduce_residual: null / unaligned peer pointer");
  }
  return launch_pdl(tp_allreduce_residual_kernel, dim3(1), dim3(1024), 0, stream, "tp_allreduce_residual launch", p, rank,
                    ws, x, ssq_out, epoch_cell, dim);
}

int launch_advance(unsigned long long* argmax_packed, int* token, int* pos, int* out_tokens, int out_cap,
                   cudaStream_t stream) {
  if (!argmax_packed || !token || !pos || !out_tokens) return set_error(TEAL_ERR_BAD_ARG, "advance: bad argument");
  return launch_pdl(advance_kernel, dim3(1), dim3(32), 0, stream, "advance launch", argmax_packed, token, pos,
                    out_tokens, out_cap);
}

int launch_sample_topk(const __half* logits, int V, float temperature, int top_k, const __half* q_ext, unsigned long long seed,
                       const int* pos, unsigned long long* argmax_packed, __half* probs_out, cudaStream_t stream) {
  if (!logits || !argmax_packed || V <= 0 || top_k < 0 || !(temperature >= 0.f))
    return set_error(TEAL_ERR_BAD_ARG, "sample: bad argument");
  int threads = ((V + 7) / 8 + 31) / 32 * 32;  // ~8 values per thread per pass, one CTA
  threads = threads < 128 ? 128 : threads > kSampleMaxThreads ? kSampleMaxThreads : threads;
  return launch_pdl(sample_topk_kernel, dim3(1), dim3(threads), 0, stream, "sample launch", logits, V, temperature, top_k, q_ext,
                    seed, pos, argmax_packed, probs_out);
}

}  // namespace teal
