// Microbenchmark: per-SM issue/throughput rate of 1-D TMA bulk copies vs cp.async (LDGSTS) vs LDG.128,
// gathering row segments from an L2-resident / HBM-resident matrix.  Build: nvcc -arch=sm_100a.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../teal_b200/csrc/common.cuh"
using namespace teal;

// Each CTA: `nwarps` producer warps issue `ncopies` bulk copies of `bytes` each (round-robin into a smem ring
// of `slots` slots, one mbarrier per slot, a slot is re-used only after its previous copy has landed).
__global__ void tma_kernel(const unsigned char* src, size_t src_bytes, int bytes, int ncopies, int slots, int hint,
                           long long* out) {
  extern __shared__ __align__(128) unsigned char sm[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm);
  unsigned char* ring = sm + 1024;  // [0,512): 64 mbarriers, [512,1024): offsets
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int i = threadIdx.x; i < slots; i += blockDim.x) mbar_init(smem_u32(&bars[i]), 1);
  fence_mbar_init();
  __syncthreads();
  const uint64_t pol = make_evict_first_policy();
  // slots is a power of two; offsets precomputed so the timed loop is LDS + issue only
  uint32_t* offs = reinterpret_cast<uint32_t*>(sm + 512);  // up to 128 entries (re-used modulo 128)
  if (threadIdx.x < 128) {
    uint32_t rng = (blockIdx.x * 7919u + threadIdx.x * 104729u + 12345u) * 1664525u + 1013904223u;
    offs[threadIdx.x] = (uint32_t)(((size_t)(rng >> 4) * 64u + blockIdx.x * 32u) % ((src_bytes - 65536) >> 7));
  }
  __syncthreads();
  const int smask = slots - 1;
  int sshift = 0; while ((1 << sshift) < slots) ++sshift;
  long long t0 = clock64();
  for (int i = warp; i < ncopies; i += nw) {
    const int slot = i & smask;
    const int use = i >> sshift;  // how many times this slot was used before
    if (use > 0) mbar_wait(smem_u32(&bars[slot]), (use - 1) & 1);
    size_t off = (size_t)offs[i & 127] << 7;
    if (elect_one()) {
      mbar_arrive_expect_tx(smem_u32(&bars[slot]), bytes);
      if (hint) bulk_g2s(smem_u32(ring + (size_t)slot * bytes), src + off, bytes, smem_u32(&bars[slot]), pol);
      else asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                        ::"r"(smem_u32(ring + (size_t)slot * bytes)), "l"(src + off), "r"(bytes), "r"(smem_u32(&bars[slot])) : "memory");
    }
  }
  long long t1 = clock64();
  // drain: wait for the last use of every slot this warp touched
  for (int i = warp; i < ncopies; i += nw) {
    const int slot = i & smask, use = i >> sshift;
    if (i + slots >= ncopies) mbar_wait(smem_u32(&bars[slot]), use & 1);
  }
  __syncthreads();
  long long t2 = clock64();
  if (threadIdx.x == 0) { out[blockIdx.x * 2] = t1 - t0; out[blockIdx.x * 2 + 1] = t2 - t0; }
}

// cp.async 16B per lane: each warp fetches row segments of `bytes` (bytes/512 instructions per row), depth groups in flight
__global__ void ldgsts_kernel(const unsigned char* src, size_t src_bytes, int bytes, int ncopies, int depth, long long* out) {
  extern __shared__ __align__(128) unsigned char sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  uint32_t* offs = reinterpret_cast<uint32_t*>(sm);
  unsigned char* mine = sm + 512 + (size_t)warp * depth * bytes;
  if (threadIdx.x < 128) {
    uint32_t rng = (blockIdx.x * 7919u + threadIdx.x * 104729u + 12345u) * 1664525u + 1013904223u;
    offs[threadIdx.x] = (uint32_t)(((size_t)(rng >> 4) * 64u + blockIdx.x * 32u) % ((src_bytes - 65536) >> 7));
  }
  __syncthreads();
  long long t0 = clock64();
  int issued = 0, dslot = 0;
  for (int i = warp; i < ncopies; i += nw, ++issued) {
    size_t off = (size_t)offs[i & 127] << 7;
    unsigned char* dst = mine + (size_t)dslot * bytes;
    if (++dslot == depth) dslot = 0;
    if (issued >= depth) asm volatile("cp.async.wait_group %0;" ::"n"(7) : "memory");
    for (int b = lane * 16; b < bytes; b += 512)
      asm volatile("cp.async.cg.shared.global.L2::128B [%0], [%1], 16;" ::"r"(smem_u32(dst + b)), "l"(src + off + b) : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  long long t1 = clock64();
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  long long t2 = clock64();
  if (threadIdx.x == 0) { out[blockIdx.x * 2] = t1 - t0; out[blockIdx.x * 2 + 1] = t2 - t0; }
}

// LDG.128 into registers, 8 loads in flight per lane
__global__ void ldg_kernel(const unsigned char* src, size_t src_bytes, int bytes, int ncopies, long long* out, float* sink) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  __shared__ uint32_t offs[128];
  if (threadIdx.x < 128) {
    uint32_t rng = (blockIdx.x * 7919u + threadIdx.x * 104729u + 12345u) * 1664525u + 1013904223u;
    offs[threadIdx.x] = (uint32_t)(((size_t)(rng >> 4) * 64u + blockIdx.x * 32u) % ((src_bytes - 65536) >> 7));
  }
  __syncthreads();
  long long t0 = clock64();
  float acc = 0.f;
  int oi = warp;
  const int per_row = bytes / 512;  // LDG.128 warp-instructions per row segment
  for (int i = warp; i < ncopies; i += nw * 8 / (per_row < 8 ? per_row : 8) ) {
    uint4 v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if ((j % per_row) == 0) oi += 7;
      size_t off = (size_t)offs[oi & 127] << 7;
      v[j] = ld_nc_u4(reinterpret_cast<const uint4*>(src + off + (j % per_row) * 512 + lane * 16));
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) acc += __uint_as_float(v[j].x) + __uint_as_float(v[j].w);
  }
  long long t1 = clock64();
  if (acc == 123.456f) sink[0] = acc;
  __syncthreads();
  long long t2 = clock64();
  if (threadIdx.x == 0) { out[blockIdx.x * 2] = t1 - t0; out[blockIdx.x * 2 + 1] = t2 - t0; }
}

int main() {
  int nsm = 0; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
  size_t big = (size_t)2 << 30, small = (size_t)48 << 20;
  unsigned char* src; cudaMalloc(&src, big); cudaMemset(src, 1, big);
  long long* out; cudaMalloc(&out, nsm * 16); float* sink; cudaMalloc(&sink, 4);
  long long* h = (long long*)malloc(nsm * 16);
  cudaFuncSetAttribute(tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  cudaFuncSetAttribute(ldgsts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto report = [&](const char* name, int bytes, int ncopies, int warps, const char* where, int extra, float ms) {
    cudaError_t ce = cudaMemcpy(h, out, nsm * 16, cudaMemcpyDeviceToHost);
    if (ce != cudaSuccess) { printf("%s B=%d n=%d w=%d %s: ERROR %s\n", name, bytes, ncopies, warps, where, cudaGetErrorString(ce)); exit(1); }
    double issue = 0, total = 0; for (int i = 0; i < nsm; ++i) { issue += h[2 * i]; total += h[2 * i + 1]; }
    issue /= nsm; total /= nsm;
    printf("%-6s B=%4d n=%4d w=%2d %-3s x%d | issue %5.1f cyc/copy | %5.1f B/clk/SM | %6.2f us %5.0f GB/s\n",
           name, bytes, ncopies, warps, where, extra, issue / ncopies, (double)bytes * ncopies / total, ms * 1e3,
           (double)bytes * ncopies * nsm / (ms * 1e-3) / 1e9);
  };
  for (int wi = 0; wi < 1; ++wi) {
    int where = 1 - wi;
    size_t sb = where ? big : small;
    const char* wn = where ? "HBM" : "L2";
    for (int bytes : {512, 1024, 2048, 4096, 8192}) {
      int per_sm = (192 * 1024) / bytes;           // everything fits in the ring: pure issue + latency
      int slots = 64; while ((size_t)slots * bytes > 196 * 1024) slots >>= 1;
      for (int warps : {1, 2, 4, 8}) for (int hint : {1}) for (int mult : {1, 4}) {
        int nc = per_sm * mult;
        for (int rep = 0; rep < 2; ++rep) {
          cudaEventRecord(e0);
          tma_kernel<<<nsm, warps * 32, 1024 + (size_t)slots * bytes>>>(src, sb, bytes, nc, slots, hint, out);
          cudaEventRecord(e1); cudaEventSynchronize(e1);
        }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        report(hint ? "tma+h" : "tma", bytes, nc, warps, wn, mult, ms);
      }
    }
    for (int bytes : {512, 2048}) for (int warps : {4, 16}) for (int mult : {1, 4}) {
      int depth = 8; int nc = (192 * 1024 / bytes) * mult;
      size_t smem = 512 + (size_t)warps * depth * bytes; if (smem > 200 * 1024) { depth = 199 * 1024 / (warps * bytes); smem = 512 + (size_t)warps * depth * bytes; }
      for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        ldgsts_kernel<<<nsm, warps * 32, smem>>>(src, sb, bytes, nc, depth, out);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
      }
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      report("ldgsts", bytes, nc, warps, wn, depth, ms);
    }
    for (int bytes : {512, 2048}) for (int warps : {16, 32}) for (int mult : {1, 4}) {
      int nc = (192 * 1024 / bytes) * mult;
      for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        ldg_kernel<<<nsm, warps * 32>>>(src, sb, bytes, nc, out, sink);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
      }
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      report("ldg128", bytes, nc, warps, wn, mult, ms);
    }
  }
  printf("err %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
