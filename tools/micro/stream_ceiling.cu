// What streaming rate can one SM-filling kernel reach on HBM-resident data (each byte read once), for
//   A: per-warp cp.async rings (LDGSTS)   B: LDG.128 into registers   C: TMA 1-D bulk copies into an mbarrier ring?
// Chunks are unique (multiplicative-hash permutation of chunk ids), addresses are computed arithmetically.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../teal_b200/csrc/common.cuh"
using namespace teal;
__device__ __forceinline__ size_t chunk_off(uint32_t id, uint32_t nchunks_mask, int bytes) {
  return (size_t)((id * 2654435761u) & nchunks_mask) * (size_t)bytes;   // odd multiplier: a permutation mod 2^k
}
template <int DEPTH>
__global__ void __launch_bounds__(1024, 1) k_ldgsts(const unsigned char* src, uint32_t mask, int per_warp, int bytes, float* sink) {
  extern __shared__ __align__(128) unsigned char sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  unsigned char* ring = sm + (size_t)warp * DEPTH * bytes;
  const uint32_t id0 = (blockIdx.x * nw + warp) * per_warp;
  float acc = 0.f;
  int islot = 0, cslot = 0;
  auto issue = [&](int i) {
    if (i < per_warp) {
      const unsigned char* s = src + chunk_off(id0 + i, mask, bytes) + lane * 16;
      const uint32_t d = smem_u32(ring + (size_t)islot * bytes) + lane * 16;
      for (int b = 0; b < bytes; b += 512)
        asm volatile("cp.async.cg.shared.global.L2::128B [%0], [%1], 16;" ::"r"(d + b), "l"(s + b) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    islot = (islot + 1 == DEPTH) ? 0 : islot + 1;
  };
  for (int i = 0; i < DEPTH - 1; ++i) issue(i);
  for (int i = 0; i < per_warp; ++i) {
    issue(i + DEPTH - 1);
    asm volatile("cp.async.wait_group %0;" ::"n"(DEPTH - 1) : "memory");
    __syncwarp();
    const unsigned char* r = ring + (size_t)cslot * bytes + lane * 16;
    for (int b = 0; b < bytes; b += 512) acc += *reinterpret_cast<const float*>(r + b);
    __syncwarp();
    cslot = (cslot + 1 == DEPTH) ? 0 : cslot + 1;
  }
  if (acc == 1.2345f) sink[0] = acc;
}
template <int U>
__global__ void __launch_bounds__(1024, 1) k_ldg(const unsigned char* src, uint32_t mask, int per_warp, int bytes, float* sink) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const uint32_t id0 = (blockIdx.x * nw + warp) * per_warp;
  const int vec_per_chunk = bytes / 512;  // LDG.128 warp-instructions per chunk
  float acc = 0.f;
  const int total = per_warp * vec_per_chunk;
  for (int t = 0; t < total; t += U) {
    uint4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int q = t + u;
      const int ci = q / vec_per_chunk, vi = q - ci * vec_per_chunk;
      v[u] = (q < total) ? ld_nc_u4(reinterpret_cast<const uint4*>(src + chunk_off(id0 + ci, mask, bytes) + vi * 512 + lane * 16)) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) acc += __uint_as_float(v[u].x) + __uint_as_float(v[u].w);
  }
  if (acc == 1.2345f) sink[0] = acc;
}
// TMA: `pw` producer warps, slot ring with full/empty mbarriers, 4 consumer warps just touch the data.
__global__ void __launch_bounds__(512, 1) k_tma(const unsigned char* src, uint32_t mask, int per_cta, int bytes, int slots, int pw, float* sink) {
  extern __shared__ __align__(128) unsigned char sm[];
  uint64_t* full = reinterpret_cast<uint64_t*>(sm);
  uint64_t* empty = full + 64;
  unsigned char* ring = sm + 1024;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ncons = (blockDim.x >> 5) - pw;
  for (int i = threadIdx.x; i < slots; i += blockDim.x) { mbar_init(smem_u32(&full[i]), 1); mbar_init(smem_u32(&empty[i]), ncons); }
  fence_mbar_init();
  __syncthreads();
  const uint32_t id0 = blockIdx.x * per_cta;
  const uint64_t pol = make_evict_first_policy();
  if (warp < pw) {
    int slot = warp % slots; uint32_t ph = 0; int wraps = 0;
    for (int i = warp; i < per_cta; i += pw) {
      // slot sequence for this warp: i % slots (pw divides slots)
      slot = i % slots; ph = (i / slots) & 1;
      mbar_wait(smem_u32(&empty[slot]), ph ^ 1u);
      if (elect_one()) {
        mbar_arrive_expect_tx(smem_u32(&full[slot]), bytes);
        bulk_g2s(smem_u32(ring + (size_t)slot * bytes), src + chunk_off(id0 + i, mask, bytes), bytes, smem_u32(&full[slot]), pol);
      }
    }
    (void)wraps;
  } else {
    float acc = 0.f;
    for (int i = 0; i < per_cta; ++i) {
      const int slot = i % slots; const uint32_t ph = (i / slots) & 1;
      mbar_wait(smem_u32(&full[slot]), ph);
      acc += *reinterpret_cast<const float*>(ring + (size_t)slot * bytes + ((warp - pw) * 32 + lane) * 4);
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&empty[slot]));
    }
    if (acc == 1.2345f) sink[0] = acc;
  }
}
int main() {
  int nsm = 0; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
  const size_t big = (size_t)1 << 30;   // 1 GiB, chunk ids are a permutation of [0, big/bytes)
  unsigned char* src; cudaMalloc(&src, big); cudaMemset(src, 0, big);
  float* sink; cudaMalloc(&sink, 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const size_t target = (size_t)512 << 20;  // bytes per launch
  auto run = [&](const char* name, int bytes, int a, int b, auto launch) {
    float ms = 0;
    for (int rep = 0; rep < 3; ++rep) { cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1); }
    cudaError_t e = cudaGetLastError();
    printf("%-7s chunk %5d B  p1 %3d p2 %3d : %7.1f us %6.0f GB/s %s\n", name, bytes, a, b, ms * 1e3, (double)target / (ms * 1e-3) / 1e9, e == cudaSuccess ? "" : cudaGetErrorString(e));
  };
#define SETSM(k) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024)
  SETSM(k_ldgsts<4>); SETSM(k_ldgsts<8>); SETSM(k_ldgsts<12>); SETSM(k_tma);
  for (int bytes : {512, 1024, 2048}) {
    const uint32_t mask = (uint32_t)(big / bytes) - 1;
    for (int warps : {16, 32}) {
      const int per_warp = (int)(target / ((size_t)nsm * warps * bytes));
      for (int depth : {4, 8, 12}) {
        const size_t smem = (size_t)warps * depth * bytes;
        if (smem > 220 * 1024) continue;
        if (depth == 4) run("ldgsts", bytes, warps, depth, [&] { k_ldgsts<4><<<nsm, warps * 32, smem>>>(src, mask, per_warp, bytes, sink); });
        if (depth == 8) run("ldgsts", bytes, warps, depth, [&] { k_ldgsts<8><<<nsm, warps * 32, smem>>>(src, mask, per_warp, bytes, sink); });
        if (depth == 12) run("ldgsts", bytes, warps, depth, [&] { k_ldgsts<12><<<nsm, warps * 32, smem>>>(src, mask, per_warp, bytes, sink); });
      }
    }
  }
  for (int bytes : {512, 2048}) {
    const uint32_t mask = (uint32_t)(big / bytes) - 1;
    for (int warps : {16, 32}) {
      const int per_warp = (int)(target / ((size_t)nsm * warps * bytes));
      run("ldg128", bytes, warps, 4, [&] { k_ldg<4><<<nsm, warps * 32>>>(src, mask, per_warp, bytes, sink); });
      run("ldg128", bytes, warps, 8, [&] { k_ldg<8><<<nsm, warps * 32>>>(src, mask, per_warp, bytes, sink); });
    }
  }
  for (int bytes : {2048, 4096, 8192, 16384}) {
    const uint32_t mask = (uint32_t)(big / bytes) - 1;
    const int per_cta = (int)(target / ((size_t)nsm * bytes));
    int slots = 16; while ((size_t)slots * 2 * bytes <= 200 * 1024 && slots < 64) slots *= 2;
    while ((size_t)slots * bytes > 200 * 1024) slots /= 2;
    for (int pw : {4, 8}) run("tma", bytes, pw, slots, [&] { k_tma<<<nsm, (pw + 4) * 32, 1024 + (size_t)slots * bytes>>>(src, mask, per_cta, bytes, slots, pw, sink); });
  }
  return 0;
}
