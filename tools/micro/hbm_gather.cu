// HBM gather bandwidth vs contiguous chunk size: every warp streams random chunks of `bytes` from a
// 4 GiB buffer (no reuse, L2 cold) through a cp.async ring.  One CTA per SM, 16 warps, DEPTH rows in flight.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include <stdint.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int DEPTH>
__global__ void __launch_bounds__(512, 1) gather(const unsigned char* src, const uint32_t* offs, int per_warp, int bytes,
                                                 int cta_chunk, float* sink) {
  extern __shared__ __align__(128) unsigned char sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // "cta_chunk" mode: the 16 warps of a CTA fetch adjacent `bytes` pieces of one contiguous cta-wide chunk
  const uint32_t* my = offs + ((size_t)blockIdx.x * 16 + (cta_chunk ? 0 : warp)) * per_warp;
  unsigned char* ring = sm + (size_t)warp * DEPTH * bytes;
  float acc = 0.f;
  auto issue = [&](int i) {
    if (i < per_warp) {
      const unsigned char* s = src + ((size_t)my[i] << 9) + (cta_chunk ? (size_t)warp * bytes : 0);
      const uint32_t d = smem_u32(ring + (size_t)(i % DEPTH) * bytes);
      for (int b = lane * 16; b < bytes; b += 512)
        asm volatile("cp.async.cg.shared.global.L2::128B [%0], [%1], 16;" ::"r"(d + b), "l"(s + b) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  for (int i = 0; i < DEPTH - 1; ++i) issue(i);
  for (int i = 0; i < per_warp; ++i) {
    issue(i + DEPTH - 1);
    asm volatile("cp.async.wait_group %0;" ::"n"(DEPTH - 1) : "memory");
    __syncwarp();
    const unsigned char* r = ring + (size_t)(i % DEPTH) * bytes;
    for (int b = lane * 16; b < bytes; b += 512) acc += *reinterpret_cast<const float*>(r + b);
    __syncwarp();
  }
  if (acc == 1.2345f) sink[0] = acc;
}
int main() {
  int nsm = 0; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
  const size_t big = (size_t)4 << 30;
  unsigned char* src; cudaMalloc(&src, big); cudaMemset(src, 0, big);
  float* sink; cudaMalloc(&sink, 4);
  cudaFuncSetAttribute(gather<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
  cudaFuncSetAttribute(gather<24>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const size_t total_target = (size_t)256 << 20;  // bytes fetched per launch
  for (int cta_chunk = 0; cta_chunk < 2; ++cta_chunk)
  for (int bytes : {512, 1024, 1536}) {
    const int per_warp = (int)(total_target / ((size_t)nsm * 16 * bytes));
    const size_t nent = (size_t)nsm * 16 * per_warp;
    std::vector<uint32_t> h(nent);
    uint64_t rng = 88172645463325252ull;
    const size_t span = cta_chunk ? (size_t)16 * bytes : bytes;      // contiguous span per gather
    const size_t slots = (big - (1 << 20)) / 32768;                  // 32 KB-spaced candidate rows
    for (size_t i = 0; i < nent; ++i) {
      rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17;
      size_t off = (rng % slots) * 32768 + ((rng >> 40) % (32768 / ((span + 511) / 512 * 512) )) * ((span + 511) / 512 * 512);
      h[i] = (uint32_t)(off >> 9);
    }
    uint32_t* d; cudaMalloc(&d, nent * 4); cudaMemcpy(d, h.data(), nent * 4, cudaMemcpyHostToDevice);
    for (int depth : {8, 24}) {
      if ((size_t)16 * depth * bytes > 220 * 1024) continue;
      float ms = 0;
      for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        if (depth == 8) gather<8><<<nsm, 512, (size_t)16 * depth * bytes>>>(src, d, per_warp, bytes, cta_chunk, sink);
        else gather<24><<<nsm, 512, (size_t)16 * depth * bytes>>>(src, d, per_warp, bytes, cta_chunk, sink);
        cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
      }
      printf("chunk/warp %5d B  contiguous span %6zu B  depth %2d (%3d KB in flight/SM): %7.1f us  %6.0f GB/s  [%s]\n", bytes, span, depth,
             16 * (depth - 1) * bytes / 1024, ms * 1e3, (double)nent * bytes / (ms * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
    }
    cudaFree(d);
  }
  return 0;
}
