// Issue rate of the three ways to do  f32 += f16 * f16  on sm_100a, per SM sub-partition, registers only:
//   A: HADD2.F32 conversions + FFMA2 (rowchunk_gemv_kernel)      B: FHFMA, fma.rn.f32.f16 (rowchunk_gemv2_kernel, FH)
//   C: conversions + scalar FFMA
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fhfma_rate fhfma_rate.cu ; run on the GPU box.
// Output: fp16 weights consumed per clock per SM for 1..16 warps per SM sub-partition.  The decode GEMVs need
// (HBM 6.5 TB/s / 2 B) / 148 SMs / 1.9 GHz = 11.6 weights per clock per SM; anything close to that is a limiter.
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ void ffma2(float2& acc, const float2 w, const float2 x) {
  asm("fma.rn.f32x2 %0, %1, %2, %0;"
      : "+l"(*reinterpret_cast<unsigned long long*>(&acc))
      : "l"(*reinterpret_cast<const unsigned long long*>(&w)), "l"(*reinterpret_cast<const unsigned long long*>(&x)));
}
__device__ __forceinline__ void fhfma2(float2& acc, uint32_t w2, unsigned short xh) {
  asm("{\n.reg .b16 lo, hi;\nmov.b32 {lo, hi}, %2;\nfma.rn.f32.f16 %0, lo, %3, %0;\nfma.rn.f32.f16 %1, hi, %3, %1;\n}\n"
      : "+f"(acc.x), "+f"(acc.y)
      : "r"(w2), "h"(xh));
}

template <int MODE>
__global__ void k(const uint4* src, float* sink, int iters, long long* cycles) {
  uint4 w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) w[i] = src[threadIdx.x * 4 + i];
  float2 acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = make_float2(0.f, 0.f);
  const unsigned short xh = (unsigned short)(0x3c00 + (threadIdx.x & 3));
  const float xs = __half2float(__ushort_as_half(xh));
  const float2 x2 = make_float2(xs, xs);
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint32_t q[4] = {w[i].x, w[i].y, w[i].z, w[i].w};
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (MODE == 0) {
          ffma2(acc[i * 4 + c], __half22float2(*reinterpret_cast<const __half2*>(&q[c])), x2);
        } else if (MODE == 1) {
          fhfma2(acc[i * 4 + c], q[c], xh);
        } else {
          const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&q[c]));
          acc[i * 4 + c].x = fmaf(f.x, xs, acc[i * 4 + c].x);
          acc[i * 4 + c].y = fmaf(f.y, xs, acc[i * 4 + c].y);
        }
      }
      w[i].x ^= (uint32_t)it;  // keep the conversions inside the loop
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i].x + acc[i].y;
  if (s == 1.2345f) sink[0] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

int main() {
  uint4* src;
  float* sink;
  long long* cyc;
  cudaMalloc(&src, 1024 * 4 * sizeof(uint4));
  cudaMemset(src, 0x3c, 1024 * 4 * sizeof(uint4));
  cudaMalloc(&sink, 4);
  cudaMalloc(&cyc, 148 * sizeof(long long));
  const int iters = 4096;
  const char* names[3] = {"cvt+FFMA2", "FHFMA", "cvt+FFMA"};
  for (int warps = 4; warps <= 32; warps *= 2) {  // per SM (4 sub-partitions)
    for (int mode = 0; mode < 3; ++mode) {
      void (*kern)(const uint4*, float*, int, long long*) = mode == 0 ? k<0> : mode == 1 ? k<1> : k<2>;
      kern<<<148, warps * 32>>>(src, sink, iters, cyc);
      cudaDeviceSynchronize();
      long long c[148];
      cudaMemcpy(c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
      double avg = 0;
      for (int i = 0; i < 148; ++i) avg += (double)c[i];
      avg /= 148;
      const double weights = (double)iters * 32.0 * warps * 32.0;  // 32 fp16 weights per thread per iteration
      printf("%-10s %2d warps/SM: %.1f fp16 weights / clk / SM  (%.0f cycles)\n", names[mode], warps, weights / avg, avg);
    }
  }
  return 0;
}
