// Stand-alone A/B of the cluster-reduced gather-GEMV (teal_b200/csrc/cluster_gemv.cuh) against the library's row-chunk
// kernel (RED + arrival counter) on the BASELINE projection shapes: checks every configuration against a naive kernel,
// then times a CUDA graph of back-to-back PDL launches over rotating weight copies (L2 defeated).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o tools/micro/cluster_gemv_bench tools/micro/cluster_gemv_bench.cu -ldl
//   tools/micro/cluster_gemv_bench [--lib teal_b200/libteal_b200.so] [--shapes o,qkv,down,gate_up] [--out file.jsonl]
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "cluster_gemv.cuh"

using namespace teal;

#define CK(x)                                                                                   \
  do {                                                                                          \
    cudaError_t e_ = (x);                                                                       \
    if (e_ != cudaSuccess) {                                                                    \
      fprintf(stderr, "CUDA error %s at %s:%d: %s\n", #x, __FILE__, __LINE__, cudaGetErrorString(e_)); \
      exit(2);                                                                                  \
    }                                                                                           \
  } while (0)

namespace teal {  // host symbols common.cuh declares but this tool does not link
const DeviceInfo& device_info() { static DeviceInfo d{}; return d; }
}

struct Shape {
  const char* name;
  int Z, N;
  bool pair, norm, resid;
};
static const Shape kShapes[] = {
    {"o", 4096, 4096, false, false, true},       {"qkv", 4096, 6144, false, true, false},
    {"down", 14336, 4096, false, false, true},   {"gate_up", 4096, 14336, true, true, false},
    {"o70", 8192, 8192, false, false, true},     {"qkv70", 8192, 10240, false, true, false},
    {"down70", 28672, 8192, false, false, true}, {"gate_up70", 8192, 28672, true, true, false},
    {"o70tp8", 1024, 8192, false, false, true},  {"qkv70tp8", 8192, 1280, false, true, false},
    {"down70tp8", 3584, 8192, false, false, true}, {"gate_up70tp8", 8192, 3584, true, true, false},
};

typedef void (*KernFn)(const CgParams);
struct Variant {
  const char* name;
  KernFn fn;
  int vec, u, warps;
  bool pair, fh;
};
#define V(VEC, U, W, PAIR, FH, DB) {#VEC "x" #U "x" #W #PAIR #FH #DB, cluster_gemv_kernel<VEC, U, W, PAIR, FH, DB>, VEC, U, W, PAIR, FH}
static const Variant kVariants[] = {
    V(1, 8, 16, false, false, false), V(1, 8, 8, false, false, false), V(2, 4, 16, false, false, false), V(1, 8, 4, false, false, false),
    V(1, 8, 16, true, true, false),   V(2, 4, 16, true, true, false),  V(1, 8, 8, true, true, false),
};

// naive check: one thread per column, fp32
__global__ void ref_kernel(const __half* x, const __half* nw, float rs, int norm, const __half* w, long long ldw, int Z, int N, float thr,
                           float* y) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float a = 0.f;
  for (int m = 0; m < Z; ++m) {
    float xf = __half2float(x[m]);
    if (norm) xf = __half2float(__hmul(__float2half_rn(xf * rs), nw[m]));
    if (!(fabsf(xf) <= thr)) a = fmaf(__half2float(w[(long long)m * ldw + n]), xf, a);
  }
  y[n] = a;
}

static uint64_t rng_state = 0x243F6A8885A308D3ull;
static inline uint32_t rnd() {
  rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17;
  return (uint32_t)(rng_state >> 32);
}
static inline float gauss() {
  float u1 = (rnd() + 1.0f) / 4294967297.0f, u2 = rnd() / 4294967296.0f;
  return sqrtf(-2.f * logf(u1)) * cosf(6.2831853f * u2);
}
__global__ void fill_kernel(__half* p, size_t n, uint32_t seed, float scale) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  for (; i < n; i += (size_t)gridDim.x * blockDim.x) {
    uint32_t h = (uint32_t)i * 2654435761u ^ seed;
    h ^= h >> 16; h *= 0x7feb352du; h ^= h >> 15; h *= 0x846ca68bu; h ^= h >> 16;
    p[i] = __float2half_rn(((h & 0xffff) / 32768.0f - 1.0f) * scale);
  }
}

static int launch_cg(const Variant& v, const CgParams& p, int C, cudaStream_t s, bool pdl) {
  const int cols = 256 * v.vec;
  size_t smem = (size_t)6 * p.lanes_per_cta;
  const size_t red = (size_t)v.warps * cols * 4;
  if (red > smem) smem = red;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(p.ntiles * C));
  cfg.blockDim = dim3(v.warps * 32);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = C; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 2 : 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, v.fn, p);
  if (e != cudaSuccess) { fprintf(stderr, "launch %s C=%d: %s\n", v.name, C, cudaGetErrorString(e)); cudaGetLastError(); return 1; }
  return 0;
}

int main(int argc, char** argv) {
  std::string libpath = "teal_b200/libteal_b200.so", shapes = "o,qkv,down,gate_up", outpath = "gpurun_out/cluster_gemv_bench.jsonl";
  int launches = 48;
  bool pdl = true;
  for (int i = 1; i < argc; ++i) {
    if (!strcmp(argv[i], "--lib")) libpath = argv[++i];
    else if (!strcmp(argv[i], "--shapes")) shapes = argv[++i];
    else if (!strcmp(argv[i], "--out")) outpath = argv[++i];
    else if (!strcmp(argv[i], "--launches")) launches = atoi(argv[++i]);
    else if (!strcmp(argv[i], "--no-pdl")) pdl = false;
  }
  FILE* out = fopen(outpath.c_str(), "a");
  // the library's production kernel as the baseline (through its C ABI)
  typedef int (*gemv_fn)(const void*, const void*, void*, int, int, int64_t, double, void*, size_t, void*);
  typedef size_t (*wsb_fn)(int);
  typedef int (*wsi_fn)(void*, size_t, void*);
  typedef int (*pdl_fn)(int);
  void* lib = dlopen(libpath.c_str(), RTLD_NOW);
  gemv_fn lib_gemv = lib ? (gemv_fn)dlsym(lib, "teal_sparse_gemv_f16") : nullptr;
  wsb_fn lib_wsb = lib ? (wsb_fn)dlsym(lib, "teal_workspace_bytes") : nullptr;
  wsi_fn lib_wsi = lib ? (wsi_fn)dlsym(lib, "teal_workspace_init") : nullptr;
  pdl_fn lib_pdl = lib ? (pdl_fn)dlsym(lib, "teal_set_pdl") : nullptr;
  if (!lib_gemv) fprintf(stderr, "library baseline unavailable (%s)\n", dlerror());

  cudaStream_t st;
  CK(cudaStreamCreate(&st));
  for (KernFn f : {(KernFn) nullptr}) (void)f;
  for (const Variant& v : kVariants) {
    CK(cudaFuncSetAttribute(v.fn, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    CK(cudaFuncSetAttribute(v.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
  }
  const int kDbgCtas = 1024;
  long long* DBG;
  CK(cudaMalloc(&DBG, (size_t)64 * kDbgCtas * 8 * sizeof(long long)));
  std::vector<long long> hdbg((size_t)64 * kDbgCtas * 8);
  void* ws = nullptr;
  size_t wsb = 0;
  if (lib_wsb) {
    wsb = lib_wsb(65536);
    CK(cudaMalloc(&ws, wsb));
    lib_wsi(ws, wsb, st);
    if (lib_pdl) lib_pdl(pdl ? 1 : 0);
  }

  for (const Shape& sh : kShapes) {
    if ((("," + shapes + ",").find(std::string(",") + sh.name + ",")) == std::string::npos) continue;
    const int Z = sh.Z, N = sh.N, nmat = sh.pair ? 2 : 1;
    const size_t per = (size_t)Z * N * 2 * nmat;
    int ncopies = (int)((640ull << 20) / per);
    if (ncopies < 2) ncopies = 2;
    if (ncopies > 32) ncopies = 32;
    __half* W;
    CK(cudaMalloc(&W, per * ncopies));
    fill_kernel<<<2048, 256, 0, st>>>(W, per * ncopies / 2, 12345u, 0.04f);
    __half *X, *NW, *Y, *R;
    float *SSQ, *SSQO, *YR;
    CK(cudaMalloc(&X, (size_t)ncopies * Z * 2));
    CK(cudaMalloc(&NW, (size_t)Z * 2));
    CK(cudaMalloc(&Y, (size_t)N * 2 * 2));
    CK(cudaMalloc(&R, (size_t)N * 2));
    CK(cudaMalloc(&SSQ, (size_t)ncopies * 512 * 4));
    CK(cudaMalloc(&SSQO, 512 * 4));
    CK(cudaMalloc(&YR, (size_t)N * 4 * 2));
    std::vector<__half> hx((size_t)ncopies * Z), hnw(Z), hr(N);
    std::vector<float> hssq((size_t)ncopies * 512, 0.f), rs(ncopies, 1.f);
    for (int c = 0; c < ncopies; ++c) {
      double ss = 0;
      for (int m = 0; m < Z; ++m) {
        const float g = gauss();
        hx[(size_t)c * Z + m] = __float2half_rn(g);
        const float gh = __half2float(hx[(size_t)c * Z + m]);
        ss += (double)gh * gh;
      }
      for (int i = 0; i < 128; ++i) hssq[(size_t)c * 512 + i] = (float)(ss / 128.0);
      float t = 0.f;
      for (int i = 0; i < 512; ++i) t += hssq[(size_t)c * 512 + i];  // (close enough to the kernel's order for the check's tolerance)
      rs[c] = 1.0f / sqrtf(t / Z + 1e-5f);
    }
    for (int m = 0; m < Z; ++m) hnw[m] = __float2half_rn(1.0f + 0.1f * gauss());
    for (int n = 0; n < N; ++n) hr[n] = __float2half_rn(gauss());
    CK(cudaMemcpy(X, hx.data(), hx.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(NW, hnw.data(), hnw.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(R, hr.data(), hr.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(SSQ, hssq.data(), hssq.size() * 4, cudaMemcpyHostToDevice));
    // threshold: the median of |x| (|N(0,1)| -> 0.6745; after the norm the scale is ~1 as well)
    const float thr = __half2float(__float2half_rn(0.6745f));

    auto params = [&](const Variant& v, int C, int copy) {
      CgParams p = {};
      p.x = X + (size_t)copy * Z;
      p.norm_w = NW;
      p.ssq_in = SSQ + (size_t)copy * 512;
      p.n_ssq_in = 512;
      p.y = Y;
      p.resid = R;
      p.ssq_out = SSQO;
      p.ssq_slots = 512;
      p.Z = Z;
      p.nseg = nmat;
      p.mode = sh.pair ? kModePairSilu : kModePlain;
      p.in_mode = sh.norm ? kInRmsNorm : kInPlain;
      p.out_mode = sh.resid ? kOutAddResidual : kOutPlain;
      p.eps = 1e-5f;
      const int cols = 256 * v.vec;
      p.ntiles = (N + cols - 1) / cols;
      p.lanes_per_cta = (((Z + C - 1) / C) + 31) / 32 * 32;
      for (int s = 0; s < nmat; ++s) {
        p.seg[s].w = W + (size_t)copy * Z * N * nmat + (size_t)s * Z * N;
        p.seg[s].ldw = N;
        p.seg[s].ncols = N;
        p.seg[s].y_col0 = 0;
        p.seg[s].tile0 = 0;
        p.seg[s].thr = thr;
      }
      return p;
    };

    // reference for copy 0
    std::vector<float> yref((size_t)N * 2);
    for (int s = 0; s < nmat; ++s)
      ref_kernel<<<(N + 127) / 128, 128, 0, st>>>(X, NW, rs[0], sh.norm ? 1 : 0, W + (size_t)s * Z * N, N, Z, N, thr, YR + (size_t)s * N);
    CK(cudaMemcpyAsync(yref.data(), YR, (size_t)N * 4 * nmat, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    std::vector<float> want(N);
    for (int n = 0; n < N; ++n) {
      if (sh.pair) {
        const float g = __half2float(__float2half_rn(yref[n])), u = __half2float(__float2half_rn(yref[N + n]));
        const float sl = __half2float(__float2half_rn(g / (1.0f + expf(-g))));
        want[n] = __half2float(__float2half_rn(sl * u));
      } else {
        want[n] = __half2float(__float2half_rn(yref[n]));
        if (sh.resid) want[n] = __half2float(__hadd(hr[n], __float2half_rn(yref[n])));
      }
    }
    int nnz = 0;
    for (int m = 0; m < Z; ++m) {
      float xf = __half2float(hx[m]);
      if (sh.norm) xf = __half2float(__hmul(__float2half_rn(xf * rs[0]), hnw[m]));
      nnz += !(fabsf(xf) <= thr);
    }
    const double alg = (double)nnz * N * 2 * nmat + Z * 2.0 + N * 2.0;

    auto time_graph = [&](auto&& enqueue) -> double {
      cudaGraph_t g;
      cudaGraphExec_t ge;
      CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeGlobal));
      for (int i = 0; i < launches; ++i) enqueue(i % ncopies);
      CK(cudaStreamEndCapture(st, &g));
      CK(cudaGraphInstantiate(&ge, g, 0));
      cudaEvent_t e0, e1;
      CK(cudaEventCreate(&e0));
      CK(cudaEventCreate(&e1));
      CK(cudaGraphLaunch(ge, st));
      CK(cudaStreamSynchronize(st));
      double tot = 0;
      const int reps = 5;
      for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0, st));
        CK(cudaGraphLaunch(ge, st));
        CK(cudaEventRecord(e1, st));
        CK(cudaStreamSynchronize(st));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        tot += ms;
      }
      CK(cudaGraphExecDestroy(ge));
      CK(cudaGraphDestroy(g));
      return tot / reps * 1e3 / launches;
    };

    if (lib_gemv && !sh.pair) {
      const double us = time_graph([&](int c) { lib_gemv(X + (size_t)c * Z, W + (size_t)c * Z * N, Y, Z, N, N, (double)thr, ws, wsb, st); });
      char line[512];
      snprintf(line, sizeof line, "{\"shape\": \"%s\", \"Z\": %d, \"N\": %d, \"kernel\": \"library rowchunk (plain in/out)\", \"us\": %.3f, \"gbs\": %.0f, \"nnz\": %d}",
               sh.name, Z, N, us, alg / us / 1e3, nnz);
      puts(line); fflush(stdout);
      if (out) { fprintf(out, "%s\n", line); fflush(out); }
    }

    for (const Variant& v : kVariants) {
      if (v.pair != sh.pair) continue;
      for (int C : {2, 4, 8, 16}) {
        const int cols = 256 * v.vec;
        if (cols % C) continue;
        CgParams p0 = params(v, C, 0);
        if (p0.lanes_per_cta > 32 * kCgMaxGroups) continue;
        if ((long long)p0.ntiles * C > 148 * 4) continue;
        if (C < 4) continue;  // more than a wave: not this design's operating point
        if (p0.lanes_per_cta / 32 < v.warps / 2 && C > 2) continue;  // fewer lane groups than half the warps: pointless
        int max_clusters = 0;
        {
          cudaLaunchConfig_t cfg = {};
          cfg.gridDim = dim3(p0.ntiles * C);
          cfg.blockDim = dim3(v.warps * 32);
          size_t smem = (size_t)6 * p0.lanes_per_cta, red = (size_t)v.warps * cols * 4;
          cfg.dynamicSmemBytes = smem > red ? smem : red;
          cudaLaunchAttribute attr[1];
          attr[0].id = cudaLaunchAttributeClusterDimension;
          attr[0].val.clusterDim.x = C; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
          cfg.attrs = attr;
          cfg.numAttrs = 1;
          if (cudaOccupancyMaxActiveClusters(&max_clusters, v.fn, &cfg) != cudaSuccess) { cudaGetLastError(); max_clusters = -1; }
        }
        CK(cudaMemsetAsync(Y, 0xff, (size_t)N * 2, st));
        if (launch_cg(v, p0, C, st, false)) continue;
        cudaError_t e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) { fprintf(stderr, "%s C=%d failed: %s\n", v.name, C, cudaGetErrorString(e)); return 3; }
        std::vector<__half> hy(N);
        CK(cudaMemcpy(hy.data(), Y, (size_t)N * 2, cudaMemcpyDeviceToHost));
        double maxerr = 0, maxref = 0;
        int bad = 0;
        for (int n = 0; n < N; ++n) {
          const double d = fabs((double)__half2float(hy[n]) - want[n]);
          const double tol = fabs(want[n]) * (sh.pair ? 4e-3 : 2e-3) + (sh.norm ? 2e-2 : 4e-3);  // (norm: rs differs in the last bits -> a lane may flip)
          if (!(d <= tol)) ++bad;
          if (d > maxerr) maxerr = d;
          if (fabs(want[n]) > maxref) maxref = fabs(want[n]);
        }
        const double us = time_graph([&](int c) { launch_cg(v, params(v, C, c), C, st, pdl); });
        // phase timeline: the same chain once more with per-CTA globaltimer stamps (one slot per launch)
        char tl[400] = "";
        const int grid = p0.ntiles * C;
        if (grid <= kDbgCtas && launches <= 64) {
          CK(cudaMemsetAsync(DBG, 0, (size_t)64 * kDbgCtas * 8 * sizeof(long long), st));
          int li = 0;
          time_graph([&](int c) { CgParams q = params(v, C, c); q.dbg_time = DBG + (size_t)(li++ % launches) * kDbgCtas * 8; launch_cg(v, q, C, st, pdl); });
          CK(cudaMemcpy(hdbg.data(), DBG, hdbg.size() * sizeof(long long), cudaMemcpyDeviceToHost));
          double ph[7] = {0, 0, 0, 0, 0, 0, 0}, span = 0, gap = 0, period = 0;
          int cnt = 0;
          long long prev_end = 0, prev_t1 = 0;
          for (int L = launches / 2; L < launches; ++L) {
            const long long* d = hdbg.data() + (size_t)L * kDbgCtas * 8;
            long long t1min = 1ll << 62, t7max = 0;
            std::vector<std::vector<double>> vals(7);
            for (int b = 0; b < grid; ++b) {
              const long long* t = d + (size_t)b * 8;
              if (t[1] < t1min) t1min = t[1];
              if (t[7] > t7max) t7max = t[7];
              for (int k = 0; k < 7; ++k) vals[k].push_back((double)(t[k + 1] - t[k]));
            }
            if (L > launches / 2) {
              for (int k = 0; k < 7; ++k) { std::sort(vals[k].begin(), vals[k].end()); ph[k] += vals[k][vals[k].size() / 2]; }
              span += (double)(t7max - t1min);
              gap += (double)(t1min - prev_end);
              period += (double)(t1min - prev_t1);
              ++cnt;
            }
            prev_end = t7max; prev_t1 = t1min;
          }
          if (cnt)
            snprintf(tl, sizeof tl, ", \"timeline_us\": {\"period\": %.2f, \"span_wait_to_end\": %.2f, \"gap_end_to_next_wait\": %.2f, \"p50\": {\"wait\": %.2f, "
                     "\"list\": %.2f, \"gather\": %.2f, \"cta_reduce\": %.2f, \"bar1\": %.2f, \"dsmem_epi\": %.2f, \"bar2\": %.2f}}",
                     period / cnt / 1e3, span / cnt / 1e3, gap / cnt / 1e3, ph[0] / cnt / 1e3, ph[1] / cnt / 1e3, ph[2] / cnt / 1e3, ph[3] / cnt / 1e3,
                     ph[4] / cnt / 1e3, ph[5] / cnt / 1e3, ph[6] / cnt / 1e3);
        }
        char line[1100];
        snprintf(line, sizeof line,
                 "{\"shape\": \"%s\", \"Z\": %d, \"N\": %d, \"kernel\": \"cluster\", \"variant\": \"%s\", \"C\": %d, \"ctas\": %d, \"max_active_clusters\": %d, "
                 "\"us\": %.3f, \"gbs\": %.0f, \"nnz\": %d, \"bad\": %d, \"max_err\": %.4g, \"max_ref\": %.3g%s}",
                 sh.name, Z, N, v.name, C, p0.ntiles * C, max_clusters, us, alg / us / 1e3, nnz, bad, maxerr, maxref, tl);
        puts(line); fflush(stdout);
        if (out) { fprintf(out, "%s\n", line); fflush(out); }
      }
    }
    CK(cudaFree(W)); CK(cudaFree(X)); CK(cudaFree(NW)); CK(cudaFree(Y)); CK(cudaFree(R)); CK(cudaFree(SSQ)); CK(cudaFree(SSQO)); CK(cudaFree(YR));
  }
  if (out) fclose(out);
  return 0;
}
