// extern "C" boundary of libteal_b200 (see include/teal_b200.h for the contract and the
// reference interfaces each entry point replaces).
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

#include "common.cuh"

namespace teal {

static thread_local char g_err[512] = "";

int set_error(int code, const char* msg) {
  snprintf(g_err, sizeof(g_err), "%s", msg);
  return code;
}
int set_cuda_error(cudaError_t e, const char* where) {
  snprintf(g_err, sizeof(g_err), "%s: %s", where, cudaGetErrorString(e));
  return TEAL_ERR_CUDA;
}

const DeviceInfo& device_info() {
  static thread_local int cached_dev = -1;
  static thread_local DeviceInfo di = {cudaSuccess, 0, 0};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) { di.err = e; return di; }
  if (dev != cached_dev) {
    di.err = cudaDeviceGetAttribute(&di.sm_count, cudaDevAttrMultiProcessorCount, dev);
    if (di.err == cudaSuccess)
      di.err = cudaDeviceGetAttribute(&di.smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (di.err == cudaSuccess) cached_dev = dev;
  }
  return di;
}

// Round-to-nearest-even to fp16, like torch demoting a Python scalar to an fp16 tensor's dtype.
float round_threshold_fp16(double t) {
  if (t != t) return t;  // NaN threshold: every compare false -> everything "survives" as in `!(a <= t)`
  __half h = __double2half(t);
  return __half2float(h);
}

static int fill_common(GemvParams& p, const void* x, void* y, int Z) {
  memset(&p, 0, sizeof(p));
  if (!x) return set_error(TEAL_ERR_BAD_ARG, "x is null");
  p.x = reinterpret_cast<const __half*>(x);
  p.y = reinterpret_cast<__half*>(y);
  p.Z = Z;
  return TEAL_OK;
}

LaunchCtx& launch_ctx() {
  static thread_local LaunchCtx ctx;
  return ctx;
}
int gemv_num_variants();

}  // namespace teal

using namespace teal;

extern "C" {

int teal_abi_version(void) { return TEAL_B200_ABI_VERSION; }
const char* teal_last_error(void) { return g_err; }
int teal_compiled_arch(void) { return 100; }
int teal_device_sm_count(void) {
  const DeviceInfo& di = device_info();
  if (di.err != cudaSuccess) return set_cuda_error(di.err, "device query");
  return di.sm_count;
}
size_t teal_workspace_bytes(int max_out_cols) { return workspace_bytes(max_out_cols < 8 ? 8 : max_out_cols); }

int teal_workspace_init(void* workspace, size_t workspace_bytes_, void* stream) {
  if (!workspace) return set_error(TEAL_ERR_BAD_ARG, "workspace is null");
  cudaError_t e = cudaMemsetAsync(workspace, 0, workspace_bytes_, reinterpret_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return set_cuda_error(e, "cudaMemsetAsync");
  return TEAL_OK;
}

// ---- launch context of the calling thread (thread-local: no process-wide state) ------------------------------------
int teal_set_config(int config_id) {
  LaunchCtx& c = launch_ctx();
  if (config_id == 0) { c.variant = -1; c.lanes_per_warp = 0; return TEAL_OK; }
  const int v = (config_id & 15) - 1, lanes = (config_id >> 4) * 32;
  if (config_id < 0 || v < 0 || v >= gemv_num_variants() || lanes > kListCap) return set_error(TEAL_ERR_BAD_ARG, "config id out of range");
  c.variant = v;
  c.lanes_per_warp = lanes;
  return TEAL_OK;
}
int teal_set_bytes_probe(void* device_u64x2) {
  launch_ctx().probe = reinterpret_cast<unsigned long long*>(device_u64x2);
  return TEAL_OK;
}
int teal_set_pdl(int enabled) {
  launch_ctx().pdl = enabled != 0;
  return TEAL_OK;
}
int teal_set_debug_timeline(void* device_buf, int slots) {
  if (device_buf && slots <= 0) return set_error(TEAL_ERR_BAD_ARG, "timeline: slots must be positive");
  LaunchCtx& c = launch_ctx();
  c.timeline = reinterpret_cast<long long*>(device_buf);
  c.timeline_slots = device_buf ? slots : 0;
  c.timeline_next = 0;
  return TEAL_OK;
}

// q | k | v column regions of the fused wqkv matrix (kernels/sparse_gemv.py:212): one region and one survivor list
// when the three thresholds agree
static void qkv_regions(GemvParams& p, const void* wt, int N, int64_t ldw, int kv_size, double threshold_q, double threshold_k,
                        double threshold_v) {
  const int n_q = N - 2 * kv_size;
  const float tq = round_threshold_fp16(threshold_q), tk = round_threshold_fp16(threshold_k),
              tv = round_threshold_fp16(threshold_v);
  const __half* w = reinterpret_cast<const __half*>(wt);
  if (tq == tk && tk == tv) {
    p.nseg = 1;
    p.seg[0] = GemvSeg{w, ldw, N, 0, tq};
    return;
  }
  const int cols[3] = {n_q, kv_size, kv_size};
  const float thr[3] = {tq, tk, tv};
  int c0 = 0;
  for (int r = 0; r < 3; ++r) {
    if (cols[r] == 0) continue;
    p.seg[p.nseg++] = GemvSeg{w + c0, ldw, cols[r], c0, thr[r]};
    c0 += cols[r];
  }
}

int teal_sparse_gemv_f16(const void* x, const void* wt, void* y, int Z, int N, int64_t ldw, double threshold,
                         void* workspace, size_t workspace_bytes_, void* stream) {
  GemvParams p;
  if (int rc = fill_common(p, x, y, Z)) return rc;
  if (!wt || !y || N <= 0 || ldw < N) return set_error(TEAL_ERR_BAD_ARG, "bad wt / y / N / ldw");
  p.mode = kModePlain;
  p.nseg = 1;
  p.seg[0].w = reinterpret_cast<const __half*>(wt);
  p.seg[0].ldw = ldw;
  p.seg[0].ncols = N;
  p.seg[0].out_col0 = 0;
  p.seg[0].thr = round_threshold_fp16(threshold);
  return launch_sparse_gemv(p, workspace, workspace_bytes_, N, reinterpret_cast<cudaStream_t>(stream));
}

int teal_sparse_qkv_gemv_f16(const void* x, const void* wt, void* y, int Z, int N, int64_t ldw, int kv_size,
                             double threshold_q, double threshold_k, double threshold_v, void* workspace,
                             size_t workspace_bytes_, void* stream) {
  GemvParams p;
  if (int rc = fill_common(p, x, y, Z)) return rc;
  if (!wt || !y || N <= 0 || ldw < N || kv_size < 0 || 2 * (int64_t)kv_size > N)
    return set_error(TEAL_ERR_BAD_ARG, "bad wt / y / N / ldw / kv_size");
  p.mode = kModePlain;
  qkv_regions(p, wt, N, ldw, kv_size, threshold_q, threshold_k, threshold_v);
  return launch_sparse_gemv(p, workspace, workspace_bytes_, N, reinterpret_cast<cudaStream_t>(stream));
}

int teal_sparse_gate_up_f16(const void* x, const void* w1t, const void* w3t, void* h, int Z, int I, int64_t ldw1,
                            int64_t ldw3, double threshold_gate, double threshold_up, void* workspace,
                            size_t workspace_bytes_, void* stream) {
  GemvParams p;
  if (int rc = fill_common(p, x, h, Z)) return rc;
  if (!w1t || !w3t || !h || I <= 0 || ldw1 < I || ldw3 < I) return set_error(TEAL_ERR_BAD_ARG, "bad w1t / w3t / h / I / ldw");
  p.mode = kModePairSilu;
  p.nseg = 2;
  p.seg[0] = GemvSeg{reinterpret_cast<const __half*>(w1t), ldw1, I, 0, round_threshold_fp16(threshold_gate)};
  p.seg[1] = GemvSeg{reinterpret_cast<const __half*>(w3t), ldw3, I, I, round_threshold_fp16(threshold_up)};
  return launch_sparse_gemv(p, workspace, workspace_bytes_, 2 * I, reinterpret_cast<cudaStream_t>(stream));
}

// ---- fused variants used by the decode pipeline --------------------------------------------------
// RMSNorm input: use the producer's sum-of-squares partials when given, else compute them here (one tiny launch)
static int norm_input(GemvParams& p, const void* norm_w, double eps, const float* ssq_in, void* workspace,
                      size_t ws_bytes, cudaStream_t stream) {
  p.in_mode = kInRmsNorm;
  p.norm_w = reinterpret_cast<const __half*>(norm_w);
  p.eps = (float)eps;
  p.n_ssq_in = kSsqParts;
  if (ssq_in) { p.ssq_in = ssq_in; return TEAL_OK; }
  if (!workspace || ws_bytes < workspace_bytes(8)) return set_error(TEAL_ERR_WORKSPACE, "workspace too small");
  float* scratch = workspace_ssq_scratch(workspace, ws_bytes);
  p.ssq_in = scratch;
  return launch_ssq(p.x, p.Z, scratch, kSsqParts, stream);
}

int teal_norm_sparse_qkv_gemv_f16(const void* x, const void* norm_w, double eps, const float* ssq_in, const void* wt,
                                  void* y, void* xn_out, int Z, int N, int64_t ldw, int kv_size, double threshold_q,
                                  double threshold_k, double threshold_v, void* workspace, size_t workspace_bytes_,
                                  void* stream) {
  GemvParams p;
  if (int rc = fill_common(p, x, y, Z)) return rc;
  if (!wt || !y || !norm_w || N <= 0 || ldw < N || kv_size < 0 || 2 * (int64_t)kv_size > N)
    return set_error(TEAL_ERR_BAD_ARG, "bad wt / y / norm_w / N / ldw / kv_size");
  p.mode = kModePlain;
  if (int rc = norm_input(p, norm_w, eps, ssq_in, workspace, workspace_bytes_, reinterpret_cast<cudaStream_t>(stream))) return rc;
  p.x_out = reinterpret_cast<__half*>(xn_out);
  qkv_regions(p, wt, N, ldw, kv_size, threshold_q, threshold_k, threshold_v);
  return launch_sparse_gemv(p, workspace, workspace_bytes_, N, reinterpret_cast<cudaStream_t>(stream));
}

int teal_norm_sparse_gate_up_f16(const void* x, const void* norm_w, double eps, const float* ssq_in, const void* w1t,
                                 const void* w3t, void* h, void* xn_out, int Z, int I, int64_t ldw1, int64_t ldw3,
                                 double threshold_gate, double threshold_up, void* workspace,
                                 size_t workspace_bytes_, void* stream) {
  GemvParams p;
  if (int rc = fill_common(p, x, h, Z)) return rc;
  if (!w1t || !w3t || !h || !norm_w || I <= 0 || ldw1 < I || ldw3 < I)
    return set_error(TEAL_ERR_BAD_ARG, "bad w1t / w3t / h / norm_w / I / ldw");
  p.mode = kModePairSilu;
  if (int rc = norm_input(p, norm_w, eps, ssq_in, workspace, workspace_bytes_, reinterpret_cast<cudaStream_t>(stream))) return rc;
  p.x_out = reinterpret_cast<__half*>(xn_out);
  p.nseg = 2;
  p.seg[0] = GemvSeg{reinterpret_cast<const __half*>(w1t), ldw1, I, 0, round_threshold_fp16(threshold_gate)};
  p.seg[1] = GemvSeg{reinterpret_cast<const __half*>(w3t), ldw3, I, I, round_threshold_fp16(threshold_up)};
  return launch_sparse_gemv(p, workspace, workspace_bytes_, 2 * I, reinterpret_cast<cudaStream_t>(stream));
}

// x is either the fp16 vector `x` or -- the down projection behind a deferred gate/up launch -- silu(gate) * up of the
// fp32 sums in `gate_up_acc` (gate at [0, Z), up at [Z, 2Z)); exactly one of the two is given
static int residual_input(GemvParams& p, const void* x, const float* gate_up_acc, void* y, int Z, float* clear, int clear_n) {
  memset(&p, 0, sizeof(p));
  if ((x == nullptr) == (gate_up_acc == nullptr)) return set_error(TEAL_ERR_BAD_ARG, "exactly one of x / gate_up_acc");
  p.x = reinterpret_cast<const __half*>(x);
  p.y = reinterpret_cast<__half*>(y);
  p.Z = Z;
  if (gate_up_acc) {
    p.in_mode = kInGateUpAcc;
    p.in_acc = gate_up_acc;
    p.in_acc_stride = Z;
  }
  p.clear = clear;
  p.clear_n = clear_n;
  return TEAL_OK;
}

int teal_sparse_gemv_residual_f16(const void* x, const void* wt, const void* resid, void* y, float* ssq_out, int Z,
                                  int N, int64_t ldw, double threshold, void* workspace, size_t workspace_bytes_,
                                  void* stream) {
  return teal_sparse_gemv_residual_ex_f16(x, nullptr, wt, resid, y, ssq_out, Z, N, ldw, threshold, nullptr, 0, workspace,
                                          workspace_bytes_, stream);
}

int teal_sparse_gemv_residual_ex_f16(const void* x, const float* gate_up_acc, const void* wt, const void* resid, void* y,
                                     float* ssq_out, int Z, int N, int64_t ldw, double threshold, float* clear, int clear_n,
                                     void* workspace, size_t workspace_bytes_, void* stream) {
  GemvParams p;
  if (int rc = residual_input(p, x, gate_up_acc, y, Z, clear, clear_n)) return rc;
  if (!wt || !y || !resid || N <= 0 || ldw < N) return set_error(TEAL_ERR_BAD_ARG, "bad wt / y / resid / N / ldw");
  p.mode = kModePlain;
  p.out_mode = kOutAddResidual;
  p.resid = reinterpret_cast<const __half*>(resid);
  p.ssq_out = ssq_out;
  p.nseg = 1;
  p.seg[0] = GemvSeg{reinterpret_cast<const __half*>(wt), ldw, N, 0, round_threshold_fp16(threshold)};
  return launch_sparse_gemv(p, workspace, workspace_bytes_, N, reinterpret_cast<cudaStream_t>(stream));
}

int teal_sparse_gemv_tp_exchange_f16(const void* x, const float* gate_up_acc, const void* wt, void* resid_x, float* ssq_out,
                                     int Z, int N, int64_t ldw, double threshold, float* clear, int clear_n,
                                     void* const* peer_recv, void* epoch_tiles, int rank, int world, void* workspace,
                                     size_t workspace_bytes_, void* stream) {
  GemvParams p;
  if (int rc = residual_input(p, x, gate_up_acc, resid_x, Z, clear, clear_n)) return rc;
  if (!wt || !resid_x || !peer_recv || !epoch_tiles || N <= 0 || ldw < N || world < 1 || world > kMaxTpWorld)
    return set_error(TEAL_ERR_BAD_ARG, "bad wt / resid_x / peers / N / ldw / world");
  TpFuse tp;
  memset(&tp, 0, sizeof(tp));
  for (int r = 0; r < world; ++r) tp.recv[r] = reinterpret_cast<uint2*>(peer_recv[r]);
  tp.epoch = reinterpret_cast<unsigned*>(epoch_tiles);
  tp.rank = rank;
  tp.world = world;
  tp.n = N;
  p.mode = kModePlain;
  p.out_mode = kOutAddResidual;
  p.resid = reinterpret_cast<const __half*>(resid_x);
  p.ssq_out = ssq_out;
  p.tp = &tp;
  p.nseg = 1;
  p.seg[0] = GemvSeg{reinterpret_cast<const __half*>(wt), ldw, N, 0, round_threshold_fp16(threshold)};
  return launch_sparse_gemv(p, workspace, workspace_bytes_, N, reinterpret_cast<cudaStream_t>(stream));
}

// ---- deferred epilogue: the producers of the two consumers that have no RMSNorm in front (attention, down) ----------
int teal_norm_sparse_qkv_gemv_acc_f16(const void* x, const void* norm_w, double eps, const float* ssq_in, const void* wt,
                                      float* qkv_acc, int Z, int N, int64_t ldw, int kv_size, double threshold_q,
                                      double threshold_k, double threshold_v, void* stream) {
  GemvParams p;
  if (int rc = fill_common(p, x, nullptr, Z)) return rc;
  if (!wt || !qkv_acc || !norm_w || !ssq_in || N <= 0 || ldw < N || kv_size < 0 || 2 * (int64_t)kv_size > N)
    return set_error(TEAL_ERR_BAD_ARG, "bad wt / qkv_acc / norm_w / ssq_in / N / ldw / kv_size");
  p.mode = kModePlain;
  p.out_mode = kOutDeferred;
  p.raw_out = qkv_acc;
  if (int rc = norm_input(p, norm_w, eps, ssq_in, nullptr, 0, reinterpret_cast<cudaStream_t>(stream))) return rc;
  qkv_regions(p, wt, N, ldw, kv_size, threshold_q, threshold_k, threshold_v);
  return launch_sparse_gemv(p, nullptr, 0, N, reinterpret_cast<cudaStream_t>(stream));
}

int teal_norm_sparse_gate_up_acc_f16(const void* x, const void* norm_w, double eps, const float* ssq_in, const void* w1t,
                                     const void* w3t, float* gate_up_acc, int Z, int I, int64_t ldw1, int64_t ldw3,
                                     double threshold_gate, double threshold_up, void* stream) {
  GemvParams p;
  if (int rc = fill_common(p, x, nullptr, Z)) return rc;
  if (!w1t || !w3t || !gate_up_acc || !norm_w || !ssq_in || I <= 0 || ldw1 < I || ldw3 < I)
    return set_error(TEAL_ERR_BAD_ARG, "bad w1t / w3t / gate_up_acc / norm_w / ssq_in / I / ldw");
  p.mode = kModePairSilu;
  p.out_mode = kOutDeferred;
  p.raw_out = gate_up_acc;
  if (int rc = norm_input(p, norm_w, eps, ssq_in, nullptr, 0, reinterpret_cast<cudaStream_t>(stream))) return rc;
  p.nseg = 2;
  p.seg[0] = GemvSeg{reinterpret_cast<const __half*>(w1t), ldw1, I, 0, round_threshold_fp16(threshold_gate)};
  p.seg[1] = GemvSeg{reinterpret_cast<const __half*>(w3t), ldw3, I, I, round_threshold_fp16(threshold_up)};
  return launch_sparse_gemv(p, nullptr, 0, 2 * I, reinterpret_cast<cudaStream_t>(stream));
}

// ---- the rest of the decode step -------------------------------------------------------------------
int teal_embed_f16(const void* tok_embeddings, const int32_t* token, void* x, float* ssq_out, int dim, void* stream) {
  return launch_embed(reinterpret_cast<const __half*>(tok_embeddings), token, reinterpret_cast<__half*>(x), ssq_out, dim,
                      reinterpret_cast<cudaStream_t>(stream));
}
int teal_attn_decode_f16(const void* qkv, const void* freqs_cis, void* k_cache, void* v_cache, const int32_t* pos,
                         void* out, int n_head, int n_kv_head, int head_dim, int max_seq, void* stream) {
  return launch_attn_decode(reinterpret_cast<const __half*>(qkv), nullptr, reinterpret_cast<const __half*>(freqs_cis),
                            reinterpret_cast<__half*>(k_cache), reinterpret_cast<__half*>(v_cache), pos,
                            reinterpret_cast<__half*>(out), n_head, n_kv_head, head_dim, max_seq,
                            reinterpret_cast<cudaStream_t>(stream));
}
int teal_attn_decode_acc_f16(const float* qkv_acc, const void* freqs_cis, void* k_cache, void* v_cache, const int32_t* pos,
                             void* out, int n_head, int n_kv_head, int head_dim, int max_seq, void* stream) {
  if (!qkv_acc) return set_error(TEAL_ERR_BAD_ARG, "attn_decode_acc: qkv_acc is null");
  return launch_attn_decode(nullptr, qkv_acc, reinterpret_cast<const __half*>(freqs_cis),
                            reinterpret_cast<__half*>(k_cache), reinterpret_cast<__half*>(v_cache), pos,
                            reinterpret_cast<__half*>(out), n_head, n_kv_head, head_dim, max_seq,
                            reinterpret_cast<cudaStream_t>(stream));
}
int teal_dense_gemv_f16(const void* x, const void* norm_w, double eps, const void* w, void* y, void* argmax_packed,
                        int Z, int V, int64_t ldw, void* stream) {
  return launch_dense_rowmajor(reinterpret_cast<const __half*>(x), reinterpret_cast<const __half*>(norm_w), (float)eps,
                               reinterpret_cast<const __half*>(w), reinterpret_cast<__half*>(y),
                               reinterpret_cast<unsigned long long*>(argmax_packed), Z, V, ldw,
                               reinterpret_cast<cudaStream_t>(stream));
}
int teal_residual_add_f16(void* x, const void* part, float* ssq_out, int dim, void* stream) {
  return launch_residual_add(reinterpret_cast<__half*>(x), reinterpret_cast<const __half*>(part), ssq_out, dim,
                             reinterpret_cast<cudaStream_t>(stream));
}
int teal_tp_allreduce_residual_f16(const void* const* peer_partials, void* const* peer_flags, int rank, int world,
                                   void* x, float* ssq_out, void* epoch_cell, int dim, void* stream) {
  return launch_tp_allreduce_residual(peer_partials, peer_flags, rank, world, reinterpret_cast<__half*>(x), ssq_out,
                                      reinterpret_cast<unsigned*>(epoch_cell), dim, reinterpret_cast<cudaStream_t>(stream));
}
int teal_sample_topk_f16(const void* logits, int V, double temperature, int top_k, const void* noise_f16, uint64_t seed,
                         const int32_t* pos, void* argmax_packed, void* probs_out, void* stream) {
  return launch_sample_topk(reinterpret_cast<const __half*>(logits), V, (float)temperature, top_k,
                            reinterpret_cast<const __half*>(noise_f16), seed, pos,
                            reinterpret_cast<unsigned long long*>(argmax_packed), reinterpret_cast<__half*>(probs_out),
                            reinterpret_cast<cudaStream_t>(stream));
}
int teal_advance(void* argmax_packed, int32_t* token, int32_t* pos, int32_t* out_tokens, int out_cap, void* stream) {
  return launch_advance(reinterpret_cast<unsigned long long*>(argmax_packed), token, pos, out_tokens, out_cap,
                        reinterpret_cast<cudaStream_t>(stream));
}

int teal_survivors_f16(const void* x, int Z, double threshold, int32_t* idx_out, int32_t* count_out, void* stream) {
  GemvParams p;
  if (int rc = fill_common(p, x, nullptr, Z)) return rc;
  if (!idx_out || !count_out) return set_error(TEAL_ERR_BAD_ARG, "null output");
  p.mode = kModeSurvivors;
  p.nseg = 0;
  p.seg[0].thr = round_threshold_fp16(threshold);
  // the production kernel writes its keep mask into the upper quarter of idx_out (Z bytes of scratch); a
  // single-CTA pass then turns it into the sorted index list (it never overtakes the mask bytes it still needs)
  p.keep_out = reinterpret_cast<unsigned char*>(idx_out) + (size_t)3 * Z;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (int rc = launch_sparse_gemv(p, nullptr, 0, 0, s)) return rc;
  return launch_mask_to_indices(p.keep_out, Z, idx_out, count_out, s);
}

int teal_sparse_gemv_f16_simple(const void* x, const void* wt, void* y, int Z, int N, int64_t ldw, double threshold,
                                void* stream) {
  if (!x || !wt || !y || Z <= 0 || N <= 0 || ldw < N) return set_error(TEAL_ERR_BAD_ARG, "bad argument");
  if ((ldw & 7) || (reinterpret_cast<uintptr_t>(wt) & 15)) return set_error(TEAL_ERR_ALIGNMENT, "wt alignment");
  return launch_simple(reinterpret_cast<const __half*>(x), reinterpret_cast<const __half*>(wt),
                       reinterpret_cast<__half*>(y), Z, N, ldw, round_threshold_fp16(threshold),
                       reinterpret_cast<cudaStream_t>(stream));
}

}  // extern "C"
