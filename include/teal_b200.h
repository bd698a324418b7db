/*
 * teal_b200 -- C ABI of the B200-native activation-sparse decode path.
 *
 * This header is the drop-in boundary.  The reference (FasterDecoding/TEAL) has no C ABI: its
 * boundary is two torch.library ops backed by Triton.  Each entry point below names the
 * reference interface it replaces (paths relative to the reference repo root).  Everything is
 * plain pointers + sizes; no torch types.  All device pointers are CUDA device pointers on the
 * current device, all launches are asynchronous on `stream` (a cudaStream_t passed as void*),
 * nothing allocates or synchronises, so every call is CUDA-graph-capturable.
 *
 * Conventions
 *   - return value: 0 on success, negative teal_status otherwise; teal_last_error() gives the
 *     message for the calling thread.
 *   - "wt" weights are the reference's "column major" layout (gpt-fast/generate.py:296-317,
 *     asserted at kernels/sparse_gemv.py:106,210): logical [N, Z] with stride (1, ldw), i.e.
 *     storage [Z][ldw] fp16 with the N output columns of one input lane contiguous.
 *   - thresholds are the reference's Python floats; the library rounds them to fp16
 *     (round-to-nearest-even) so that the survivor set equals the reference's
 *     `x.abs().gt(threshold)` on an fp16 tensor (utils/utils.py:51-52).  A lane survives iff
 *     |x| > fp16(threshold); NaN lanes are also gathered so that NaN poisons the output exactly
 *     as `mask * x` does in the reference.
 *   - the teal_set_* knobs (config, PDL, probes) are THREAD-LOCAL launch context of the calling host
 *     thread; the compute entry points themselves keep no state besides the workspace.
 *   - workspace: teal_workspace_bytes() bytes of device memory, zeroed ONCE with
 *     teal_workspace_init(); kernels leave it zeroed again on exit.  One workspace must not be
 *     used by two streams at the same time.
 */
#ifndef TEAL_B200_H_
#define TEAL_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum teal_status {
  TEAL_OK = 0,
  TEAL_ERR_BAD_ARG = -1,     /* null pointer, negative size */
  TEAL_ERR_ALIGNMENT = -2,   /* pointer / leading dimension not 16-byte aligned, N or Z % 8 != 0 */
  TEAL_ERR_UNSUPPORTED = -3, /* shape outside compiled limits (Z > 65536, ...) */
  TEAL_ERR_WORKSPACE = -4,   /* workspace too small */
  TEAL_ERR_CUDA = -5         /* CUDA runtime error (message has the cudaError string) */
} teal_status;

#define TEAL_B200_ABI_VERSION 5
#define TEAL_SSQ_PARTS 64    /* partial sums of squares (one per producer column tile, unused tail zero) */
#define TEAL_SSQ_REPLICAS 32 /* identical copies of the partials: ~4000 consumer warps read them at the same instant, one copy
                              * would be a single hot L2 line (measured: 1.4-2 us of every RMSNorm-fused launch) */
#define TEAL_SSQ_FLOATS (TEAL_SSQ_PARTS * TEAL_SSQ_REPLICAS) /* floats in a sum-of-squares buffer: replica r at [r * PARTS] */

int teal_abi_version(void);
const char* teal_last_error(void);

/* Number of SMs the library will spread work over / compiled-for arch (100). */
int teal_device_sm_count(void);
int teal_compiled_arch(void);

/* Bytes of workspace needed for any call whose total number of output columns
 * (N, or 2*I for the fused gate/up call) is <= max_out_cols. */
size_t teal_workspace_bytes(int max_out_cols);
int teal_workspace_init(void* workspace, size_t workspace_bytes, void* stream);

/* y[n] = sum_{m : |x[m]| > fp16(threshold)} wt[m][n] * x[m]          (fp32 accumulate, one fp16 round)
 * Replaces torch.ops.teal.sparse_gemv's decode branch: splitk_sparse_gemv()
 * kernels/sparse_gemv.py:87-142 + splitk_sparse_gemv_kernel :45-83 (incl. its Y.zero_() pre-hook).
 * x: fp16[Z], wt: fp16[Z][ldw], y: fp16[N]. */
int teal_sparse_gemv_f16(const void* x, const void* wt, void* y, int Z, int N, int64_t ldw,
                         double threshold, void* workspace, size_t workspace_bytes, void* stream);

/* Fused QKV with one threshold per output region: columns [0, N - 2*kv_size) use threshold_q,
 * the next kv_size use threshold_k, the last kv_size threshold_v -- selected per column exactly.
 * Replaces torch.ops.teal.sparse_qkv_gemv's decode branch: qkv_gemv() kernels/sparse_gemv.py:196-237
 * + qkv_kernel :147-194. */
int teal_sparse_qkv_gemv_f16(const void* x, const void* wt, void* y, int Z, int N, int64_t ldw,
                             int kv_size, double threshold_q, double threshold_k,
                             double threshold_v, void* workspace, size_t workspace_bytes,
                             void* stream);

/* h[i] = fp16(silu(gate[i])) * up[i] with gate = sparse_gemv(x, w1t, threshold_gate) and
 * up = sparse_gemv(x, w3t, threshold_up), both rounded to fp16 first -- the first half of
 * _new_ffn_forward (gpt-fast/model.py:258-259): two gemv1 calls + the SiLU*mul, in one kernel.
 * x: fp16[Z], w1t/w3t: fp16[Z][ldw*], h: fp16[I]. */
int teal_sparse_gate_up_f16(const void* x, const void* w1t, const void* w3t, void* h, int Z, int I,
                            int64_t ldw1, int64_t ldw3, double threshold_gate,
                            double threshold_up, void* workspace, size_t workspace_bytes,
                            void* stream);

/* ---- fused variants used by the decode pipeline (one launch per projection group) ---------------
 * RMSNorm (gpt-fast/model.py:286-291: fp32 norm, fp16 cast, fp16 multiply by norm_w) of the fp16
 * residual stream x is computed in the kernel prologue, then thresholded; `xn_out` (nullable)
 * receives the normalised vector.  Replaces attention_norm + gemv1 (model.py:159,168) ...
 * Sum-of-squares hand-off: no CTA reads the whole vector for the norm.  The producer of x leaves
 * TEAL_SSQ_PARTS fp32 partial sums of squares, replicated TEAL_SSQ_REPLICAS times (`ssq_out` of
 * teal_embed_f16 / teal_sparse_gemv_residual_f16, a buffer of TEAL_SSQ_FLOATS floats: every replica
 * is fully rewritten on every call) and the consumer takes the buffer as `ssq_in` (each CTA reads
 * one replica); with ssq_in == NULL the library computes them with one extra tiny launch. */
int teal_norm_sparse_qkv_gemv_f16(const void* x, const void* norm_w, double eps, const float* ssq_in,
                                  const void* wt, void* y, void* xn_out, int Z, int N, int64_t ldw,
                                  int kv_size,
                                  double threshold_q, double threshold_k, double threshold_v,
                                  void* workspace, size_t workspace_bytes, void* stream);
/* ... ffn_norm + gemv1(gate) + gemv1(up) + SiLU*mul (model.py:161,259) ... */
int teal_norm_sparse_gate_up_f16(const void* x, const void* norm_w, double eps, const float* ssq_in,
                                 const void* w1t, const void* w3t, void* h, void* xn_out, int Z, int I,
                                 int64_t ldw1,
                                 int64_t ldw3, double threshold_gate, double threshold_up,
                                 void* workspace, size_t workspace_bytes, void* stream);
/* ... and gemv2 + residual add: y = resid + fp16(sparse_gemv(x, wt)) in fp16
 * (model.py:160-161,188,259).  y may alias resid. */
int teal_sparse_gemv_residual_f16(const void* x, const void* wt, const void* resid, void* y,
                                  float* ssq_out, int Z, int N, int64_t ldw, double threshold,
                                  void* workspace, size_t workspace_bytes, void* stream);

/* ---- the rest of one batch-1 decode step -----------------------------
 * All of them read the token / position from device memory so a captured CUDA graph replays for
 * every token.
 * x[dim] = tok_embeddings[*token]  (gpt-fast/model.py:137). */
int teal_embed_f16(const void* tok_embeddings, const int32_t* token, void* x, float* ssq_out, int dim,
                   void* stream);
/* RoPE(q, k) with the fp16 (cos, sin) table freqs_cis[max_seq][head_dim/2][2], KV-cache append at
 * *pos, attention of the one query token over positions [0, *pos], heads concatenated
 * (gpt-fast/model.py:170-186).  qkv: fp16 [n_head*hd | n_kv*hd | n_kv*hd]; caches fp16
 * [n_kv][max_seq][hd] (the reference's [1, H_kv, S, D], model.py:84); out: fp16 [n_head*hd].
 * *pos outside [0, max_seq) touches no memory and writes NaN (the caller must bound the position).
 * Under teal_set_pdl(1) the kernel requests *pos, its freqs_cis row and the cache rows [0, *pos) BEFORE it waits for the
 * launch in front of it (only qkv is read after the wait): they must have been final before that launch started -- true
 * in a decode step, where the position comes from the previous step's teal_advance and the launch in front is the QKV
 * projection. */
int teal_attn_decode_f16(const void* qkv, const void* freqs_cis, void* k_cache, void* v_cache,
                         const int32_t* pos, void* out, int n_head, int n_kv_head, int head_dim,
                         int max_seq, void* stream);
/* Dense row-major GEMV y[v] = sum_z w[v][z] * xn[z] for the never-sparsified lm_head
 * (gpt-fast/model.py:141-142); xn = RMSNorm(x) with norm_w when norm_w != NULL, else x.
 * argmax_packed (nullable, uint64[1], zero before the call) receives the packed arg-max. */
int teal_dense_gemv_f16(const void* x, const void* norm_w, double eps, const void* w, void* y,
                        void* argmax_packed, int Z, int V, int64_t ldw, void* stream);
/* Tensor-parallel path: x += part (the all-reduced o/down partial, fp16 add of gpt-fast/model.py:160-161
 * after the all-reduce of tp.py:120-121,139-140) and the sum-of-squares partials of the new x. */
int teal_residual_add_f16(void* x, const void* part, float* ssq_out, int dim, void* stream);
/* Fused exchange step of the tensor-parallel path: one-shot all-reduce of the row-parallel partials over
 * NVLink peer memory + fp16 residual add + sum of squares, one kernel (replaces the ncclAllReduce of
 * gpt-fast/tp.py:120-121,139-140 and the add of model.py:160-161).  peer_partials[r] / peer_flags[r] (HOST arrays of
 * `world` device pointers, peer-mapped, e.g. from torch symmetric memory) are rank r's fp16 partial buffer [dim]
 * and uint32 flag array [world] (zero-initialised); epoch_cell is a local device uint32 (zero-initialised) the
 * kernel increments.  Consecutive calls must alternate between two partial buffers. */
int teal_tp_allreduce_residual_f16(const void* const* peer_partials, void* const* peer_flags, int rank,
                                   int world, void* x, float* ssq_out, void* epoch_cell, int dim,
                                   void* stream);
/* Top-k / temperature sampling of the next token -- gpt-fast/generate.py:49-66 (`sample`, `logits_to_probs`,
 * `multinomial_sample_one_no_sync`): l = fp16(logits / max(T, 1e-5)); values below the top_k-th largest l -> -inf;
 * p = fp16(softmax_fp32(l)); token = argmax fp16(p / q), q ~ Exponential(1) in fp16.  Runs after teal_dense_gemv_f16
 * and replaces the greedy arg-max in `argmax_packed` (consumed by teal_advance).  top_k = 0 or >= V: no top-k filter.
 * noise_f16: optional fp16[V] Exponential(1) draws (tests: the reference's own torch noise => same token as the
 * reference); NULL = Philox4x32-10 keyed by (seed; vocabulary index, *pos).  probs_out: optional fp16[V] (the `probs`
 * the reference returns).  One CTA, four passes over the L2-resident logits, graph-capturable. */
int teal_sample_topk_f16(const void* logits, int V, double temperature, int top_k, const void* noise_f16, uint64_t seed,
                         const int32_t* pos, void* argmax_packed, void* probs_out, void* stream);
/* Row-parallel projection (o / down under tensor parallelism, gpt-fast/tp.py:117,130) with its exchange step
 * (tp.py:120-121,139-140) and the residual add (model.py:160-161) in ONE kernel: resid_x[N] += sum over ranks of the
 * rank's partial y = W_shard * mask(x_shard), in place, plus the per-tile sums of squares for the next RMSNorm.  The
 * finaliser of every column tile PUSHES the rank's fp16 partial of that tile into every rank's receive buffer as 8-byte
 * words {2 x fp16, epoch} (data and flag in one store: no fence, no remote read), polls its own buffer until all `world`
 * words of a column pair carry the epoch, and adds them in rank order (bit-identical replicas).
 * peer_recv[r]: rank r's receive buffer, world * N / 2 eight-byte words = world * N * 4 bytes (peer-mapped, e.g. torch
 * symmetric memory; zero-initialised; HOST array of `world` device pointers); epoch_tiles: local device uint32
 * [TEAL_SSQ_PARTS], zero-initialised.  Consecutive exchanges must alternate between two (recv, epochs) sets. */
int teal_sparse_gemv_tp_exchange_f16(const void* x, const float* gate_up_acc, const void* wt, void* resid_x, float* ssq_out,
                                     int Z, int N, int64_t ldw, double threshold, float* clear, int clear_n,
                                     void* const* peer_recv, void* epoch_tiles, int rank, int world, void* workspace,
                                     size_t workspace_bytes, void* stream);
   /* (x / gate_up_acc, clear / clear_n: as in teal_sparse_gemv_residual_ex_f16 below) */

/* ---- deferred epilogue (decode step) ------------------------------------------------------------------------------
 * Attention and the down projection consume a projection's output with no RMSNorm in between, so the producer's tile
 * finaliser (fence -> arrival counter -> last arrival loads the fp32 sums, rounds, stores fp16, re-zeroes) can be left
 * out: the producer's CTAs end with their fp32 `red.add` into a caller-owned buffer, and the CONSUMER kernel applies the
 * same single round-to-nearest (and, for gate / up, fp16(silu(fp16 g)) * fp16 u of gpt-fast/model.py:259) when it loads
 * its input -- bit-identical values, ~1 us less on the critical path of each of the two launches.  The buffers must be
 * zero before their producer runs; `clear` on a later launch of the step re-zeroes them once every reader is done (the
 * o projection clears both: it starts after attention, i.e. after the last reader of q | k | v, and long after the
 * previous layer's down projection).  No workspace: nothing is counted or finalised.
 * qkv_acc: fp32 [N]; gate_up_acc: fp32 [2 * I], gate at [0, I), up at [I, 2 I); both 16-byte aligned. */
int teal_norm_sparse_qkv_gemv_acc_f16(const void* x, const void* norm_w, double eps, const float* ssq_in, const void* wt,
                                      float* qkv_acc, int Z, int N, int64_t ldw, int kv_size, double threshold_q,
                                      double threshold_k, double threshold_v, void* stream);
int teal_norm_sparse_gate_up_acc_f16(const void* x, const void* norm_w, double eps, const float* ssq_in, const void* w1t,
                                     const void* w3t, float* gate_up_acc, int Z, int I, int64_t ldw1, int64_t ldw3,
                                     double threshold_gate, double threshold_up, void* stream);
/* teal_attn_decode_f16 reading q | k | v from the fp32 sums of teal_norm_sparse_qkv_gemv_acc_f16. */
int teal_attn_decode_acc_f16(const float* qkv_acc, const void* freqs_cis, void* k_cache, void* v_cache,
                             const int32_t* pos, void* out, int n_head, int n_kv_head, int head_dim, int max_seq,
                             void* stream);
/* teal_sparse_gemv_residual_f16 whose input is EITHER the fp16 vector x OR (x == NULL) h = silu(gate) * up of
 * gate_up_acc (Z = I); clear (nullable) / clear_n (multiple of 4): floats this launch zeroes after its producer is done
 * -- a buffer whose readers have all completed by then, never anything this launch itself reads (its CTAs clear their
 * slices while other CTAs are still loading their input). */
int teal_sparse_gemv_residual_ex_f16(const void* x, const float* gate_up_acc, const void* wt, const void* resid, void* y,
                                     float* ssq_out, int Z, int N, int64_t ldw, double threshold, float* clear, int clear_n,
                                     void* workspace, size_t workspace_bytes, void* stream);

/* Greedy step bookkeeping: *token = arg-max index, out_tokens[*pos] = *token, ++*pos, reset cell. */
int teal_advance(void* argmax_packed, int32_t* token, int32_t* pos, int32_t* out_tokens, int out_cap,
                 void* stream);

/* Survivor set of one threshold, produced by the same device code the GEMV kernels use
 * (parity hook: "surviving-index set bit-exact").  idx_out: int32[Z] (sorted ascending, first
 * *count_out entries valid; its upper quarter is used as scratch for the keep mask),
 * count_out: int32[1].  Both device pointers. */
int teal_survivors_f16(const void* x, int Z, double threshold, int32_t* idx_out,
                       int32_t* count_out, void* stream);

/* Independent slow cross-check of teal_sparse_gemv_f16 (one warp per 256 columns, no K split,
 * no workspace); used by tests to triage, never by the decode path. */
int teal_sparse_gemv_f16_simple(const void* x, const void* wt, void* y, int Z, int N, int64_t ldw,
                                double threshold, void* stream);

/* ---- launch context of the CALLING THREAD -------------------------------------------------------------------------
 * The four setters below configure how the calling host thread launches; the state is thread-local (no process-wide
 * state: two decoders driven from two threads never see each other's settings) and is read at launch (or capture) time.
 *
 * Kernel selection for benchmarks and parity tests: 0 = library default (variant and K chunk chosen by launch size);
 * otherwise config_id = (variant + 1) + 16 * (lanes_per_warp / 32), variant = 0 (256-column tiles, 2 warps: the
 * default) | 1 (256, 4 warps) | 2 (512-column tiles, 4 warps) | 3 (512, 2 warps) | 4 (256, 1 warp); lanes_per_warp =
 * input lanes one warp owns (multiple of 32, <= 512; 0 = chosen by launch size). */
int teal_set_config(int config_id);

/* Launch with programmatic dependent launch (the prologue overlaps the previous kernel's tail).  Off by default; the
 * decode pipeline switches it on around its graph capture. */
int teal_set_pdl(int enabled);

/* Measurement hook: when non-null (device uint64[2]) every sparse GEMV launch adds its gathered
 * weight bytes (nnz * N * 2) to [0] and the dense equivalent (Z * N * 2) to [1]. */
int teal_set_bytes_probe(void* device_u64x2);

/* Profiling hook: while set, sparse GEMV launch i (i < slots, in launch order, also inside a CUDA-graph capture) runs
 * the TIMED twin of the kernel whose CTAs write 16 int64 per CTA into slot i of device_buf (slot stride = 4096 CTAs x
 * 16 x 8 bytes): globaltimer ns at [0] CTA start, [1] producer done (griddepcontrol.wait returned), [2] survivor list
 * built, [4] gather done, [5] partials reduced + REDs issued, [6] arrival counted, [7] tile finalised (last arrival
 * only); [8] SM id, [9] survivors of warp 0, [10] last-arrival flag.  NULL switches it off.  tools/gemv_timeline.py
 * reads it. */
int teal_set_debug_timeline(void* device_buf, int slots);

#ifdef __cplusplus
}
#endif
#endif /* TEAL_B200_H_ */
