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
 *   - the teal_set_* knobs (config, PDL, probes) are process-global and not thread-safe: they exist for
 *     benchmarks and profiling; the compute entry points themselves keep no state besides the workspace.
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

#define TEAL_B200_ABI_VERSION 2
#define TEAL_SSQ_PARTS 64 /* floats in a sum-of-squares partials buffer */

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
 * TEAL_SSQ_PARTS fp32 partial sums of squares (`ssq_out` of teal_embed_f16 /
 * teal_sparse_gemv_residual_f16: it writes its tiles' partials, the unused tail of the buffer must
 * be zero) and the consumer takes them as `ssq_in`; with ssq_in == NULL the library computes them
 * with one extra tiny launch. */
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
 * [n_kv][max_seq][hd] (the reference's [1, H_kv, S, D], model.py:84); out: fp16 [n_head*hd]. */
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
/* Flag-chained MLP pair (opt-in; checked on the CPU emulation with the two launches running concurrently, not yet run
 * on hardware).  Same arithmetic as teal_norm_sparse_gate_up_f16 followed by teal_sparse_gemv_residual_f16 on the same
 * workspace and stream, but the second launch does not wait for the first grid to finish: the gate/up finalisers
 * release h tile by tile (CTAs run group-major so that the first tiles are final early) and every down-projection
 * CTA waits only for the h tiles that cover its K chunk.  Both calls must use the library's default dispatch, be
 * issued back to back (PDL enabled) with Z = dim, I = intermediate size, N = dim; ssq_in is required. */
int teal_norm_sparse_gate_up_chain_f16(const void* x, const void* norm_w, double eps, const float* ssq_in, const void* w1t,
                                       const void* w3t, void* h, int Z, int I, int64_t ldw1, int64_t ldw3, double threshold_gate,
                                       double threshold_up, int wait_all, void* workspace, size_t workspace_bytes, void* stream);
int teal_sparse_gemv_residual_chain_f16(const void* h, const void* wt, const void* resid, void* y, float* ssq_out, int I, int N,
                                        int64_t ldw, double threshold, int release_all_ctas, void* workspace, size_t workspace_bytes,
                                        void* stream);
/* Block-to-block hand-off of the chained step (the two RMSNorm points): the last block of a chain (o-proj, down) called with
 * release_all_ctas = teal_chain_consumer_ctas(...) of the NEXT launch counts its final tiles and releases the whole residual
 * stream + sum-of-squares partials; that next launch (gate/up, the next layer's QKV) called with wait_all = 1 skips
 * griddepcontrol.wait and waits for the release instead.  wait_all = 0 / release_all_ctas = 0: an ordinary grid-wide wait
 * (first layer after the embedding, last layer before the lm_head).  kind 0 = QKV (a = n_head, b = n_kv_head, c = head_dim,
 * thresholds select the tiling), kind 1 = gate/up (a = intermediate size). */
int teal_chain_consumer_ctas(int kind, int dim, int a, int b, int c, double threshold_q, double threshold_k, double threshold_v);

/* Flag-chained attention block (opt-in, same status as the MLP pair): teal_norm_sparse_qkv_gemv_f16 -> teal_attn_decode_f16
 * -> teal_sparse_gemv_residual_f16 (o-proj), issued back to back on one workspace and stream, without a grid-wide wait
 * between them: the QKV finalisers release their column tiles (k / v tiles first), the attention CTA of a head waits for the
 * tiles holding its q / k / v slices, the o-proj CTAs for the heads covering their K chunk.  Z = dim, N = (n_head + 2 n_kv)
 * head_dim; the three calls must agree on n_head / n_kv_head / head_dim and on the QKV thresholds (they select the tiling). */
int teal_norm_sparse_qkv_gemv_chain_f16(const void* x, const void* norm_w, double eps, const float* ssq_in, const void* wt, void* y,
                                        int Z, int N, int64_t ldw, int kv_size, double threshold_q, double threshold_k,
                                        double threshold_v, int n_head, int n_kv_head, int head_dim, int wait_all, void* workspace,
                                        size_t workspace_bytes, void* stream);
int teal_attn_decode_chain_f16(const void* qkv, const void* freqs_cis, void* k_cache, void* v_cache, const int32_t* pos, void* out,
                               int n_head, int n_kv_head, int head_dim, int max_seq, int dim, double threshold_q, double threshold_k,
                               double threshold_v, void* workspace, size_t workspace_bytes, void* stream);
int teal_sparse_o_proj_chain_f16(const void* att, const void* wt, const void* resid, void* y, float* ssq_out, int Z, int N, int64_t ldw,
                                 double threshold, int n_head, int n_kv_head, int head_dim, int release_all_ctas, void* workspace,
                                 size_t workspace_bytes, void* stream);

/* Row-parallel projection (o / down under tensor parallelism, gpt-fast/tp.py:117,130) with its exchange step
 * (tp.py:120-121,139-140) and the residual add (model.py:160-161) in ONE kernel: resid_x[N] += sum over ranks of the
 * rank's partial y = W_shard * mask(x_shard), in place, plus the per-tile sums of squares for the next RMSNorm.  The
 * finaliser of every column tile writes the rank's fp16 partial of that tile into peer_partials[rank], raises
 * peer_flags[q][rank * TEAL_SSQ_PARTS + tile] on every peer q, waits for the peers' flags of the same tile and adds the
 * partials in rank order (bit-identical replicas).  peer_partials[r]: rank r's fp16 [N] buffer, peer_flags[r]: rank r's
 * uint32 [world * TEAL_SSQ_PARTS] flag array (both peer-mapped, zero-initialised; HOST arrays of `world` device
 * pointers); epoch_tiles: local device uint32 [TEAL_SSQ_PARTS], zero-initialised.  Consecutive exchanges must alternate
 * between two (partials, flags, epochs) sets, like teal_tp_allreduce_residual_f16.  Opt-in: checked on the CPU
 * emulation with concurrent emulated ranks, not yet run on hardware. */
int teal_sparse_gemv_tp_exchange_f16(const void* x, const void* wt, void* resid_x, float* ssq_out, int Z, int N, int64_t ldw,
                                     double threshold, void* const* peer_partials, void* const* peer_flags, void* epoch_tiles,
                                     int rank, int world, void* workspace, size_t workspace_bytes, void* stream);
/* Greedy step bookkeeping: *token = arg-max index, out_tokens[*pos] = *token, ++*pos, reset cell. */
int teal_advance(void* argmax_packed, int32_t* token, int32_t* pos, int32_t* out_tokens, int out_cap,
                 void* stream);

/* ---- step program: the whole decode step as ONE persistent data-flow kernel ------------------------
 * The same step the calls above perform launch by launch (decode_one_token, gpt-fast/generate.py:73-77
 * over model.py:133-190,258-259), planned once into a table of phases and run by one kernel per token:
 * CTAs take (column tile, K chunk) items from a global ticket counter in dependency order and wait only
 * on the completion flags of the producer tiles they read -- the down projection starts on the first
 * finished tiles of gate/up, attention heads on their q/k/v tiles, the o projection on finished heads.
 * Only the two RMSNorm points of a layer remain full barriers.
 *
 *   teal_step_program_sizes  bytes of the program blob and of the device scratch buffer
 *   teal_step_program_plan   fills a HOST blob (device pointers embedded; no CUDA call, runs without a GPU);
 *                            the caller uploads it verbatim to device memory
 *   teal_step_program_run    one launch = one token; scratch must have been zeroed once before the first
 *                            run and whenever a program is re-planned into it
 *   teal_step_program_item   CPU-side introspection of one ticket (schedule verifier / tests)
 * Layout and thresholds as in the calls above; a negative threshold keeps every lane. */
typedef struct teal_layer_desc {
  const void* wqkv;      /* fp16 [dim][ld_qkv], columns [q | k | v]; NULL = no attention block */
  const void* wo;        /* fp16 [n_head*head_dim][ld_o] */
  const void* w1;        /* gate, fp16 [dim][ld_1]; NULL = no MLP block */
  const void* w3;        /* up,   fp16 [dim][ld_3] */
  const void* w2;        /* down, fp16 [inter][ld_2] */
  const void* attn_norm; /* fp16 [dim] */
  const void* ffn_norm;  /* fp16 [dim] */
  void* k_cache;         /* fp16 [n_kv_head][max_seq][head_dim] */
  void* v_cache;
  int64_t ld_qkv, ld_o, ld_1, ld_3, ld_2;
  double thr[7];         /* q, k, v, o, gate, up, down */
} teal_layer_desc;

typedef struct teal_step_desc {
  int32_t dim, n_head, n_kv_head, head_dim, inter, vocab, n_layer, max_seq;
  double eps;
  const void* tok_emb;     /* fp16 [vocab][dim]; NULL = x is preset by the caller and *ssq_init = sum(x^2) */
  const int32_t* token;    /* device int32[1] */
  const float* ssq_init;   /* device float[1], only read when tok_emb == NULL */
  const void* lm_head;     /* fp16 [vocab][ld_head] row-major; NULL = stop after the last layer */
  const void* final_norm;  /* fp16 [dim] */
  int64_t ld_head;
  const void* freqs_cis;   /* fp16 [max_seq][head_dim/2][2] */
  int32_t* pos;            /* device int32[1] */
  int32_t* out_tokens;     /* device int32[out_cap] */
  void* argmax_packed;     /* device uint64[1], zero */
  void* x;                 /* fp16 [dim] residual stream */
  void* qkv;               /* fp16 [(n_head + 2 n_kv_head) * head_dim] */
  void* att;               /* fp16 [n_head * head_dim] */
  void* h;                 /* fp16 [inter] */
  void* logits;            /* fp16 [vocab] */
  void* scratch;           /* device, teal_step_program_sizes() bytes, zeroed once */
  const teal_layer_desc* layers; /* HOST array [n_layer] */
  int32_t out_cap;
  /* tuning knobs, 0 = library default */
  int32_t rows_small;      /* K chunk of the 256-column-tile GEMVs (default 64) */
  int32_t rows_large;      /* K chunk of the 512-column-tile GEMVs (default 128) */
  int32_t group_qkv;       /* column tiles per order group of the QKV phase (default 4) */
  int32_t group_gate_up;   /* column tiles per order group of the gate/up phase (even; default 8) */
  int32_t head_rows;       /* vocabulary rows per lm_head item (default 32) */
  int32_t wide_melems;     /* projections with >= this many million weights use 512-column tiles (default 40) */
  int32_t ctas_per_sm;     /* cap on resident CTAs per SM (0 = default: occupancy) */
  int32_t poll_ns;         /* back-off between two polls of a completion flag, ns (default 32; negative = none) */
  int32_t opts;            /* bit 0: a CTA that has seen a whole-phase dependency satisfied does not poll it again;
                            * bit 1: the RMSNorm scale is computed once per (CTA, phase) instead of per warp and item */
  void* debug_times;       /* profiling: device uint64[4 * phases] = per phase {first ticket taken, first item ready
                            * (dependencies seen), last item finished} in globaltimer ns (min, min, max; the caller
                            * initialises {~0, ~0, 0, 0}) and the summed busy ns of its items; NULL = off */
  int32_t warps;           /* warps per CTA: 4 (default, 7 CTAs per SM) or 2 (14 CTAs per SM) */
  int32_t stagger_ns;      /* CTAs released by a whole-phase wait start hash(cta) in [0,8) x stagger_ns apart (default 0) */
  int32_t rows_kind[4];    /* K chunk per projection kind {qkv, o, gate/up, down}; 0 = rows_small / rows_large */
  int32_t tiles_per_item;  /* experimental: an item sweeps this many column tiles of one K chunk (1..4, default 1): x is
                            * loaded once and the first rows of the next tile are in flight while a tile is reduced */
  int32_t reserved2;
} teal_step_desc;

typedef struct teal_item_info {
  int32_t phase, kind, layer, item;       /* kind: 0 embed, 1 gemv, 2 attention, 3 lm_head, 4 advance */
  int32_t ctile, kchunk, k_begin, k_end;  /* gemv items */
  int32_t pub_index;                      /* publication (tile / head / item) this item contributes to */
  int32_t pub_flag;                       /* absolute completion-flag index of that publication, -1 = none */
  int32_t pub_cell;                       /* done cell it increments */
  int32_t arrivals;                       /* items that must arrive before the publication completes */
  int32_t dep_kind, dep_cell, dep_per_epoch; /* 0 none, 1 all publications of a phase, 2/3 flag list */
  int32_t n_deps;
  int32_t deps[32];
  int32_t n_pubs;                         /* publications this item contributes to (column tiles it sweeps) */
  int32_t pub_indices[4];
  int32_t pub_flags[4];
  int32_t ctiles[4];
} teal_item_info;

int teal_step_program_sizes(const teal_step_desc* desc, size_t* program_bytes, size_t* scratch_bytes);
int teal_step_program_plan(const teal_step_desc* desc, void* host_program, size_t program_bytes);
int teal_step_program_run(const void* host_program, const void* device_program, void* stream);
int teal_step_program_item(const void* host_program, int ticket, teal_item_info* out);
/* number of tickets (items) of one step, or a negative teal_status */
int teal_step_program_items(const void* host_program);

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

/* Kernel selection for benchmarks and parity tests: 0 = library default (rowchunk_gemv: many small CTAs, variant
 * chosen by launch size); 30+v = rowchunk variant v (v >= 12: generation 2, v >= 20: warp-wise); 100+v = variant v without wave
 * balancing; 1..6 = TMA-bulk-ring kernel; 10+T = gather_gemv (cp.async rings) with T column tiles per matrix. */
int teal_set_config(int config_id);

/* Opt-in experimental knobs of the default dispatch (process-wide, like teal_set_config).
 *   TEAL_TUNE_GENERATION  1 (default, hardware-validated) | 2 (rowchunk_gemv2_kernel: shared-memory survivor list,
 *                         rolling load ring, FHFMA) | 3 (the same with warp-wise completion: every warp reduces and
 *                         arrives on its own, no __syncthreads); 2 and 3 pass the CPU emulation of tests/emu and are
 *                         SASS-checked, but have not run on hardware yet
 *   TEAL_TUNE_SPEC_ROWS   generation 2: rows per warp the first-wave CTAs pull into L2 before their producer is done
 *                         (0 = off, default)
 *   TEAL_TUNE_SPEC_MODE   1 = prefetch.global.L2 per 128-byte line (default), 2 = cp.async.bulk.prefetch.L2 per row
 *   TEAL_TUNE_FHFMA       generation 2: 1 = FHFMA arithmetic (default), 0 = conversions + FFMA2 like generation 1
 *   TEAL_TUNE_ROWS_WIDE / TEAL_TUNE_ROWS_NARROW   input lanes (K chunk) per CTA of the wide-tile (>= 40 M-element
 *                         launches) / narrow-tile default variant; 0 = the variant's own size with wave balancing
 *   TEAL_TUNE_V1_OPTS     generation 1 with ONE change at a time (bit mask): 1 = x / norm weights requested together with
 *                         the sum-of-squares partials, 2 = FHFMA arithmetic, 4 = 32-bit row offsets
 *   TEAL_TUNE_WIDE_KELEMS  launches of at least this many thousand weight elements use the wide-tile (512-column) variant
 *                         (0 = the measured default, 40 000) */
enum { TEAL_TUNE_GENERATION = 0, TEAL_TUNE_SPEC_ROWS = 1, TEAL_TUNE_SPEC_MODE = 2, TEAL_TUNE_FHFMA = 3, TEAL_TUNE_ROWS_WIDE = 4,
       TEAL_TUNE_ROWS_NARROW = 5, TEAL_TUNE_V1_OPTS = 6, TEAL_TUNE_WIDE_KELEMS = 7, TEAL_TUNE_COUNT = 8 };
int teal_set_tuning(int knob, int value);

/* Launch the GEMV kernels with programmatic dependent launch (prologue overlaps the previous
 * kernel's tail).  Off by default; the decode pipeline switches it on. */
int teal_set_pdl(int enabled);

/* Measurement hook: when non-null (device uint64[2]) every sparse GEMV launch adds its gathered
 * weight bytes (nnz * N * 2) to [0] and the dense equivalent (Z * N * 2) to [1]. */
int teal_set_bytes_probe(void* device_u64x2);

/* Profiling hook: when non-null, every CTA of the GEMV kernels writes 16 int64 phase timestamps
 * (clock64 / globaltimer) to device_buf[blockIdx.x * 16 ...]; pass NULL to switch it off. */
int teal_set_debug_timestamps(void* device_buf);

/* Profiling hook of the production row-chunk kernel: while set, launch i (i < slots, in launch order, also inside a
 * CUDA-graph capture) runs a TIMED twin of the default kernel whose CTAs write 16 int64 per CTA into slot i of
 * device_buf (slot stride = 4096 CTAs x 16 x 8 bytes): globaltimer ns at [0] CTA start, [1] producer done
 * (griddepcontrol.wait returned), [2] first survivor set known, [3] first rows consumed, [4] gather done, [5] partials
 * reduced + REDs issued, [6] arrival counted, [7] tile finalised (last arrival only); [8] SM id, [9] survivors of
 * warp 0, [10] last-arrival flag.  NULL switches it off.  tools/rowchunk_timeline.py reads it. */
int teal_set_debug_timeline(void* device_buf, int slots);

#ifdef __cplusplus
}
#endif
#endif /* TEAL_B200_H_ */
