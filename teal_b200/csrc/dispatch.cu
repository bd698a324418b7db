// Host-side kernel selection and process-wide knobs of libteal_b200 (no device code): which of the three GEMV
// designs serves a call, the workspace layout, the benchmark / experiment switches.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace teal {

static int g_forced_config = 0;
static long long* g_dbg_time = nullptr;

static bool g_pdl = false;
void set_debug_timestamps(long long* buf) { g_dbg_time = buf; }
long long* debug_timestamps() { return g_dbg_time; }
// Per-launch timeline slots for the TIMED row-chunk kernels: launch i of a timed region writes slot i.
static long long* g_timeline = nullptr;
static int g_timeline_slots = 0, g_timeline_next = 0;
void set_debug_timeline(long long* buf, int slots) { g_timeline = buf; g_timeline_slots = buf ? slots : 0; g_timeline_next = 0; }
long long* next_debug_timeline_slot() {
  if (!g_timeline || g_timeline_next >= g_timeline_slots) return nullptr;
  return g_timeline + (size_t)(g_timeline_next++) * kTimelineSlotInts;
}
int forced_config() { return g_forced_config; }
void set_pdl(bool on) { g_pdl = on; }
bool pdl_enabled() { return g_pdl; }
static unsigned long long* g_probe = nullptr;
void set_bytes_probe(unsigned long long* q) { g_probe = q; }
unsigned long long* bytes_probe() { return g_probe; }
static bool g_wave_balance = true;
bool wave_balance_enabled() { return g_wave_balance; }
static int g_sub_cols = 0;  // config ids k*1000 + id additionally cap the per-warp sub-tile of gather_gemv at k*256 columns
int sub_cols_override() { return g_sub_cols; }
static int g_tuning[TEAL_TUNE_COUNT] = {1, 0, 1, 1, 0, 0, 0, 0};  // generation 1, no speculation, per-line prefetch, FHFMA, own K chunks, no opts
int tuning(int knob) { return (knob >= 0 && knob < TEAL_TUNE_COUNT) ? g_tuning[knob] : 0; }
int set_tuning(int knob, int value) {
  if (knob < 0 || knob >= TEAL_TUNE_COUNT) return set_error(TEAL_ERR_BAD_ARG, "tuning knob out of range");
  if (knob == TEAL_TUNE_GENERATION && (value < 1 || value > 5)) return set_error(TEAL_ERR_BAD_ARG, "generation must be 1 .. 5");
  if (knob == TEAL_TUNE_SPEC_ROWS && (value < 0 || value > 32)) return set_error(TEAL_ERR_BAD_ARG, "spec_rows must be 0..32");
  if (knob == TEAL_TUNE_SPEC_MODE && value != 1 && value != 2) return set_error(TEAL_ERR_BAD_ARG, "spec_mode must be 1 or 2");
  if (knob == TEAL_TUNE_FHFMA && value != 0 && value != 1) return set_error(TEAL_ERR_BAD_ARG, "fhfma must be 0 or 1");
  if ((knob == TEAL_TUNE_ROWS_WIDE || knob == TEAL_TUNE_ROWS_NARROW) && (value < 0 || value > 4096 || (value & 31)))
    return set_error(TEAL_ERR_BAD_ARG, "rows per CTA must be a multiple of 32 in [0, 4096]");
  if (knob == TEAL_TUNE_WIDE_KELEMS && value < 0) return set_error(TEAL_ERR_BAD_ARG, "wide_kelems must be >= 0");
  if (knob == TEAL_TUNE_V1_OPTS && (value < 0 || value > 11)) return set_error(TEAL_ERR_BAD_ARG, "v1_opts is 0..11");
  g_tuning[knob] = value;
  return TEAL_OK;
}

static long long wide_min_elems() { return g_tuning[TEAL_TUNE_WIDE_KELEMS] > 0 ? 1000ll * g_tuning[TEAL_TUNE_WIDE_KELEMS] : 40ll * 1000 * 1000; }

int set_config(int id) {
  g_wave_balance = true;
  if (id >= 100 && id < 200) { g_wave_balance = false; id -= 70; }  // 100 + v: rowchunk variant v without wave balancing
  g_sub_cols = 0;
  if (id >= 1000) { g_sub_cols = (id / 1000) * 256; id %= 1000; }  // e.g. 2014: sub-tile 512, 4 tiles
  if (id < 0 || (id > kNumTmaConfigs && (id < 11 || id > 26) && (id < 30 || id > 79)) || g_sub_cols > 1024)
    return set_error(TEAL_ERR_BAD_ARG, "config id out of range");
  g_forced_config = id;
  return TEAL_OK;
}

// Config ids: 0 = default = rowchunk_gemv (many small CTAs), 30 + v = rowchunk variant v
// (bit 0: 512-column tiles, bits 1-2: rows per CTA = 128 << n);
// 1..6 = the TMA-bulk-ring kernel (sparse_gemv.cu); 10 + T = gather_gemv (cp.async rings) with T tiles per matrix.
// The two alternative kernels are kept as measured design points; they do not
// implement the sum-of-squares hand-off, so the fused-RMSNorm entry points always use rowchunk_gemv.
int dispatch_gemv(GemvParams& p, void* workspace, size_t ws_bytes, int out_cols_total, cudaStream_t stream) {
  const int id = g_forced_config;
  const bool needs_rc = p.in_mode != kInPlain || p.ssq_out || p.keep_out || p.mode == kModeSurvivors || p.tp || p.chain;
  if (id >= 1 && id <= kNumTmaConfigs && !needs_rc && p.out_mode == kOutPlain && !p.x_out)
    return launch_gemv(p, workspace, ws_bytes, out_cols_total, stream);
  if (id >= 11 && id <= 26 && !needs_rc)
    return launch_gather_gemv(p, workspace, ws_bytes, out_cols_total, stream, id - 10);
  int variant;
  if (id >= 30 && id <= 79 && ((!p.tp && !p.chain) || id == 31 || id == 32)) {  // (the fused TP exchange exists for variants 1 and 2)
    variant = id - 30;
  } else {  // measured on B200 (profiles/r01_bench_gemv_history.jsonl): wide tiles pay off once a launch moves >= ~40 MB
    long long elems = 0;
    for (int s = 0; s < p.nseg; ++s) elems += (long long)p.Z * p.seg[s].ncols;
    variant = (elems >= wide_min_elems()) ? 1 : 2;
    if (p.tp || p.chain) {
      // the fused TP exchange and the flag-chained pair live in the default generation
    } else if (g_tuning[TEAL_TUNE_GENERATION] == 2)  // opt-in, see rowchunk_gemv2_kernel
      variant = g_tuning[TEAL_TUNE_FHFMA] ? ((variant == 1) ? 13 : 15) : ((variant == 1) ? 12 : 14);
    if (!p.tp && !p.chain && g_tuning[TEAL_TUNE_GENERATION] == 3) variant = (variant == 1) ? 20 : 21;  // ... with warp-wise completion
    if (!p.tp && !p.chain && g_tuning[TEAL_TUNE_GENERATION] == 4) variant = (variant == 1) ? 24 : 25;  // cp.async rings
    if (!p.tp && !p.chain && g_tuning[TEAL_TUNE_GENERATION] == 5)  // lean list-driven loop, batch structure
      variant = g_tuning[TEAL_TUNE_FHFMA] ? ((variant == 1) ? 30 : 31) : ((variant == 1) ? 32 : 33);
  }
  return launch_rowchunk(p, workspace, ws_bytes, out_cols_total, stream, variant);
}

// Tiling of a flag-chained MLP pair, derived once for both launches with the default dispatch's size rule: the producer
// (gate/up, Z = dim, 2 x inter columns) must know the consumer's (down, Z = inter, dim columns) K chunk and tile count.
void chain_describe(int dim, int inter, int* rows_c, int* ctiles_c, int* cols_p) {
  const bool wide_p = (long long)dim * 2 * inter >= wide_min_elems(), wide_c = (long long)inter * dim >= wide_min_elems();
  *cols_p = wide_p ? 512 : 256;
  *rows_c = wide_c ? 128 : 64;
  *ctiles_c = (dim + (wide_c ? 512 : 256) - 1) / (wide_c ? 512 : 256);
}

// Geometry of a flag-chained attention block, derived once for its three launches (same size rule as the dispatch).
int chain_attn_describe(ChainAttn* ca, void* workspace, size_t ws_bytes, int dim, int n_head, int n_kv, int head_dim, bool uniform) {
  if (!workspace || ws_bytes < workspace_bytes(8)) return set_error(TEAL_ERR_WORKSPACE, "workspace too small");
  if (n_head <= 0 || n_kv <= 0 || n_head % n_kv || n_head > 128 || (head_dim != 64 && head_dim != 128))
    return set_error(TEAL_ERR_UNSUPPORTED, "chained attention block: heads / head_dim");
  const int N = (n_head + 2 * n_kv) * head_dim;
  ca->n_head = n_head; ca->n_kv = n_kv; ca->hd = head_dim;
  ca->cols = ((long long)dim * N >= wide_min_elems()) ? 512 : 256;
  const int widths[3] = {n_head * head_dim, n_kv * head_dim, n_kv * head_dim};
  ca->nseg = uniform ? 1 : 3;
  int col = 0, tile = 0;
  for (int s = 0; s < 3; ++s) ca->seg_col0[s] = ca->seg_ncols[s] = ca->seg_tile0[s] = 0;
  for (int s = 0; s < ca->nseg; ++s) {
    const int w = uniform ? N : widths[s];
    ca->seg_col0[s] = col; ca->seg_ncols[s] = w; ca->seg_tile0[s] = tile;
    col += w;
    tile += (w + ca->cols - 1) / ca->cols;
  }
  if (tile > 64) return set_error(TEAL_ERR_UNSUPPORTED, "chained attention block: more than 64 QKV tiles");
  for (int t = 0; t < 64; ++t) {
    const unsigned r = t < tile ? chain_qkv_readers(*ca, t) : 0u;
    if (r > 255u) return set_error(TEAL_ERR_UNSUPPORTED, "chained attention block: more than 255 readers of one QKV tile");
    ca->readers[t] = (unsigned char)r;
  }
  const bool wide_o = (long long)n_head * head_dim * dim >= wide_min_elems();
  ca->rows_o = wide_o ? 128 : 64;
  ca->ctiles_o = (dim + (wide_o ? 512 : 256) - 1) / (wide_o ? 512 : 256);
  unsigned* cells = reinterpret_cast<unsigned*>(workspace) + 3072;  // the ready cells live in the unused tail of the counter region
  ca->ready_qkv = cells + 128;
  ca->ready_att = cells + 256;
  return TEAL_OK;
}

// Grid size of an RMSNorm consumer launched with chain_wait_all (nominal K chunk, no wave balancing): what the previous
// block's last finaliser writes into the hand-off cell.  kind 0 = QKV (a = n_head, b = n_kv, c = head_dim, uniform),
// kind 1 = gate/up (a = inter).
int chain_consumer_ctas(int kind, int dim, int a, int b, int c, bool uniform) {
  if (kind == 1) {
    const bool wide = (long long)dim * 2 * a >= wide_min_elems();
    const int cols = wide ? 512 : 256, rows = wide ? 128 : 64;
    return 2 * ((a + cols - 1) / cols) * ((dim + rows - 1) / rows);
  }
  const int N = (a + 2 * b) * c;
  const bool wide = (long long)dim * N >= wide_min_elems();
  const int cols = wide ? 512 : 256, rows = wide ? 128 : 64;
  const int widths[3] = {a * c, b * c, b * c};
  int tiles = 0;
  if (uniform) tiles = (N + cols - 1) / cols;
  else for (int s = 0; s < 3; ++s) tiles += (widths[s] + cols - 1) / cols;
  return tiles * ((dim + rows - 1) / rows);
}

size_t workspace_bytes(int max_out_cols) {
  size_t acc = ((size_t)max_out_cols * 4 + 255) & ~(size_t)255;
  return (size_t)kMaxCounters * 4 + acc + (size_t)kSsqParts * 4;  // [counters][acc][ssq scratch]
}
float* workspace_ssq_scratch(void* workspace, size_t ws_bytes) {
  return reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(workspace) + ws_bytes - (size_t)kSsqParts * 4);
}

}  // namespace teal
