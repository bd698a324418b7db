// Step program: the whole batch-1 decode step as ONE persistent kernel that walks a table of phases
// (embedding, RMSNorm+QKV GEMV, RoPE+KV+attention, o GEMV+residual, RMSNorm+gate/up GEMV+SiLU*mul,
// down GEMV+residual, lm_head+arg-max, bookkeeping) item by item.  Items are handed out by one global
// ticket counter in dependency order; an item waits on the completion flags of exactly the producer
// tiles it reads (data-flow), not on a kernel boundary.
//
// This header holds the phase table layout and the item -> (tile, chunk, dependency) arithmetic.  The
// arithmetic is __host__ __device__: the planner, the CPU-side verifier (teal_step_program_item, used by
// tests/test_step_program.py) and the kernel all run the same code.
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

#include "common.cuh"

namespace teal {

constexpr int kSpWarpsDefault = 4;       // warps per CTA (teal_step_desc::warps: 4 or 2)
constexpr int kSpMaxDeps = 32;          // completion flags one item may wait on (one polling lane each)
constexpr unsigned kSpMagic = 0x54454132u;  // "TEA2"
constexpr int kSpMaxTpi = 4;            // column tiles one item may sweep (same K chunk: x and survivor masks are reused)
constexpr int kSpBcast = 32;            // replicas of a phase-complete broadcast word
constexpr int kSpBcastStride = 16;      // in 8-byte words: one 128-byte line per replica

enum { kPhEmbed = 0, kPhGemv = 1, kPhAttn = 2, kPhHead = 3, kPhAdvance = 4 };
enum { kDepNone = 0, kDepAll = 1, kDepLanes = 2, kDepQkvHead = 3 };

// What an item waits for before it reads its input.
struct SpDep {
  int kind;
  int cell;        // kDepAll: done[cell] must reach epoch * per_epoch (every publication of the producer phase)
  int per_epoch;
  int flag0;       // kDepLanes / kDepQkvHead: first completion flag of the producer phase
  int div;         // kDepLanes: input lanes per producer flag (producer column-tile width, or head_dim)
  int bcast0;      // kDepAll, experimental: replica block of the producer's "phase complete" broadcast word (-1: none)
  int pad[2];
};

// What a finished tile / head / item publishes.
struct SpPub {
  int flag0;       // flags[flag0 + i] = epoch when publication i of this phase is complete (-1: no flags)
  int cell;        // done[cell] += 1 per publication
  int count;       // publications per epoch
  int bcast0;      // experimental: the publication that completes the phase writes {epoch, RMSNorm scale} to
                   // kSpBcast replicas (one L2 sector each) instead of every consumer polling one cell (-1: off)
  const float* ssq;  // ... computing the scale of the NEXT RMSNorm from this phase's sum-of-squares partials
  int n_ssq, rs_Z;
  float rs_eps;
  int pad[3];
};

struct SpSeg {
  const __half* w;  // fp16 [Z][ldw], first column of the region
  long long ldw;
  int ncols, out_col0;
  float thr;        // fp16-representable
  int ctile0, nctiles;
  int pad;
};

struct SpGemv {
  const __half* x;
  const __half* norm_w;
  const float* ssq_in;
  __half* y;
  const __half* resid;
  float* ssq_out;
  float* acc;          // fp32 accumulator region of this projection kind (zero on entry, re-zeroed by the finaliser)
  unsigned* counters;  // per-tile arrival counters of this projection kind
  int Z, mode, in_mode, out_mode, nseg, n_ssq_in;
  int variant;         // 0: 256-column tiles, 8 rows in flight; 1: 512-column tiles, 4 rows in flight
  int rows_per_item;   // K chunk
  int lpw;             // input lanes per warp per pass (<= 32): rows_per_item = warps * lpw * passes
  int nkchunks, total_ctiles, target;
  int group;           // order positions per group: items run K-chunk-major inside a group, groups in sequence
  int rot;             // plain mode: order position p -> column tile (p + rot) % total_ctiles (k, v tiles before q)
  float eps;
  int tpi;             // column tiles per item (experimental, <= kSpMaxTpi; 0 / 1 = one tile per item)
  SpSeg seg[3];
};

struct SpAttn {
  const __half* qkv;
  const __half* freqs;
  __half* k_cache;
  __half* v_cache;
  const int* pos;
  __half* out;
  int n_head, n_kv, hd, max_seq;
  float scale;
  int prod_cols;       // column-tile width of the producing QKV phase
  int prod_nseg;
  int prod_col0[3], prod_ncols[3], prod_ctile0[3];
};

struct SpHead {
  const __half* x;
  const __half* norm_w;
  const __half* W;     // [V][ldw] row-major (never sparsified, gpt-fast/model.py:141-142)
  __half* y;
  unsigned long long* argmax;
  long long ldw;
  int Z, V, rows_per_item;
  float eps;
};

struct SpEmbed {
  const __half* emb;
  const int* token;
  __half* x;
  float* ssq_out;
  int dim, pad;
};

struct SpAdvance {
  unsigned long long* argmax;
  int* token;
  int* pos;
  int* out_tokens;
  int out_cap, pad;
};

struct alignas(16) SpPhase {
  int kind, item0, nitems, layer;
  SpDep dep;
  SpPub pub;
  union {
    SpGemv g;
    SpAttn a;
    SpHead h;
    SpEmbed e;
    SpAdvance v;
  } u;
};
static_assert(sizeof(SpPhase) % 16 == 0, "SpPhase is copied to shared memory in 16-byte words");

struct alignas(16) SpHeader {
  unsigned magic;
  int nphases, total_items, smem_bytes;
  unsigned* ticket;   // control block in the scratch buffer: [ticket][exited][epoch]
  unsigned* exited;
  unsigned* epoch;
  unsigned* flags;
  unsigned* done;
  int nflags, ndone;
  unsigned long long scratch_bytes;
  int ctas_per_sm;    // cap on resident CTAs per SM (0 = occupancy)
  int poll_ns;        // back-off between two polls of a flag
  int opts;           // bit 0: cache satisfied whole-phase dependencies per CTA; bit 1: RMSNorm scale once per (CTA, phase)
  int warps;          // warps per CTA the plan was made for (4 or 2)
  unsigned long long* dbg;  // optional per-phase timestamps [nphases][4] (teal_step_desc::debug_times)
  int stagger_ns;     // de-synchronise the CTAs released by a whole-phase wait
  int ext;            // 1: needs the kernel instantiation with the experimental paths
  unsigned long long* bcast;  // [nphases][kSpBcast] broadcast words, kSpBcastStride apart (opts bit 2), or null
};
static_assert(sizeof(SpHeader) % 16 == 0, "phase table must stay 16-byte aligned");

// ------------------------------------------------------------------------------------------------
// item -> (column tile, K chunk).  Order position p runs over the column tiles in "publication order":
//   pair mode : gate tile j and up tile j are neighbours (p = 2j, 2j+1), so tile pair j completes early;
//   plain mode: rotated by `rot` (QKV: the k and v tiles first, every attention head needs them).
// Positions are cut into groups of `group`; inside a group the K chunk is the slow index (neighbouring
// CTAs read neighbouring pieces of the same weight rows), groups follow each other, so the tiles of group
// 0 are complete -- and their consumers may start -- while later groups still stream.
// ------------------------------------------------------------------------------------------------
__host__ __device__ inline void sp_gemv_decode(const SpGemv& g, int item, int* ct, int* kc) {
  const int T = g.total_ctiles, nk = g.nkchunks, G = g.group;
  const int per_group = G * nk;
  const int grp = item / per_group;
  const int r = item - grp * per_group;
  const int left = T - grp * G;
  const int gsz = left < G ? left : G;
  const int k = r / gsz;
  const int p = grp * G + (r - k * gsz);
  int c;
  if (g.mode == kModePairSilu) {
    c = (p & 1) ? (T >> 1) + (p >> 1) : (p >> 1);
  } else {
    c = p + g.rot;
    if (c >= T) c -= T;
  }
  *ct = c;
  *kc = k;
}

// order position -> column tile
__host__ __device__ inline int sp_gemv_pos_tile(const SpGemv& g, int p) {
  const int T = g.total_ctiles;
  if (g.mode == kModePairSilu) return (p & 1) ? (T >> 1) + (p >> 1) : (p >> 1);
  int c = p + g.rot;
  if (c >= T) c -= T;
  return c;
}

// Multi-tile items (experimental): an item sweeps up to `tpi` consecutive order positions of one group for ONE K chunk
// (the x lanes are loaded once; the first rows of tile t+1 are on their way from DRAM while tile t is reduced).
// group % tpi == 0 unless the phase is a single group; the last group may end with a shorter item.
__host__ __device__ inline void sp_gemv_decode_multi(const SpGemv& g, int item, int* p0, int* ntiles, int* kc) {
  const int T = g.total_ctiles, nk = g.nkchunks, G = g.group, tpi = g.tpi;
  const int per_full = (G / tpi) * nk;            // items of a full group
  const int full_groups = (G % tpi == 0) ? T / G : 0;  // a group that is not whole items is the only group (G == T)
  int grp = full_groups ? item / per_full : 0;
  if (grp > full_groups) grp = full_groups;
  const int r = item - grp * per_full;
  const int left = T - grp * G;
  const int gsz = left < G ? left : G;
  const int ipk = (gsz + tpi - 1) / tpi;          // items per K chunk in this group
  const int k = r / ipk;
  const int j = r - k * ipk;
  *p0 = grp * G + j * tpi;
  const int rem = gsz - j * tpi;
  *ntiles = rem < tpi ? rem : tpi;
  *kc = k;
}
__host__ __device__ inline int sp_gemv_items(const SpGemv& g) {
  if (g.tpi <= 1) return g.total_ctiles * g.nkchunks;
  const int T = g.total_ctiles, G = g.group, tpi = g.tpi;
  const int full_groups = (G % tpi == 0) ? T / G : 0;
  const int tail = T - full_groups * G;
  return (full_groups * (G / tpi) + (tail + tpi - 1) / tpi) * g.nkchunks;
}

// publication index of column tile ct (pair mode: gate tile j and up tile j share completion counter j)
__host__ __device__ inline int sp_gemv_cidx(const SpGemv& g, int ct) {
  return (g.mode == kModePairSilu && ct >= g.seg[1].ctile0) ? ct - g.seg[1].ctile0 : ct;
}

__host__ __device__ inline int sp_attn_col_tile(const SpAttn& a, int col) {
  for (int s = 0; s < a.prod_nseg; ++s)
    if (col >= a.prod_col0[s] && col < a.prod_col0[s] + a.prod_ncols[s])
      return a.prod_ctile0[s] + (col - a.prod_col0[s]) / a.prod_cols;
  return 0;
}

// Absolute indices of the completion flags item `item` of phase `ph` must see before it may read its input
// (kDepLanes / kDepQkvHead).  Returns the count, at most kSpMaxDeps (the planner guarantees the bound).
template <bool MULTI = false>
__host__ __device__ inline int sp_item_deps(const SpPhase& ph, int item, int* out) {
  int n = 0;
  if (ph.dep.kind == kDepLanes) {
    const SpGemv& g = ph.u.g;
    int ct, kc;
    if (MULTI && g.tpi > 1) {
      int nt;
      sp_gemv_decode_multi(g, item, &ct, &nt, &kc);
    } else {
      sp_gemv_decode(g, item, &ct, &kc);
    }
    const int kbeg = kc * g.rows_per_item;
    int kend = kbeg + g.rows_per_item;
    if (kend > g.Z) kend = g.Z;
    for (int f = kbeg / ph.dep.div; f <= (kend - 1) / ph.dep.div && n < kSpMaxDeps; ++f) out[n++] = ph.dep.flag0 + f;
  } else if (ph.dep.kind == kDepQkvHead) {
    const SpAttn& a = ph.u.a;
    const int h = item, kvh = h / (a.n_head / a.n_kv);
    const int c0[3] = {h * a.hd, a.n_head * a.hd + kvh * a.hd, (a.n_head + a.n_kv) * a.hd + kvh * a.hd};
    for (int r = 0; r < 3; ++r) {
      const int t0 = sp_attn_col_tile(a, c0[r]), t1 = sp_attn_col_tile(a, c0[r] + a.hd - 1);
      for (int t = t0; t <= t1 && n < kSpMaxDeps; ++t) {
        bool dup = false;
        for (int j = 0; j < n; ++j) dup = dup || (out[j] == ph.dep.flag0 + t);
        if (!dup) out[n++] = ph.dep.flag0 + t;
      }
    }
  }
  return n;
}

// host side (step_program.cu)
struct StepScratchLayout {
  size_t control, done, flags, counters[4], acc[4], ssq[2], bcast, total;
  int nflags, ndone, ncounters[4];
};

}  // namespace teal
