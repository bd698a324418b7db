/* oracle/teal_oracle_c.c -- plain-C restatement of the hot path (TEST INFRASTRUCTURE ONLY, never linked into
 * teal_b200).  A second, torch-free statement of what one projection computes, used by tests to cross-check the
 * Python oracle's rounding claims:
 *
 *   survivors : m such that |x[m]| > fp16(t)            utils/utils.py:51-52 (`x.abs().gt(threshold)` on an fp16
 *                                                       tensor: the Python float is demoted to fp16, strict >)
 *   y[n]      : fp16( sum_{m in survivors, ascending} fp32(W[n][m]) * fp32(x[m]) )   kernels/sparse_gemv.py:60-83
 *               with the weight in the reference's "column major" storage wt[m*ldw + n] (generate.py:296-317)
 *
 * Half <-> float conversions are done bit-exactly here (round-to-nearest-even), no compiler fp16 support needed.
 * Build: gcc -O2 -shared -fPIC -o oracle/_build/libteal_oracle_c.so oracle/teal_oracle_c.c   (see oracle/Makefile)
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

static float h2f(uint16_t h) {
  uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
  uint32_t exp = (h >> 10) & 0x1fu;
  uint32_t man = h & 0x3ffu;
  uint32_t bits;
  if (exp == 0) {
    if (man == 0) {
      bits = sign;
    } else { /* subnormal: normalise */
      int e = -1;
      do { man <<= 1; ++e; } while (!(man & 0x400u));
      bits = sign | ((uint32_t)(127 - 15 - e) << 23) | ((man & 0x3ffu) << 13);
    }
  } else if (exp == 31) {
    bits = sign | 0x7f800000u | (man << 13);
  } else {
    bits = sign | ((exp + 127 - 15) << 23) | (man << 13);
  }
  float f;
  memcpy(&f, &bits, 4);
  return f;
}

/* float/double -> half, round to nearest even (what torch does when it demotes a scalar / rounds a result) */
static uint16_t d2h(double d) {
  if (d != d) return 0x7e00u;
  uint16_t sign = 0;
  if (d < 0 || (d == 0 && 1.0 / d < 0)) { sign = 0x8000u; d = -d; }
  if (d >= 65520.0) return sign | 0x7c00u;            /* rounds to inf */
  if (d < 5.9604644775390625e-08 / 2) return sign;     /* below half of the smallest subnormal */
  int e;
  double m = frexp(d, &e);                             /* d = m * 2^e, m in [0.5, 1) */
  int he = e - 1 + 15;                                 /* biased exponent for 1.xxx * 2^(e-1) */
  double scaled;
  if (he <= 0) {                                       /* subnormal: units of 2^-24 */
    scaled = ldexp(d, 24);
    double r = nearbyint(scaled);                      /* default rounding mode: to nearest even */
    return sign | (uint16_t)r;                         /* r may reach 1024 == smallest normal: still correct */
  }
  scaled = ldexp(m, 11);                               /* in [1024, 2048) */
  double r = nearbyint(scaled);
  if (r >= 2048.0) { r = 1024.0; ++he; }
  if (he >= 31) return sign | 0x7c00u;
  return sign | (uint16_t)(he << 10) | (uint16_t)((uint32_t)r - 1024u);
}

/* effective threshold of the reference compare */
double teal_ref_round_threshold(double t) { return (double)h2f(d2h(t)); }

/* idx_out[Z]; returns the number of survivors (NaN lanes do not survive the compare, as in torch) */
int teal_ref_survivors(const uint16_t* x, int Z, double threshold, int32_t* idx_out) {
  const float t16 = h2f(d2h(threshold));
  int n = 0;
  for (int m = 0; m < Z; ++m) {
    const float a = fabsf(h2f(x[m]));
    if (a > t16) idx_out[n++] = m;
  }
  return n;
}

/* y (fp16 bits) and y32 (optional fp32, unrounded) for wt stored [Z][ldw] */
void teal_ref_sparse_gemv(const uint16_t* x, const uint16_t* wt, int Z, int N, long long ldw, double threshold,
                          uint16_t* y, float* y32) {
  const float t16 = h2f(d2h(threshold));
  for (int n = 0; n < N; ++n) {
    float acc = 0.f;
    for (int m = 0; m < Z; ++m) {
      const float xv = h2f(x[m]);
      if (fabsf(xv) > t16) acc += h2f(wt[(size_t)m * (size_t)ldw + n]) * xv; /* exact product, fp32 sum */
    }
    if (y32) y32[n] = acc;
    y[n] = d2h((double)acc);
  }
}

/* half round trip helpers exposed for the tests */
float teal_ref_h2f(uint16_t h) { return h2f(h); }
uint16_t teal_ref_f2h(double d) { return d2h(d); }
