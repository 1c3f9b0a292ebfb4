/*
 * hydrobayes_b200_rng.h -- specification of the native-RNG draw transforms (host + device, plain C).
 *
 * The reference draws from R's Mersenne-Twister (R/internals.R:119-194, 241); the B200 sampler draws from
 * Philox4x32-10 keyed by (seed; chain, generation, slot).  Everything that turns Philox words into the values the
 * sampler uses is defined here with IEEE-754 correctly rounded operations ONLY (+, -, *, /, sqrt, conversions), in a
 * fixed order, so that any conforming implementation -- the CUDA kernels and the CPU oracle -- produces bit-identical
 * draws.  That matters: for d = 1 the DREAM population map has a large Lyapunov exponent (gamma = 1.68), a 1e-18
 * difference in one jitter value grows to 1e-9 within ~40 generations.
 *
 *   hb_rng_u53(x)       64 random bits  -> double in the open interval (0, 1), 53 bits
 *   hb_rng_u32(w)       32 random bits  -> double in (0, 1), exact
 *   hb_rng_lambda(w,c)  U(-c, c)                                        (R/internals.R:188)
 *   hb_rng_normal(a,b)  N(0, 1) by Box-Muller in binary32, from two 32-bit words    (zeta, R/internals.R:194,158)
 *
 * hb_rng_normal has float accuracy (about 1e-7 relative, |z| <= 6.66): zeta is a jitter of scale c_star (1e-12 by
 * default) whose only role is ergodicity (Vrugt 2016, Sect. 3.3), not a quantity that is ever reported.
 */
#ifndef HYDROBAYES_B200_RNG_H
#define HYDROBAYES_B200_RNG_H

#include <stdint.h>

#if defined(__CUDA_ARCH__)
#define HB_RNG_FN __device__ __forceinline__
#define HB_FADD(a, b) __fadd_rn((a), (b))
#define HB_FMUL(a, b) __fmul_rn((a), (b))
#define HB_FDIV(a, b) __fdiv_rn((a), (b))
#define HB_FSQRT(a) __fsqrt_rn((a))
#define HB_DADD(a, b) __dadd_rn((a), (b))
#define HB_DMUL(a, b) __dmul_rn((a), (b))
#else
/* host: compile with -ffp-contract=off (no fused multiply-add); x86-64 SSE arithmetic is IEEE binary32/binary64 */
#if defined(__CUDACC__)
#define HB_RNG_FN __host__ __device__ __forceinline__
#else
#define HB_RNG_FN static inline
#endif
#define HB_FADD(a, b) ((float)((float)(a) + (float)(b)))
#define HB_FMUL(a, b) ((float)((float)(a) * (float)(b)))
#define HB_FDIV(a, b) ((float)((float)(a) / (float)(b)))
#define HB_FSQRT(a) hb_rng_sqrtf_host((a))
#define HB_DADD(a, b) ((double)((double)(a) + (double)(b)))
#define HB_DMUL(a, b) ((double)((double)(a) * (double)(b)))
#include <math.h>
static inline float hb_rng_sqrtf_host(float a) { return sqrtf(a); } /* sqrtf is correctly rounded (IEEE 754) */
#endif

HB_RNG_FN double hb_rng_u53(uint64_t x) { return HB_DADD(HB_DMUL((double)(x >> 11), 0x1.0p-53), 0x1.0p-54); }
HB_RNG_FN double hb_rng_u32(uint32_t w) { return HB_DMUL(HB_DADD((double)w, 0.5), 0x1.0p-32); }
HB_RNG_FN double hb_rng_lambda(uint32_t w, double c_val) {
  return HB_DADD(-c_val, HB_DMUL(HB_DMUL(2.0, c_val), hb_rng_u32(w)));
}

/* ln(u) for a binary32 u in (0, 1]:  u = m * 2^e, m in [sqrt(1/2), sqrt(2));  ln m = 2 atanh(s), s = (m-1)/(m+1) */
HB_RNG_FN float hb_rng_logf(float u) {
  union { float f; uint32_t i; } v;
  v.f = u;
  int e = (int)(v.i >> 23) - 127;
  v.i = (v.i & 0x007fffffu) | 0x3f800000u; /* m in [1, 2) */
  float m = v.f;
  if (m > 1.41421356f) { m = HB_FMUL(m, 0.5f); e += 1; }
  const float s = HB_FDIV(HB_FADD(m, -1.0f), HB_FADD(m, 1.0f));
  const float s2 = HB_FMUL(s, s);
  /* 2 (s + s^3/3 + s^5/5 + s^7/7 + s^9/9), Horner in s2 */
  float p = 0.111111111f;
  p = HB_FADD(HB_FMUL(p, s2), 0.142857143f);
  p = HB_FADD(HB_FMUL(p, s2), 0.2f);
  p = HB_FADD(HB_FMUL(p, s2), 0.333333333f);
  p = HB_FADD(HB_FMUL(p, s2), 1.0f);
  const float lnm = HB_FMUL(HB_FMUL(2.0f, s), p);
  return HB_FADD(HB_FMUL((float)e, 0.693147181f), lnm);
}

/* cos(2 pi t) for a binary32 t in (0, 1]: quadrant reduction (exact) + Taylor polynomials on [-pi/4, pi/4] */
HB_RNG_FN float hb_rng_cos2pif(float t) {
  const float q4 = HB_FMUL(t, 4.0f);
  const int q = (int)HB_FADD(q4, 0.5f);                       /* round to nearest quadrant 0..4 */
  const float r = HB_FADD(t, -HB_FMUL((float)q, 0.25f));      /* exact: |r| <= 1/8 */
  const float y = HB_FMUL(r, 6.28318531f);
  const float y2 = HB_FMUL(y, y);
  float c = 2.48015873e-5f;                                    /* 1/8! */
  c = HB_FADD(HB_FMUL(c, y2), -1.38888889e-3f);
  c = HB_FADD(HB_FMUL(c, y2), 4.16666667e-2f);
  c = HB_FADD(HB_FMUL(c, y2), -0.5f);
  c = HB_FADD(HB_FMUL(c, y2), 1.0f);
  float sn = 2.75573192e-6f;                                   /* 1/9! */
  sn = HB_FADD(HB_FMUL(sn, y2), -1.98412698e-4f);
  sn = HB_FADD(HB_FMUL(sn, y2), 8.33333333e-3f);
  sn = HB_FADD(HB_FMUL(sn, y2), -0.166666667f);
  sn = HB_FADD(HB_FMUL(sn, y2), 1.0f);
  sn = HB_FMUL(sn, y);
  switch (q & 3) {
    case 0: return c;
    case 1: return -sn;
    case 2: return -c;
    default: return sn;
  }
}

HB_RNG_FN double hb_rng_normal(uint32_t wa, uint32_t wb) {
  const float u1 = (float)hb_rng_u32(wa);                     /* (0, 1]: rounds to 1.0f for the largest words */
  const float u2 = (float)hb_rng_u32(wb);
  const float r = HB_FSQRT(HB_FMUL(-2.0f, hb_rng_logf(u1)));
  return (double)HB_FMUL(r, hb_rng_cos2pif(u2));
}

#endif /* HYDROBAYES_B200_RNG_H */
