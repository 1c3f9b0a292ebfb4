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
 *   hb_rng_u53(x)        64 random bits  -> double in the open interval (0, 1), 53 bits
 *   hb_rng_u32(w)        32 random bits  -> double in (0, 1), exact
 *   hb_rng_u16(h)        16 random bits  -> double in (0, 1), exact: (h + 0.5) 2^-16
 *   hb_rng_lambda16(h,c) U(-c, c) on 65536 symmetric levels               (R/internals.R:188)
 *   hb_rng_normal32(w)   approximately N(0, 1), integer-only, from one 32-bit word  (zeta, R/internals.R:194,158)
 *
 * Per-dimension draws (slot map v2): one Philox4x32 call serves TWO dimensions.  For dimension k the call is
 * slot HB_RNG_SLOT_DIM0 + (k >> 1); dimension k uses words (0,1) when k is even and words (2,3) when k is odd:
 *   lo word: bits 0-15  -> zt_k      = hb_rng_u16(lo & 0xffff)            (crossover test, R/internals.R:175-177)
 *            bits 16-31 -> lambda_k  = hb_rng_lambda16(lo >> 16, c_val)   (:188)
 *   hi word:            -> zeta_k    = c_star * hb_rng_normal32(hi)       (:194; snooker :158)
 * 64 random bits per dimension: the proposal kernel is instruction-bound on Philox, so the draws are sized to what
 * the sampler can observe (a crossover probability exact to 2^-16, a jitter on 65536 levels).
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

HB_RNG_FN double hb_rng_u16(uint32_t h) { return HB_DMUL(HB_DADD((double)(h & 0xffffu), 0.5), 0x1.0p-16); }
HB_RNG_FN double hb_rng_lambda16(uint32_t h, double c_val) {
  return HB_DADD(-c_val, HB_DMUL(HB_DMUL(2.0, c_val), hb_rng_u16(h)));
}
/* zt_k < CR  <=>  (h + 0.5) 2^-16 < CR  <=>  h < ceil(CR 2^16 - 0.5): the exact integer form of the crossover test */
HB_RNG_FN uint32_t hb_rng_zt_threshold16(double cr) {
  const double t = HB_DADD(HB_DMUL(cr, 65536.0), -0.5);
  uint32_t n = (uint32_t)t;                 /* truncation; t >= 0 for cr >= 2^-17 */
  if ((double)n < t) n += 1u;               /* ceil */
  return n;
}

/*
 * zeta's standard deviate: a bit-exact, integer-only approximation of N(0, 1) from 64 random bits.
 *   B = popcount(48 bits) ~ Binomial(48, 1/2)  (mean 24, variance 12: a lattice normal by the CLT)
 *   T = u8 + u8' - 255                          (triangular kernel of width 2 lattice steps, in units of 1/256)
 *   z = ((B - 24) * 256 + T) / (256 * sqrt(12 + (256^2 - 1) / (6 * 256^2)))
 * The result has mean 0, variance 1, a continuous piecewise-linear-like density, |z| <= 7.17, excess kurtosis -1/24
 * (Kolmogorov distance to N(0,1) about 3e-3).  zeta = c_star * z is the 1e-12-scale jitter of R/internals.R:194
 * whose only role is to keep the chain ergodic (Vrugt 2016, Sect. 3.3); its law is never observable in the output,
 * so ~10 integer instructions per dimension replace a ~100-instruction Box-Muller in the HBM-bound proposal kernel.
 * Replay (tape) mode does not use this function: there zeta comes from the tape as R's rnorm returned it.
 */
HB_RNG_FN int hb_rng_popc32(uint32_t x) {
#if defined(__CUDA_ARCH__)
  return __popc(x);
#else
  return __builtin_popcount(x);
#endif
}
HB_RNG_FN double hb_rng_normal(uint32_t wa, uint32_t wb) {
  const int b = hb_rng_popc32(wa) + hb_rng_popc32(wb >> 16);
  const int t = (int)(wb & 0xffu) + (int)((wb >> 8) & 0xffu) - 255;
  const int v = (b - 24) * 256 + t;
  return HB_DMUL((double)v, 0.00111988718555951);   /* 1 / (256 * sqrt(12 + 65535/393216)) */
}

/*
 * 32-bit variant used by the per-dimension draws (slot map v2):
 *   B = popcount(24 bits) ~ Binomial(24, 1/2)   (mean 12, variance 6)
 *   T = n4 + n4' - 15                            (triangular kernel, in units of 1/16 lattice step)
 *   z = ((B - 12) * 16 + T) / sqrt(6 * 256 + 2 * (16^2 - 1) / 12) = v / sqrt(1578.5)
 * mean 0, variance 1, |z| <= 5.21, excess kurtosis -0.079, Kolmogorov distance to N(0,1) 1.9e-3.
 */
HB_RNG_FN double hb_rng_normal32(uint32_t w) {
  const int b = hb_rng_popc32(w & 0xffffffu);
  const int t = (int)((w >> 24) & 0xfu) + (int)(w >> 28) - 15;
  const int v = (b - 12) * 16 + t;
  return HB_DMUL((double)v, 0x1.9c614ae8b41e8p-6);   /* 1 / sqrt(1578.5) */
}
#define HB_RNG_SLOT_DIM0 8u

#endif /* HYDROBAYES_B200_RNG_H */
