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
 *   hb_rng_normal(a,b)  approximately N(0, 1), integer-only, from two 32-bit words  (zeta, R/internals.R:194,158)
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

#endif /* HYDROBAYES_B200_RNG_H */
