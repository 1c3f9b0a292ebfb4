// hb_common.cuh -- shared device helpers: Philox4x32-10, uniform conversions, group (sub-warp) primitives.
//
// Philox slot map (DESIGN.md "Philox slot map"): counter = (chain_lo, chain_hi, generation, slot), key = seed.
//   slot 0      words 0-1 -> u_snooker (53 bit)        words 2-3 -> D = 1 + mulhi64(r, delta)      R/internals.R:119,123
//   slot 1..4   two 64-bit draws each -> partial Fisher-Yates partner draws 0..7 (delta <= 4)     R/internals.R:134,142
//   slot 5      words 0-1 -> u_id (crossover index)     words 2-3 -> u_gamma                        R/internals.R:173,192
//   slot 6      words 0-1 -> u_accept                   words 2-3 -> g_snooker = 1.2 + u            R/internals.R:241,156
//   slot 8+(k>>1)  dimension k uses words (0,1) if k is even, (2,3) if odd (slot map v2, hydrobayes_b200_rng.h):
//               lo bits 0-15 -> zt_k, lo bits 16-31 -> lambda_k, hi -> zeta_k (hb_rng_normal32)   R/internals.R:175,188,194
//   outlier replacement draws: chain = (0xFFFFFFFF << 32) | run, slot = q >> 1, pair q & 1 (host side) R/internals.R:81
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/hydrobayes_b200_rng.h"

#define HB_FLOOR (-708.3964185322641)  // log(.Machine$double.xmin), R/internals.R:19,65

// error bits raised by kernels into hb_ctrl::err
enum : unsigned {
  HB_FLAG_PROP_NONFINITE = 1u,   // R/internals.R:209
  HB_FLAG_LP_NONFINITE = 2u,     // R/internals.R:67
  HB_FLAG_LL_NONFINITE = 4u,     // R/internals.R:21
  HB_FLAG_J_NAN = 8u,            // if(any(J > 0)) on NaN, R/dream_parallel.R:407
  HB_FLAG_BAD_TAPE = 16u,
  HB_FLAG_TARGET_DOMAIN = 32u,   // HydroModel: param < 0, R/test_HydroModel.R:18
  HB_FLAG_PEER_TIMEOUT = 64u     // HB_EXCHANGE_DELTA: a peer GPU did not reach the generation barrier
};

enum { HB_SLOT_HEAD = 0, HB_SLOT_PARTNER0 = 1, HB_SLOT_CHOICE = 5, HB_SLOT_ACCEPT = 6, HB_SLOT_DIM0 = 8 };

#define HB_MAX_NCR 16
#define HB_MAX_DELTA 4

// Device-resident control block: everything the reference keeps in R scalars/vectors between generations.
struct hb_ctrl {
  double pCR[HB_MAX_NCR];        // R/dream_parallel.R:192
  double J[HB_MAX_NCR];          // :183
  double n_id[HB_MAX_NCR];       // :183 (doubles holding integers, like R)
  unsigned long long n_id_gen[HB_MAX_NCR];  // per-generation integer counters (atomics), folded by the finalize kernel
  unsigned long long n_acc_gen;
  double n_acc;                  // :184, per AR interval
  double n_eval;                 // :185
  double cum_eval;               // out_CR[i_ar-1,1] + n_eval, :430
  long long i_ar;                // :250
  unsigned err;                  // HB_FLAG_*
  unsigned pad;
  // per-crossover acceptance counters of the sequential sweep (R/dream.R:143,241,266)
  unsigned long long n_acc_id_gen[HB_MAX_NCR];
  double n_acc_id[HB_MAX_NCR];
};

__host__ __device__ __forceinline__ uint4 hb_philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                          uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    const uint32_t n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    const uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}

// same function with the 10 round keys precomputed (they depend on the seed only): ks[2r], ks[2r+1]
__host__ __device__ __forceinline__ uint4 hb_philox4x32_10_ks(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                             const uint32_t* __restrict__ ks) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ ks[2 * r];
    const uint32_t n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ ks[2 * r + 1];
    const uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
  }
  return make_uint4(c0, c1, c2, c3);
}
inline void hb_philox_key_schedule(uint64_t seed, uint32_t ks[20]) {
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  for (int r = 0; r < 10; ++r) { ks[2 * r] = k0; ks[2 * r + 1] = k1; k0 += 0x9E3779B9u; k1 += 0xBB67AE85u; }
}

struct hb_phx {
  uint32_t k0, k1, c0, c1, gen;
  __host__ __device__ __forceinline__ hb_phx(uint64_t seed, uint64_t gchain, uint32_t generation)
      : k0((uint32_t)seed), k1((uint32_t)(seed >> 32)), c0((uint32_t)gchain), c1((uint32_t)(gchain >> 32)),
        gen(generation) {}
  __host__ __device__ __forceinline__ uint4 slot(uint32_t s) const { return hb_philox4x32_10(c0, c1, gen, s, k0, k1); }
};

__host__ __device__ __forceinline__ uint64_t hb_pair64(const uint4& w, int which) {
  return which ? (((uint64_t)w.w << 32) | w.z) : (((uint64_t)w.y << 32) | w.x);
}
// open-interval uniforms, identical arithmetic on host and device (exact in binary floating point)
__host__ __device__ __forceinline__ double hb_u53(uint64_t x) { return hb_rng_u53(x); }
__host__ __device__ __forceinline__ double hb_u32d(uint32_t x) { return hb_rng_u32(x); }

#ifdef __CUDACC__
__device__ __forceinline__ uint64_t hb_mulhi64(uint64_t a, uint64_t b) { return __umul64hi(a, b); }

// zeta ~ N(0, c_star^2): bit-reproducible integer-CLT deviate of include/hydrobayes_b200_rng.h
__device__ __forceinline__ double hb_zeta32(uint32_t hi, double c_star) {
  return __dmul_rn(c_star, hb_rng_normal32(hi));
}

// sparse partial Fisher-Yates over 0..n-1 (R's sample(n, k) without replacement: y = x[r]; x[r] = x[--n])
struct hb_fy {
  long long pos[2 * HB_MAX_DELTA];
  long long val[2 * HB_MAX_DELTA];
  int cnt;
  __device__ __forceinline__ hb_fy() : cnt(0) {}
  __device__ __forceinline__ long long get(long long p) const {
    long long v = p;
#pragma unroll
    for (int i = 0; i < 2 * HB_MAX_DELTA; ++i)
      if (i < cnt && pos[i] == p) v = val[i];  // later entries override earlier ones
    return v;
  }
  __device__ __forceinline__ long long draw(uint64_t rnd, long long& n_rem) {
    const long long r = (long long)hb_mulhi64(rnd, (uint64_t)n_rem);
    const long long y = get(r);
    const long long last = get(n_rem - 1);
#pragma unroll
    for (int i = 0; i < 2 * HB_MAX_DELTA; ++i)
      if (i == cnt) { pos[i] = r; val[i] = last; }
    cnt++;
    n_rem--;
    return y;
  }
};

// bound_par, R/internals.R:89-113.  reflect / fold are R's literal while-loops for up to HB_BOUND_MAX_ITER passes (bit-identical
// to the reference wherever the reference terminates in reasonable time: a proposal up to that many box widths outside);
// beyond that the remaining whole periods are removed in closed form (fmod on the 2w period of reflect, w of fold) and the
// loop finishes the job, so a gamma = 1 jump of an over-dispersed population cannot spin a kernel for 1e15 iterations.
// A non-finite value (already flagged as HB_FLAG_PROP_NONFINITE by the caller) is returned unchanged: R's loop would never
// end on it.  hb_set_bounds rejects min > max, and min == max for reflect / fold.
#define HB_BOUND_MAX_ITER 1024
__device__ __forceinline__ double hb_bound_par(double par, double mn, double mx, int handle) {
  double o = par;
  if (par < mn || par > mx) {
    if (handle == 1) {
      o = fmax(fmin(par, mx), mn);
    } else if (handle == 2 || handle == 3) {
      if (!(fabs(par) <= 1.7976931348623157e308)) return par;
      const double w = mx - mn;
      for (int pass = 0; pass < 2; ++pass) {
        int it = 0;
        if (handle == 2) {
          while ((o < mn || o > mx) && it < HB_BOUND_MAX_ITER) {
            if (o < mn) o = mn + fabs(mn - o);
            if (o > mx) o = mx - fabs(mx - o);
            ++it;
          }
        } else {
          while ((o < mn || o > mx) && it < HB_BOUND_MAX_ITER) {
            if (o < mn) o = mx - fabs(mn - o);
            if (o > mx) o = mn + fabs(mx - o);
            ++it;
          }
        }
        if (!(o < mn || o > mx) || pass == 1) break;
        // still outside after HB_BOUND_MAX_ITER passes: drop the whole periods, then let the loop finish
        const double period = (handle == 2) ? 2.0 * w : w;
        if (o > mx) o = mx + fmod(o - mx, period); else o = mn - fmod(mn - o, period);
      }
    }
  }
  return o;
}

// ---- sub-warp group primitives: G consecutive lanes (G a power of two <= 32) cooperate on one chain -------
// All take the group's lane mask so that groups of one warp may diverge from one another.
template <int G>
__device__ __forceinline__ unsigned hb_group_mask(int lane) {
  return (G == 32) ? 0xffffffffu : (((1u << (G & 31)) - 1u) << (lane & ~(G - 1)));
}
template <int G>
__device__ __forceinline__ double hb_group_sum(double v, unsigned gmask = 0xffffffffu) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(gmask, v, o);
  return v;
}
template <int G>
__device__ __forceinline__ int hb_group_sum_int(int v, unsigned gmask = 0xffffffffu) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(gmask, v, o);
  return v;
}
template <int G>
__device__ __forceinline__ unsigned hb_group_ballot(bool p, int lane, unsigned gmask = 0xffffffffu) {
  const unsigned b = __ballot_sync(gmask, p);
  if (G == 32) return b;
  return (b >> (lane & ~(G - 1))) & ((1u << (G & 31)) - 1u);
}

// ---- 1-D bulk TMA (cp.async.bulk) + mbarrier helpers ----------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// 1-D bulk TMA copy global -> shared, completion on an mbarrier (SASS: UBLKCP.S.G + SYNCS.*)
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// 1-D bulk TMA copy shared -> global (also peer-mapped global: NVLink), tracked by the issuing thread's bulk group
__device__ __forceinline__ void tma_store_1d(void* dst_gmem, const void* src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(dst_gmem), "r"(smem_u32(src_smem)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;\n" ::"n"(N) : "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait() { asm volatile("cp.async.bulk.wait_group %0;\n" ::"n"(N) : "memory"); }

// ---- fp64 tensor-core tile (mma.sync m8n8k4 f64 -> SASS DMMA) and the block-triangular layout of the t-distribution's
// W = chol(C)^-1 (hb_targets.cu: hb_tdist_dmma_kernel; hb_k1_propose.cu: the product phase of the fused small-population kernels) ----
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

__host__ __device__ constexpr int tdist_blk_base(int nb, int KB) {   // blocks of the n-blocks before nb
  int s = 0;
  for (int j = 0; j < nb; ++j) s += (2 * j + 2 < KB) ? 2 * j + 2 : KB;
  return s;
}
__host__ __device__ constexpr int tdist_n_blocks(int KB) { return tdist_blk_base((KB + 1) / 2, KB); }


__device__ __forceinline__ double hb_warp_sum(double v) { return hb_group_sum<32>(v); }
// two warp sums for the price of one: after the first exchange the lower half of the warp carries a, the upper half b;
// on return every lane holds both totals.  Fixed order (deterministic).
__device__ __forceinline__ void hb_warp_sum2(double& a, double& b, int lane) {
  const bool up = lane >= 16;
  const double give = up ? a : b;
  double v = (up ? b : a) + __shfl_xor_sync(0xffffffffu, give, 16);
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  a = __shfl_sync(0xffffffffu, v, 0);
  b = __shfl_sync(0xffffffffu, v, 16);
}
// ---- the structured evaluation of the reference's t_distribution (hb_targets.cu), shared by every kernel that computes it.
// The log-density must not depend on which kernel produced it (a shard of a multi-GPU run and the single-GPU run of the same
// chains may take different kernels and are compared bit for bit), so the order of the sums is fixed here: lane L of a warp
// accumulates the coordinate pairs k0 = 2 (32 m + L), k0 + 1 for m = 0, 1, ... (hb_tdist_acc, explicit IEEE operations: no
// contraction the compiler could apply in one kernel and not in another), then hb_warp_sum2, then hb_tdist_logdens.
__device__ __forceinline__ void hb_tdist_acc(double& s1, double& s2, double x0, double x1, double r0, double r1) {
  const double u0 = __dmul_rn(x0, r0), u1 = __dmul_rn(x1, r1);          // u = x / sqrt(k)  (callers pass x1 = r1 = 0 past d)
  s1 = __dadd_rn(s1, __dadd_rn(u0, u1));
  s2 = __fma_rn(u0, u0, s2); s2 = __fma_rn(u1, u1, s2);
}
// x' C^-1 x = 2 (s2 - s1^2 / (d + 1)) from s1 = sum u, s2 = sum u^2, then dmvt's log-density (R/test_t_distribution.R:43)
__device__ __forceinline__ double hb_tdist_logdens(double s1, double s2, int d, double konst, double df) {
  const double rss = __dmul_rn(2.0, __fma_rn(-__ddiv_rn(s1, (double)(d + 1)), s1, s2));
  return __fma_rn(-__dmul_rn(0.5, __dadd_rn(df, (double)d)), log1p(__ddiv_rn(rss, df)), konst);
}
#endif  // __CUDACC__
