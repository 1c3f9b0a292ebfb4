// hb_k1_propose.cu -- K1: differential-evolution proposal for every chain of the population.
//
// Replaces calc_prop (R/internals.R:115-219) + bound_par (:89-113) + prior_pdf (:49-69) as called for all nc chains
// from the generation-start state by dream_parallel (R/dream_parallel.R:269-275).
//
// Mapping: G = min(32, pow2 >= d) consecutive lanes own one chain; lane lg owns dimensions k = m*G + lg,
// m < ITER.  A warp therefore reads a partner row as ITER fully coalesced 256-byte segments, and the rows of the
// 1 + 2D chains it touches are the only HBM traffic (48d + 16 bytes per chain-generation).  Loads of
// dimensions outside the crossover subspace are predicated off.  Random draws are either read from the decision
// tape or generated with Philox4x32-10 keyed by (seed; chain, generation, slot): one Philox call per two dimensions
// (zt, lambda, zeta; slot map v2 in include/hydrobayes_b200_rng.h) plus six per chain, the latter computed by
// different lanes and broadcast.
#include <cstdio>
#include <cstdlib>
#include "hb_kernels.h"

namespace {

// error-free sum of up to HB_MAX_DELTA pair differences: reproduces R's long-double colSums (R/internals.R:196)
// to the last bit except for (rare) double-rounding cases
__device__ __forceinline__ void two_sum(double a, double b, double& s, double& e) {
  s = __dadd_rn(a, b);
  const double bb = __dadd_rn(s, -a);
  e = __dadd_rn(__dadd_rn(a, -__dadd_rn(s, -bb)), __dadd_rn(b, -bb));
}

// compensated accumulation / group reduction: the snooker projections are dot products that R sums in long double
// (R/internals.R:257, sum()); carrying the rounding errors reproduces that sum to the last bit in practice
__device__ __forceinline__ void comp_add(double& hi, double& lo, double v) {
  double e;
  two_sum(hi, v, hi, e);
  lo = __dadd_rn(lo, e);
}
template <int G>
__device__ __forceinline__ double comp_group_sum(double hi, double lo, unsigned gmask) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) {
    const double oh = __shfl_xor_sync(gmask, hi, o), ol = __shfl_xor_sync(gmask, lo, o);
    double e;
    two_sum(hi, oh, hi, e);
    lo = __dadd_rn(__dadd_rn(lo, ol), e);
  }
  return __dadd_rn(hi, lo);
}

// One chain's jump (calc_prop up to dx, R/internals.R:119-200) for the G lanes that own chain `base + lane / G`: the
// per-chain draws, the partner gather and dx.  Shared by the proposal kernel and the fused small-population kernel.
// population loads of the proposal: plain loads, or L2-coherent loads (ld.global.cg) inside the persistent kernel, where the
// population is rewritten by other CTAs of the SAME launch between generations (L1 / the read-only path could be stale)
template <bool CG>
__device__ __forceinline__ double k1_ld(const double* q) { return CG ? __ldcg(q) : *q; }

// The per-chain draws of a proposal (R/internals.R:119-146, 173, 192): everything calc_prop decides before it touches the
// population.  Split from the sweep so that the persistent kernel can compute generation i + 1's draws while the grid barrier
// of generation i is in flight.
struct k1_headv {
  double g_sn;
  int D, g_is_one, id;
  bool snooker, valid;
  long long pa[HB_MAX_DELTA], pb[HB_MAX_DELTA], ps2;   // a = first D partners, b = next D (R/internals.R:143-144); third snooker row
  long long jl, gl, rbase, rix, gch;
};

template <int G, bool TAPE>
__device__ __forceinline__ void k1_chain_head(const hb_k1_params& p, uint32_t gen, long long tape_g, long long base, int lane,
                                              k1_headv& hv, unsigned& errbits) {
  constexpr unsigned FULL = 0xffffffffu;
  const int lg = lane & (G - 1);
  const int gbase = lane & ~(G - 1);
  const unsigned gmask = hb_group_mask<G>(lane);
  const int nCR = p.nCR, delta = p.delta;
  const long long jl_raw = base + lane / G;
  const bool valid = jl_raw < p.n_local;
  const long long jl = valid ? jl_raw : p.n_local - 1;
  const long long gl = p.chain0 + jl * p.stride;   // chain index within the handle (run-major)
  const long long run = p.n_runs > 1 ? gl / p.nc : 0;
  const long long j = gl - run * p.nc;         // chain within the run
  const long long rbase = gl - j;              // first chain of the run
  const long long gch = p.gchain0 + gl;        // global chain id: Philox counter
  const long long rix = tape_g * p.tape.nct + gl;   // tape column

  // ---------------- per-chain draws: snooker test, D, partners, crossover id, gamma choice ----------------
  double u_sn, g_sn = 0.0;
  int D, g_is_one, id;
  long long pa[HB_MAX_DELTA], pb[HB_MAX_DELTA];   // a = first D partners, b = next D (R/internals.R:143-144)
  long long ps2 = 0;                               // third snooker row (:127)
  const hb_phx ph(p.seed, (uint64_t)gch, gen);
  if (TAPE) {
    u_sn = p.tape.u_snooker ? p.tape.u_snooker[rix] : 1.0;            // R/internals.R:119
    D = p.tape.D[rix];                                               // :123
    id = p.tape.id[rix];                                             // :173
    g_is_one = p.tape.g_is_one[rix];                                 // :192
    if (p.tape.g_snooker) g_sn = p.tape.g_snooker[rix];              // :156
    if (D < 1 || D > delta || id >= nCR) { errbits |= HB_FLAG_BAD_TAPE; D = 1; id = 0; }
    const int32_t* pr = p.tape.partners + rix * (2 * delta);
#pragma unroll
    for (int q = 0; q < HB_MAX_DELTA; ++q) {
      pa[q] = (q < D) ? pr[q] : 0;
      pb[q] = (q < D) ? pr[D + q] : 0;
    }
    if (p.past_sample && u_sn < p.psnooker) { pa[0] = pr[0]; pb[0] = pr[1]; ps2 = pr[2]; }
  } else {
    uint4 hw[6];
#pragma unroll
    for (int s0 = 0; s0 < 6; s0 += G) {
      uint4 w = make_uint4(0, 0, 0, 0);
      const int s = s0 + lg;
      if (s < 6) w = ph.slot((uint32_t)s);
#pragma unroll
      for (int q = 0; q < G && s0 + q < 6; ++q) {
        if (G == 1) { hw[s0 + q] = w; }
        else {
          hw[s0 + q].x = __shfl_sync(FULL, w.x, gbase + q);
          hw[s0 + q].y = __shfl_sync(FULL, w.y, gbase + q);
          hw[s0 + q].z = __shfl_sync(FULL, w.z, gbase + q);
          hw[s0 + q].w = __shfl_sync(FULL, w.w, gbase + q);
        }
      }
    }
    u_sn = hb_u53(hb_pair64(hw[HB_SLOT_HEAD], 0));
    D = 1 + (int)hb_mulhi64(hb_pair64(hw[HB_SLOT_HEAD], 1), (uint64_t)delta);
    const double u_id = hb_u53(hb_pair64(hw[HB_SLOT_CHOICE], 0));
    double cum = 0.0;
    id = nCR - 1;
    bool found = false;
    for (int q = 0; q < nCR; ++q) {
      cum += p.ctrl[run].pCR[q];
      if (!found && u_id < cum) { id = q; found = true; }
    }
    g_is_one = hb_u53(hb_pair64(hw[HB_SLOT_CHOICE], 1)) >= 1.0 - p.p_g;
    const bool sn = p.past_sample && (u_sn < p.psnooker);
    if (sn) g_sn = 1.2 + hb_u53(hb_pair64(ph.slot(HB_SLOT_ACCEPT), 1));
    const int n_part = sn ? 3 : 2 * D;
    hb_fy fy;
    long long n_rem = p.past_sample ? p.m_arch : p.nc - 1;
    long long y[2 * HB_MAX_DELTA];
#pragma unroll
    for (int q = 0; q < 2 * HB_MAX_DELTA; ++q) {
      y[q] = 0;
      if (q < n_part) {
        long long v = fy.draw(hb_pair64(hw[HB_SLOT_PARTNER0 + (q >> 1)], q & 1), n_rem);
        if (!p.past_sample && v >= j) v += 1;                        // sample((1:nc)[-j], ...)  :142
        y[q] = v;
      }
    }
#pragma unroll
    for (int q = 0; q < HB_MAX_DELTA; ++q) {
      pa[q] = y[q];
      pb[q] = (D == 1) ? y[(1 + q) & 7] : (D == 2) ? y[(2 + q) & 7] : (D == 3) ? y[(3 + q) & 7] : y[(4 + q) & 7];
    }
    if (sn) { pb[0] = y[1]; ps2 = y[2]; }
  }
  const bool snooker = p.past_sample && (u_sn < p.psnooker);
  {
    const long long n_src = p.past_sample ? p.m_arch : p.nc;
#pragma unroll
    for (int q = 0; q < HB_MAX_DELTA; ++q) {
      const bool used = snooker ? (q == 0) : (q < D);
      if (used && (pa[q] < 0 || pa[q] >= n_src || pb[q] < 0 || pb[q] >= n_src)) {
        errbits |= HB_FLAG_BAD_TAPE; pa[q] = 0; pb[q] = 0;
      }
    }
    if (snooker && (ps2 < 0 || ps2 >= n_src)) { errbits |= HB_FLAG_BAD_TAPE; ps2 = 0; }
  }

  hv.g_sn = g_sn; hv.D = D; hv.g_is_one = g_is_one; hv.id = id; hv.snooker = snooker; hv.valid = valid; hv.ps2 = ps2;
#pragma unroll
  for (int q = 0; q < HB_MAX_DELTA; ++q) { hv.pa[q] = pa[q]; hv.pb[q] = pb[q]; }
  hv.jl = jl; hv.gl = gl; hv.rbase = rbase; hv.rix = rix; hv.gch = gch;
}

// The population-independent part of the parallel direction update (R/internals.R:173-192): the crossover subspace
// A = {k : zt_k < CR[id]} (or which.min), d*, gamma, and the Philox words behind lambda / zeta.  Split from the sweep so that
// the persistent kernel can run it for generation i + 1 while the grid barrier of generation i is in flight.
template <int ITER>
struct k1_prev {
  unsigned selmask[ITER];
  uint32_t wl[ITER], wz[ITER];     // slot map v2: lo word >> 16 -> lambda; hi word -> zeta
  int d_star, kmin;
  double gam;
};
template <int G, int ITER, bool TAPE>
__device__ __forceinline__ void k1_chain_pre(const hb_k1_params& p, uint32_t gen, const k1_headv& hv, int lane, k1_prev<ITER>& pv) {
  const int lg = lane & (G - 1);
  const unsigned gmask = hb_group_mask<G>(lane);
  const int d = p.d, nCR = p.nCR;
  const int D = hv.D, id = hv.id;
  const long long rix = hv.rix;
  const hb_phx ph(p.seed, (uint64_t)hv.gch, gen);
  const double CRv = (double)(id + 1) / (double)nCR;               // CR <- (1:nCR)/nCR
  double ztv[ITER];
  int d_star = 0;
#pragma unroll
  for (int m = 0; m < ITER; ++m) {
    const int k = m * G + lg;
    ztv[m] = 2.0; pv.wl[m] = 0; pv.wz[m] = 0;
    if (k < d) {
      if (TAPE) ztv[m] = p.tape.zt[rix * d + k];                   // :175
      else {
        const uint4 w = ph.slot(HB_SLOT_DIM0 + (uint32_t)(k >> 1));
        const uint32_t lo = (k & 1) ? w.z : w.x;
        ztv[m] = hb_rng_u16(lo); pv.wl[m] = lo >> 16; pv.wz[m] = (k & 1) ? w.w : w.y;
      }
    }
    pv.selmask[m] = hb_group_ballot<G>(k < d && ztv[m] < CRv, lane, gmask); // A <- which(zt < CR[id])  :177
    d_star += __popc(pv.selmask[m]);
  }
  int kmin = -1;
  if (d_star == 0) {                             // :181-184  A <- which.min(zt)
    double bv = 3.0; int bk = 0x7fffffff;
#pragma unroll
    for (int m = 0; m < ITER; ++m) {
      const int k = m * G + lg;
      if (k < d && ztv[m] < bv) { bv = ztv[m]; bk = k; }
    }
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) {
      const double ov = __shfl_xor_sync(gmask, bv, o);
      const int ok = __shfl_xor_sync(gmask, bk, o);
      if (ov < bv || (ov == bv && ok < bk)) { bv = ov; bk = ok; }
    }
    kmin = bk; d_star = 1;
  }
  // :190  2.38/sqrt(2 D d*): the host-built table when there is one (same IEEE sqrt and division: bit-identical), which
  // spares the warp ~40 fp64 instructions
  const double gamma_d = p.gamma_tab ? __ldg(p.gamma_tab + (D - 1) * d + (d_star - 1)) : 2.38 / sqrt(2.0 * (double)D * (double)d_star);
  pv.gam = hv.g_is_one ? 1.0 : gamma_d;                            // :192
  pv.d_star = d_star; pv.kmin = kmin;
}

// The sweep of a proposal (R/internals.R:150-200): partner rows, crossover subspace, dx.
// STAGE (one chain per warp, G == 32): the rows the chain needs -- its own state and the 2D partner rows -- are fetched into
// a per-warp shared-memory stage by 1-D bulk TMA copies first, the Philox work of the crossover pass runs while they are in
// flight, and the sweep reads shared memory.
// stage: [1 + 2 delta (>= 4)][ldx] doubles, rows 0 = x, 1.. = a, 1 + delta.. = b (snooker: 1, 2, 3); sbar: its mbarrier;
// sphase: completions so far (parity of the next wait).  Bulk copies read through L2, so they are coherent with rows other
// CTAs of a persistent kernel wrote before the last grid barrier.
template <int G, int ITER, bool TAPE, bool CG = false, bool STAGE = false>
__device__ __forceinline__ void k1_chain_body(const hb_k1_params& p, uint32_t gen, const k1_headv& hv, int lane, double (&xk)[ITER],
                                              double (&dxk)[ITER], unsigned& errbits,
                                              double* stage = nullptr, uint64_t* sbar = nullptr, unsigned* sphase = nullptr,
                                              const double* Xcur = nullptr, const k1_prev<ITER>* pre = nullptr) {
  static_assert(!STAGE || G == 32, "row staging is written for one chain per warp");
  const int lg = lane & (G - 1);
  const unsigned gmask = hb_group_mask<G>(lane);
  const int d = p.d, ldx = p.ldx, nCR = p.nCR, delta = p.delta;
  const double* Xpop = Xcur ? Xcur : p.X;   // the persistent kernel alternates between two population buffers
  const double* src = p.past_sample ? p.Z : Xpop;
  const double g_sn = hv.g_sn;
  const int D = hv.D, g_is_one = hv.g_is_one, id = hv.id;
  const bool snooker = hv.snooker;
  const long long ps2 = hv.ps2, gl = hv.gl, rbase = hv.rbase, rix = hv.rix;
  long long pa[HB_MAX_DELTA], pb[HB_MAX_DELTA];
#pragma unroll
  for (int q = 0; q < HB_MAX_DELTA; ++q) { pa[q] = hv.pa[q]; pb[q] = hv.pb[q]; }
  const hb_phx ph(p.seed, (uint64_t)hv.gch, gen);
  const double* xrow = Xpop + gl * (long long)ldx;
  const double* srcr = p.past_sample ? src : src + rbase * (long long)ldx;   // partner rows of this run
  // row q of the a / b partner sets (global memory, or the staged copy)
  auto row_a = [&](int q) -> const double* { return STAGE ? stage + (size_t)(1 + q) * ldx : srcr + pa[q] * (long long)ldx; };
  auto row_b = [&](int q) -> const double* { return STAGE ? stage + (size_t)(1 + delta + q) * ldx : srcr + pb[q] * (long long)ldx; };
  if (STAGE) {
    const uint32_t row_bytes = (uint32_t)ldx * 8u;
    const int n_rows = snooker ? 4 : 1 + 2 * D;
    __syncwarp();
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // the previous chain's reads of the stage precede these copies
    if (lane == 0) mbar_expect_tx(sbar, (uint32_t)n_rows * row_bytes);
    __syncwarp();
    if (lane < n_rows) {
      const double* g = xrow;
      int srow = 0;
      if (snooker) {
        if (lane == 1) g = srcr + pa[0] * (long long)ldx;
        if (lane == 2) g = srcr + pb[0] * (long long)ldx;
        if (lane == 3) g = srcr + ps2 * (long long)ldx;
        srow = lane;
      } else {
#pragma unroll
        for (int q = 0; q < HB_MAX_DELTA; ++q) {
          if (q < D && lane == 1 + q) { g = srcr + pa[q] * (long long)ldx; srow = 1 + q; }
          if (q < D && lane == 1 + D + q) { g = srcr + pb[q] * (long long)ldx; srow = 1 + delta + q; }
        }
      }
      tma_load_1d(stage + (size_t)srow * ldx, g, row_bytes, sbar);
    }
    xrow = stage;
  }
  auto rows_ready = [&]() { if (STAGE) { mbar_wait(sbar, *sphase & 1u); *sphase += 1u; } };

  if (snooker) {
    // ---------------- snooker update, R/internals.R:150-163 with orth_proj :250-263 ----------------
    const double* r1 = STAGE ? stage + (size_t)1 * ldx : srcr + pa[0] * (long long)ldx;
    const double* r2 = STAGE ? stage + (size_t)2 * ldx : srcr + pb[0] * (long long)ldx;
    const double* r3 = STAGE ? stage + (size_t)3 * ldx : srcr + ps2 * (long long)ldx;
    rows_ready();
    double uu[ITER], a2[ITER], a3[ITER];
    double s2 = 0, s3 = 0, su = 0, l2 = 0, l3 = 0, lu = 0;
    int differs = 0;
#pragma unroll
    for (int m = 0; m < ITER; ++m) {
      const int k = m * G + lg;
      xk[m] = 0; uu[m] = 0; a2[m] = 0; a3[m] = 0;
      if (k < d) {
        xk[m] = k1_ld<CG && !STAGE>(xrow + k);
        const double b = k1_ld<CG && !STAGE>(r1 + k);
        a2[m] = k1_ld<CG && !STAGE>(r2 + k); a3[m] = k1_ld<CG && !STAGE>(r3 + k);
        uu[m] = __dadd_rn(xk[m], -b);
        differs |= (xk[m] != b);
        comp_add(s2, l2, __dmul_rn(__dadd_rn(a2[m], -xk[m]), uu[m]));
        comp_add(s3, l3, __dmul_rn(__dadd_rn(a3[m], -xk[m]), uu[m]));
        comp_add(su, lu, __dmul_rn(uu[m], uu[m]));
      }
    }
    s2 = comp_group_sum<G>(s2, l2, gmask); s3 = comp_group_sum<G>(s3, l3, gmask); su = comp_group_sum<G>(su, lu, gmask);
    differs = hb_group_sum_int<G>(differs, gmask);
#pragma unroll
    for (int m = 0; m < ITER; ++m) {
      const int k = m * G + lg;
      dxk[m] = 0;
      if (k < d) {
        double zeta;
        if (TAPE) zeta = p.tape.zeta[rix * d + k];
        else { const uint4 w = ph.slot(HB_SLOT_DIM0 + (uint32_t)(k >> 1)); zeta = hb_zeta32((k & 1) ? w.w : w.y, p.c_star); }
        double rp2 = a2[m], rp3 = a3[m];
        if (differs) {
          rp2 = __dadd_rn(xk[m], __ddiv_rn(__dmul_rn(uu[m], s2), su));   // q <- a + u*sum((p-a)*u)/sum(u*u)
          rp3 = __dadd_rn(xk[m], __ddiv_rn(__dmul_rn(uu[m], s3), su));
        }
        if (!isfinite(rp2) || !isfinite(rp3)) errbits |= HB_FLAG_PROP_NONFINITE;   // :260
        dxk[m] = __dadd_rn(zeta, __dmul_rn(g_sn, __dadd_rn(rp2, -rp3)));          // :163
      }
    }
  } else {
    // ---------------- parallel direction update, R/internals.R:166-200 ----------------
    k1_prev<ITER> pv_local;
    if (!pre) k1_chain_pre<G, ITER, TAPE>(p, gen, hv, lane, pv_local);
    const k1_prev<ITER>& pv = pre ? *pre : pv_local;
    const int kmin = pv.kmin;
    const double gam = pv.gam;
    rows_ready();
    int rank_base = 0;
#pragma unroll
    for (int m = 0; m < ITER; ++m) {
      const int k = m * G + lg;
      const bool sel = (kmin >= 0) ? (k == kmin) : ((pv.selmask[m] >> lg) & 1u);
      xk[m] = 0; dxk[m] = 0;
      if (k < d) {
        xk[m] = k1_ld<CG && !STAGE>(xrow + k);
        if (sel) {
          double lambda, zeta;
          if (TAPE) {
            const int rank = (kmin >= 0) ? 0 : rank_base + __popc(pv.selmask[m] & ((1u << lg) - 1u));
            lambda = p.tape.lambda[rix * d + rank];                  // :188 (by rank within A)
            zeta = p.tape.zeta[rix * d + rank];                      // :194
          } else {
            lambda = hb_rng_lambda16(pv.wl[m], p.c_val);
            zeta = hb_zeta32(pv.wz[m], p.c_star);
          }
          // jump_diff <- colSums(r1[,A] - r2[,A])   :196
          double s = 0.0, comp = 0.0;
#pragma unroll
          for (int q = 0; q < HB_MAX_DELTA; ++q) {
            if (q < D) {
              const double a = k1_ld<CG && !STAGE>(row_a(q) + k);
              const double b = k1_ld<CG && !STAGE>(row_b(q) + k);
              const double t = __dadd_rn(a, -b);
              if (q == 0) s = t;
              else { double e; two_sum(s, t, s, e); comp = __dadd_rn(comp, e); }
            }
          }
          const double jd = __dadd_rn(s, comp);
          // dx[A] <- (zeta + (1+lambda)*g*jump_diff) * beta0   :198-200
          const double v = __dadd_rn(zeta, __dmul_rn(__dmul_rn(__dadd_rn(1.0, lambda), gam), jd));
          dxk[m] = __dmul_rn(v, p.beta0);
        }
      }
      rank_base += __popc(pv.selmask[m]);
    }
  }

}


template <int G, int ITER, bool TAPE, bool CG = false, bool STAGE = false>
__device__ __forceinline__ void k1_chain_jump(const hb_k1_params& p, long long base, int lane, double (&xk)[ITER],
                                              double (&dxk)[ITER], bool& valid, long long& jl, int& id, unsigned& errbits,
                                              double* stage = nullptr, uint64_t* sbar = nullptr, unsigned* sphase = nullptr) {
  k1_headv hv;
  k1_chain_head<G, TAPE>(p, p.gen, p.tape_g, base, lane, hv, errbits);
  k1_chain_body<G, ITER, TAPE, CG, STAGE>(p, p.gen, hv, lane, xk, dxk, errbits, stage, sbar, sphase);
  valid = hv.valid; jl = hv.jl; id = hv.id;
}

template <int G, int ITER, bool TAPE>
__global__ void __launch_bounds__(256) hb_k1_kernel(const hb_k1_params p) {
  constexpr int CPW = 32 / G;
  const int lane = threadIdx.x & 31;
  const int lg = lane & (G - 1);
  const unsigned gmask = hb_group_mask<G>(lane);
  const long long warp_id = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int d = p.d, ldx = p.ldx;
  unsigned errbits = 0;

  for (long long base = warp_id * CPW; base < p.n_local; base += n_warps * CPW) {
    double xk[ITER], dxk[ITER];
    double lp_part = 0.0;
    bool valid; long long jl; int id;
    k1_chain_jump<G, ITER, TAPE>(p, base, lane, xk, dxk, valid, jl, id, errbits);

    // ---------------- xp <- x + dx ; finite check ; bounds ; prior ----------------
#pragma unroll
    for (int m = 0; m < ITER; ++m) {
      const int k = m * G + lg;
      if (k < d) {
        double xp = __dadd_rn(xk[m], dxk[m]);                          // :206
        if (!isfinite(xp)) errbits |= HB_FLAG_PROP_NONFINITE;          // :209
        if (p.bound) xp = hb_bound_par(xp, p.pmin[k], p.pmax[k], p.bound);   // :210-214
        if (p.prior == 1) {                                            // dunif(log = TRUE), :57
          const double a = p.pmin[k], b = p.pmax[k];
          lp_part += (a <= xp && xp <= b) ? -log(b - a) : -INFINITY;
        }
        if (valid) p.XP[jl * (long long)ldx + k] = xp;
      }
    }
    if (p.prior == 1) lp_part = hb_group_sum<G>(lp_part, gmask);
    if (valid && lg == 0) {
      double lp = (p.prior == 1) ? lp_part : 0.0;                      // flat prior: 0  :55
      if (!(lp >= HB_FLOOR)) lp = (lp != lp) ? lp : HB_FLOOR;          // :65
      if (!isfinite(lp)) errbits |= HB_FLAG_LP_NONFINITE;              // :67
      p.lp_xp[jl] = lp;
      p.id_out[jl] = id;
    }
  }
  if (errbits) atomicOr(p.err, errbits);
}


// what changes from one generation to the next in the fused kernels (kept in registers; everything else is read from the
// kernel parameters in the constant bank)
struct fused_gen {
  const double* Xin; double* Xout;          // population read by the proposals / written by the update
  double* hist_x; double* hist_l; double* hist_fx; double* lpost_hist_row; uint8_t* accept_rec;
  const double* u_accept;                    // tape row (replay) or nullptr
  uint32_t gen;
};
__device__ __forceinline__ fused_gen fused_gen_of(const hb_k3_params& q) {
  fused_gen g;
  g.Xin = q.X; g.Xout = q.Xout; g.hist_x = q.hist_x; g.hist_l = q.hist_l; g.hist_fx = q.hist_fx;
  g.lpost_hist_row = q.lpost_hist_row; g.accept_rec = q.accept_rec; g.u_accept = q.u_accept; g.gen = q.gen;
  return g;
}

// One chain's whole generation for the warp that owns it: proposal, t-distribution log-density, Metropolis decision, state
// update into q.Xout and history store.  Shared by the one-launch-per-generation kernel and the persistent kernel below.
// st (adaptation period only, else nullptr): this warp's three statistics rows [3][ITER * 32] in shared memory.
template <int ITER, bool TAPE, bool CG = false>
__device__ __forceinline__ void fused_chain_step(const hb_k1_params& p, const hb_k3_params& q, const hb_fused_tdist& td, const fused_gen& fg,
                                                 const double* __restrict__ rs_s, double* __restrict__ xrow_s, const k1_headv& hv, int lane,
                                                 unsigned long long* s_cnt,
                                                 unsigned& errbits, double* stage, uint64_t* sbar, unsigned& sphase, bool has_chain,
                                                 long long* tm = nullptr, const double* shift_g = nullptr, double* st = nullptr,
                                                 int* s_id = nullptr, const double* logu_pre = nullptr,
                                                 const k1_prev<ITER>* sweep_pre = nullptr) {
  constexpr unsigned FULL = 0xffffffffu;
  constexpr int SLD = ITER * 32;
  const int d = p.d, ldx = p.ldx;
    if (!has_chain) {                                                   // no chain: no contribution to the statistics
      if (st) {
#pragma unroll
        for (int m = 0; m < ITER; ++m) { st[m * 32 + lane] = 0.0; st[SLD + m * 32 + lane] = 0.0; st[2 * SLD + m * 32 + lane] = 0.0; }
        if (lane == 0) *s_id = -1;
      }
      return;
    }
    // ---------------- proposal (K1) ----------------
    double xk[ITER], dxk[ITER], xp[ITER];
    const long long jl = has_chain ? hv.jl : 0; const int id = has_chain ? hv.id : 0;
    // the chain's current log posterior: fetched now, needed by the Metropolis decision at the end (an L2 round trip)
    double lpost_early = 0.0;
    if (lane == 0 && has_chain) lpost_early = q.lpost[q.chain0 + jl * q.stride - q.state0];
    if (tm) tm[0] = clock64();
#pragma unroll
    for (int m = 0; m < ITER; ++m) { xk[m] = 0.0; dxk[m] = 0.0; }
    if (has_chain) k1_chain_body<32, ITER, TAPE, CG, true>(p, fg.gen, hv, lane, xk, dxk, errbits, stage, sbar, &sphase, fg.Xin, sweep_pre);
    if (tm) tm[1] = clock64();
    double lp_part = 0.0;
#pragma unroll
    for (int m = 0; m < ITER; ++m) {
      const int k = m * 32 + lane;
      xp[m] = 0.0;
      if (k < d && has_chain) {
        double v = __dadd_rn(xk[m], dxk[m]);                            // R/internals.R:206
        if (!isfinite(v)) errbits |= HB_FLAG_PROP_NONFINITE;            // :209
        if (p.bound) v = hb_bound_par(v, p.pmin[k], p.pmax[k], p.bound);   // :210-214
        if (p.prior == 1) {                                             // dunif(log = TRUE), :57
          const double a = p.pmin[k], b = p.pmax[k];
          lp_part += (a <= v && v <= b) ? -log(b - a) : -INFINITY;
        }
        xp[m] = v;
      }
    }
    if (p.prior == 1) lp_part = hb_group_sum<32>(lp_part);
    double lpx = (p.prior == 1) ? lp_part : 0.0;                        // flat prior: 0  :55
    if (!(lpx >= HB_FLOOR)) lpx = (lpx != lpx) ? lpx : HB_FLOOR;        // :65
    if (!isfinite(lpx)) errbits |= HB_FLAG_LP_NONFINITE;                // :67

    // ---------------- log-density (K2): dmvt, R/test_t_distribution.R:43 ----------------
    // the structured evaluation of the target (hb_targets.cu: hb_tdist_struct_kernel) on the proposal in registers:
    // x' C^-1 x = 2 (sum u^2 - (sum u)^2 / (d + 1)), u_k = xp_k / sqrt(k); rs_s: 1 / sqrt(k) in shared memory, zero past d
    // (the sums run in the fixed order of hb_tdist_acc -- lane L takes the coordinate pairs 2 (32 m + L) -- so the proposal goes
    // through this warp's scratch row once)
    if (tm) tm[2] = clock64();
    double dens;
    {
#pragma unroll
      for (int m = 0; m < ITER; ++m) xrow_s[m * 32 + lane] = xp[m];       // zeros past d
      __syncwarp();
      double s1 = 0.0, s2 = 0.0;
#pragma unroll
      for (int m = 0; m < (ITER + 1) / 2; ++m) {
        const int k0 = 2 * (m * 32 + lane);
        if (k0 < ITER * 32) {
          const double2 x2 = *reinterpret_cast<const double2*>(xrow_s + k0);
          const double2 r2 = *reinterpret_cast<const double2*>(rs_s + k0);
          if (k0 < d) hb_tdist_acc(s1, s2, x2.x, x2.y, r2.x, r2.y);
        }
      }
      __syncwarp();
      hb_warp_sum2(s1, s2, lane);
      dens = hb_tdist_logdens(s1, s2, d, td.konst, td.df);
    }
    if (tm) tm[7] = clock64();
    if (tm) tm[3] = clock64();

    // ---------------- accept / update / store (K3, outside the adaptation period) ----------------
    const long long j = q.chain0 + jl * q.stride;
    const long long si = j - q.state0;
    const long long gch = q.gchain0 + j;
    int accf = 0;
    double lln = 0, lpostx = 0, lpost_old = 0;
    if (lane == 0) {
      double llr = (td.lik == 1) ? log(dens) : dens;                    // calc_ll, R/internals.R:4-7
      if (!(llr >= HB_FLOOR)) llr = (llr != llr) ? llr : HB_FLOOR;      // :19
      if (!isfinite(llr)) errbits |= HB_FLAG_LL_NONFINITE;              // :21
      lln = llr;
      lpostx = lpx + lln;                                               // R/dream_parallel.R:291
      lpost_old = lpost_early;
      double logu;
      if (TAPE) logu = log(fg.u_accept[j]);
      else if (logu_pre) logu = *logu_pre;                              // the persistent kernel draws it while the grid barrier is in flight
      else { const hb_phx ph(q.seed, (uint64_t)gch, fg.gen); logu = log(hb_u53(hb_pair64(ph.slot(HB_SLOT_ACCEPT), 0))); }
      const double p_acc = fmin(fmax(-100.0, lpostx - lpost_old), 0.0); // R/internals.R:239
      accf = p_acc > logu;                                              // :241
    }
    accf = __shfl_sync(FULL, accf, 0);
#pragma unroll
    for (int m = 0; m < ITER; ++m) {
      const int k = m * 32 + lane;
      if (k < d) {
        const double xn = accf ? xp[m] : xk[m];                         // R/dream_parallel.R:358
        fg.Xout[j * (long long)ldx + k] = xn;
        if (fg.hist_x) fg.hist_x[si * (long long)d + k] = xn;             // chain[ind_out, 1:d, ] :388
      }
    }
    if (lane == 0) {
      double lp_c, ll_c, lpost_c;
      if (accf) {
        q.lp[si] = lpx; q.ll[si] = lln; q.lpost[si] = lpostx;           // :359-361
        if (q.fx) q.fx[si] = dens;                                      // :362
        lp_c = lpx; ll_c = lln; lpost_c = lpostx;
      } else {
        lpost_c = lpost_old;                                            // :365
        lp_c = 0; ll_c = 0;
        if (fg.hist_l) { lp_c = q.lp[si]; ll_c = q.ll[si]; }
      }
      if (fg.hist_l) { fg.hist_l[si * 3 + 0] = lp_c; fg.hist_l[si * 3 + 1] = ll_c; fg.hist_l[si * 3 + 2] = lpost_c; }   // :389-391
      if (fg.hist_fx && q.fx) fg.hist_fx[si] = q.fx[si];                  // :394
      if (fg.lpost_hist_row) fg.lpost_hist_row[si] = lpost_c;
      if (fg.accept_rec) fg.accept_rec[si] = (uint8_t)accf;
      atomicAdd(&s_cnt[1 + id], 1ull);                                  // n_id  :368-369
      if (accf) { atomicAdd(&s_cnt[0], 1ull); atomicAdd(&s_cnt[1 + HB_MAX_NCR + id], 1ull); }   // n_acc :374
    }
    if (st) {
      // adaptation period: this chain's terms of the std_x / jump statistics (R/dream_parallel.R:376-378; the sums of
      // hb_k3_kernel: shifted post-update state, its square, dx^2 filed under the crossover index); the CTA folds the rows of
      // its warps in warp order after the next block barrier
#pragma unroll
      for (int m = 0; m < ITER; ++m) {
        const int k = m * 32 + lane;
        double cc = 0.0, dd = 0.0;
        if (k < d) {
          const double xn = accf ? xp[m] : xk[m];
          cc = xn - __ldcg(shift_g + k);
          const double dxv = accf ? xp[m] - xk[m] : 0.0;                // :357 (after bound handling), 0 if rejected
          dd = dxv * dxv;
        }
        st[k] = cc; st[SLD + k] = cc * cc; st[2 * SLD + k] = dd;
      }
      if (lane == 0) *s_id = id;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Fused generation for SMALL populations (BASELINE C2: nc = 1024, d = 100): proposal + t-distribution log-density +
// Metropolis accept/update/store in ONE launch, one warp per chain.  At this size the three kernels of the large-population
// path take a few microseconds each and the generation time is their launch gaps; fusing them removes two launches and the
// XP / ll round trips through L2.  The synchronous sweep (R/dream_parallel.R:269, all proposals from the generation-start
// state) is kept without a grid-wide barrier by double buffering the population: every chain reads partners from X and
// writes its post-generation row into Xout (its proposal if accepted, its old row otherwise); the host swaps the buffers.
// This kernel runs outside the adaptation period only (the persistent kernel below also covers the adaptation period).
//   proposal      k1_chain_head + k1_chain_body -- the same code, draw for draw, as hb_k1_kernel<32, ITER>
//   log-density   the structured evaluation of the built-in t_distribution (hb_targets.cu) on the proposal in registers,
//                 summed in the fixed order of hb_tdist_acc: bit-identical to the plugin's value
//   accept        the non-adaptation path of hb_k3_kernel (R/internals.R:239-241, R/dream_parallel.R:348-366, 386-394)
// ------------------------------------------------------------------------------------------------------------------
template <int ITER, bool TAPE>
__global__ void __launch_bounds__(256) hb_fused_small_kernel(const hb_k1_params p, const hb_k3_params q, const hb_fused_tdist td) {
  extern __shared__ __align__(16) double fused_smem[];               // 1 / sqrt(k) [ITER * 32] | row stages [warps][rows][ldx]
  __shared__ unsigned long long s_cnt[2 * HB_MAX_NCR + 1];
  __shared__ __align__(8) uint64_t sbars[8];
  const int lane = threadIdx.x & 31;
  const int n_wib = blockDim.x >> 5, wib = threadIdx.x >> 5;
  const int stage_rows = (1 + 2 * p.delta) > 4 ? (1 + 2 * p.delta) : 4;
  double* rs_s = fused_smem;
  double* xrow_s = fused_smem + ITER * 32 * (1 + wib);              // this warp's scratch row
  double* stage = fused_smem + ITER * 32 * (1 + n_wib) + (size_t)wib * stage_rows * p.ldx;
  unsigned sphase = 0;
  if (lane == 0) { mbar_init(&sbars[wib], 1); asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
  for (int k = threadIdx.x; k < ITER * 32; k += blockDim.x) rs_s[k] = (k < p.d) ? td.rs[k] : 0.0;
  const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
  unsigned errbits = 0;
  if (threadIdx.x <= 2 * HB_MAX_NCR) s_cnt[threadIdx.x] = 0ull;
  __syncthreads();

  for (long long base = (long long)blockIdx.x * n_wib + wib; base < p.n_local; base += n_warps) {   // one chain per warp and round
    k1_headv hv;
    k1_chain_head<32, TAPE>(p, p.gen, p.tape_g, base, lane, hv, errbits);
    fused_chain_step<ITER, TAPE>(p, q, td, fused_gen_of(q), rs_s, xrow_s, hv, lane, s_cnt, errbits, stage, &sbars[wib], sphase, true);
  }
  __syncthreads();
  if (threadIdx.x <= 2 * HB_MAX_NCR) {
    const unsigned long long v = s_cnt[threadIdx.x];
    if (v) {
      if (threadIdx.x == 0) atomicAdd(&q.ctrl->n_acc_gen, v);
      else if (threadIdx.x <= HB_MAX_NCR) atomicAdd(&q.ctrl->n_id_gen[threadIdx.x - 1], v);
      else atomicAdd(&q.ctrl->n_acc_id_gen[threadIdx.x - 1 - HB_MAX_NCR], v);
    }
  }
  if (errbits) atomicOr(&q.ctrl->err, errbits);
}

template <int ITER>
cudaError_t launch_fused(const hb_k1_params& p, const hb_k3_params& q, const hb_fused_tdist& td, bool tape, int sm_count,
                         cudaStream_t st) {
  const int threads = 256, warps = threads / 32;
  long long blocks = (p.n_local + warps - 1) / warps;
  const long long cap = (long long)sm_count * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  const int stage_rows = (1 + 2 * p.delta) > 4 ? (1 + 2 * p.delta) : 4;
  if (!td.rs || p.d > ITER * 32 || (p.ldx & 1)) return cudaErrorInvalidValue;
  const size_t smem = ((size_t)ITER * 32 * (1 + warps) + (size_t)warps * stage_rows * p.ldx) * sizeof(double);
  if (smem > 220 * 1024) return cudaErrorInvalidValue;
  static bool attr_set[2] = {false, false};
  if (!attr_set[tape]) {
    cudaError_t e = tape ? cudaFuncSetAttribute(hb_fused_small_kernel<ITER, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024)
                         : cudaFuncSetAttribute(hb_fused_small_kernel<ITER, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    if (e != cudaSuccess) return e;
    attr_set[tape] = true;
  }
  if (tape) hb_fused_small_kernel<ITER, true><<<(unsigned)blocks, threads, smem, st>>>(p, q, td);
  else hb_fused_small_kernel<ITER, false><<<(unsigned)blocks, threads, smem, st>>>(p, q, td);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------------------------
// Persistent generation loop for small populations (BASELINE C2 is latency bound: 20 000 strictly dependent generations of
// 1 024 chains).  ONE cooperative launch runs all generations of a slice that lie behind the adaptation
// period -- or, with pp.adapt, the generations of the adaptation period up to the next updateInterval boundary: one warp per
// chain, the population double buffered (generation i reads buffer A and writes B, i + 1 the other way round), and a
// grid-wide barrier per generation in place of a kernel boundary -- every CTA arrives on a counter, then works out what
// generation i + 1 needs that does not depend on the population (the per-chain draws, the crossover subspace, the logarithm
// of the Metropolis uniform) while the others arrive; outside the adaptation period CTA 0 folds the AR / CR bookkeeping at
// updateInterval boundaries (R/dream_parallel.R:425-433; the non-adaptation part of hb_finalize_kernel) before it releases
// the others; inside it every CTA parks its std_x / jump statistic sums of the generation in a slot and the finalize kernels
// fold the slots after the launch.  The per-chain step is fused_chain_step, i.e. draw for draw and operation for operation the
// one-launch-per-generation kernel; the population is read with L2-coherent loads (other CTAs rewrite it between barriers).
// ------------------------------------------------------------------------------------------------------------------
template <int ITER, bool TAPE>
__global__ void __launch_bounds__(256) hb_persist_small_kernel(const hb_k1_params p0, const hb_k3_params q0, const hb_fused_tdist td,
                                                               const hb_persist_params pp) {
  extern __shared__ __align__(16) double fused_smem[];               // 1 / sqrt(k) [SLD] | statistics rows [8][3][SLD] | row stages
  __shared__ unsigned long long s_cnt[2 * HB_MAX_NCR + 1];
  __shared__ __align__(8) uint64_t sbars[8];
  __shared__ int s_ids[8];
  constexpr int SLD = ITER * 32;
  const int lane = threadIdx.x & 31;
  constexpr int n_wib = 8;                                           // warps per CTA, one chain each
  const int wib = threadIdx.x >> 5;
  const int stage_rows = (1 + 2 * p0.delta) > 4 ? (1 + 2 * p0.delta) : 4;
  double* rs_s = fused_smem;
  double* xrow_s = fused_smem + SLD * (1 + wib);                     // this warp's scratch row
  double* stats = fused_smem + SLD * (1 + n_wib);                    // warp w: rows [w][0..2]: shifted state, its square, dx^2
  double* stage = stats + (size_t)n_wib * 3 * SLD + (size_t)wib * stage_rows * p0.ldx;
  unsigned sphase = 0;
  if (lane == 0) { mbar_init(&sbars[wib], 1); asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
  for (int k = threadIdx.x; k < SLD; k += blockDim.x) rs_s[k] = (k < p0.d) ? td.rs[k] : 0.0;
  if (threadIdx.x <= 2 * HB_MAX_NCR) s_cnt[threadIdx.x] = 0ull;
  __syncthreads();
  const long long warp_id = (long long)blockIdx.x * n_wib + wib;     // this warp's chain
  const long long nl = p0.n_local;
  const bool active = warp_id < nl;                    // one chain per warp (the launcher sizes the grid that way)
  unsigned errbits = 0;
  long long ind_out = pp.ind_out0, pending = pp.pending0;
  k1_headv hv;
  if (active) k1_chain_head<32, TAPE>(p0, (uint32_t)pp.i0, pp.i0 - 2, warp_id, lane, hv, errbits);
  // log of the Metropolis uniform of generation i (R/internals.R:241): like the head, it does not depend on the population
  const long long acc_gch = q0.gchain0 + q0.chain0 + warp_id * q0.stride;
  auto draw_logu = [&](uint32_t gen) -> double {
    const hb_phx ph(q0.seed, (uint64_t)acc_gch, gen);
    return log(hb_u53(hb_pair64(ph.slot(HB_SLOT_ACCEPT), 0)));
  };
  double logu = 0.0;
  if (!TAPE && active && lane == 0) logu = draw_logu((uint32_t)pp.i0);
  // ... and so is the crossover subspace of the proposal (k1_chain_pre): d*, gamma and the words behind lambda / zeta
  k1_prev<ITER> pv;
  if (active) k1_chain_pre<32, ITER, TAPE>(p0, (uint32_t)pp.i0, hv, lane, pv);
  for (long long i = pp.i0; i < pp.i1; ++i) {
    const bool store = ((double)i > pp.burnin * (double)pp.t) && (i % pp.thin == 0);   // R/dream_parallel.R:263, :386
    if (store) ind_out++;
    fused_gen fg;                             // this generation's pointers (uniform; every thread derives them itself)
    {
      const bool flip = ((i - pp.i0) & 1) != 0;
      fg.Xin = flip ? pp.XB : pp.XA; fg.Xout = flip ? pp.XA : pp.XB; fg.gen = (uint32_t)i;
      fg.hist_x = (store && pp.hist_x) ? pp.hist_x + (size_t)(ind_out - 1) * nl * p0.d : nullptr;
      fg.hist_l = (store && pp.hist_l) ? pp.hist_l + (size_t)(ind_out - 1) * nl * 3 : nullptr;
      fg.hist_fx = (store && pp.hist_fx) ? pp.hist_fx + (size_t)(ind_out - 1) * nl : nullptr;
      fg.lpost_hist_row = (i <= pp.H && pp.lpost_hist) ? pp.lpost_hist + (size_t)(i - 1) * nl : nullptr;
      fg.accept_rec = pp.accept_rec ? pp.accept_rec + (size_t)(i - 2) * nl : nullptr;
      fg.u_accept = pp.u_accept ? pp.u_accept + (size_t)(i - 2) * pp.tape_nct : nullptr;
    }
    long long tm[8];
    const bool tr = pp.trace && threadIdx.x == 0 && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1) && i >= pp.i0 + 2 && i < pp.i0 + 6;
    if (tr) tm[4] = clock64();
    fused_chain_step<ITER, TAPE, true>(p0, q0, td, fg, rs_s, xrow_s, hv, lane, s_cnt, errbits, stage, &sbars[wib], sphase, active,
                                       tr ? tm : nullptr, pp.adapt ? q0.shift : nullptr, pp.adapt ? stats + (size_t)wib * 3 * SLD : nullptr,
                                       &s_ids[wib], TAPE ? nullptr : &logu, &pv);
    if (tr) tm[5] = clock64();
    __syncthreads();
    if (tr) tm[6] = clock64();
    const bool upd = (i % pp.update_interval == 0);
    pending += p0.nc;
    const int E = (p0.nCR + 2) * p0.d;
    if (pp.adapt) {
      // this CTA's partial sums, folded over its warps in warp order: [s1 | s2 | Q[0] .. Q[nCR-1]] (hb_k3_kernel's layout)
      for (int e = threadIdx.x; e < E; e += blockDim.x) {
        double sum = 0.0;
        if (e < 2 * p0.d) {
          const double* col = stats + (e < p0.d ? e : SLD + e - p0.d);
#pragma unroll
          for (int w = 0; w < 8; ++w) sum += col[(size_t)w * 3 * SLD];
        } else {
          const int qq = (e - 2 * p0.d) / p0.d, k = (e - 2 * p0.d) - qq * p0.d;
#pragma unroll
          for (int w = 0; w < 8; ++w) sum += (s_ids[w] == qq) ? stats[(size_t)w * 3 * SLD + 2 * SLD + k] : 0.0;
        }
        pp.partials[((size_t)(pp.slot0 + (int)(i - pp.i0)) * gridDim.x + blockIdx.x) * E + e] = sum;
      }
    }
    // the per-generation counters stay in shared memory and reach the control block only when somebody reads them: at
    // updateInterval boundaries (AR / CR rows) and at the end of the launch
    if ((upd || i == pp.i1 - 1) && threadIdx.x <= 2 * HB_MAX_NCR) {
      const unsigned long long v = s_cnt[threadIdx.x];
      if (v) {
        if (threadIdx.x == 0) atomicAdd(&q0.ctrl->n_acc_gen, v);
        else if (threadIdx.x <= HB_MAX_NCR) atomicAdd(&q0.ctrl->n_id_gen[threadIdx.x - 1], v);
        else atomicAdd(&q0.ctrl->n_acc_id_gen[threadIdx.x - 1 - HB_MAX_NCR], v);
        s_cnt[threadIdx.x] = 0ull;
      }
    }
    if (upd || i == pp.i1 - 1) __syncthreads();
    // ---- grid-wide barrier: arrive, compute the NEXT generation's draws (they do not depend on the population) while the
    // other CTAs arrive, then wait.  At updateInterval boundaries CTA 0 folds the AR / CR bookkeeping before it releases
    // the others (they must not add the next generation's counts before the fold has zeroed the counters).
    const unsigned round = (unsigned)(i - pp.i0 + 1);
    if (threadIdx.x == 0) { __threadfence(); atomicAdd(&pp.bar[0], 1u); }
    if (active && i + 1 < pp.i1) {
      k1_chain_head<32, TAPE>(p0, (uint32_t)(i + 1), i - 1, warp_id, lane, hv, errbits);
      k1_chain_pre<32, ITER, TAPE>(p0, (uint32_t)(i + 1), hv, lane, pv);
      if (!TAPE && lane == 0) logu = draw_logu((uint32_t)(i + 1));
    }
    if (threadIdx.x == 0) {
      const unsigned target = gridDim.x * round;
      // adaptation period: nothing to fold here -- the statistics of the interval's generations are parked in slots and the
      // finalize kernels consume them after the launch (which ends at the updateInterval boundary)
      if (!upd || pp.adapt) {
        while (*reinterpret_cast<volatile unsigned*>(&pp.bar[0]) < target) {}
      } else if (blockIdx.x == 0) {
        while (*reinterpret_cast<volatile unsigned*>(&pp.bar[0]) < target) {}
        __threadfence();
        hb_fin_params f{};
        f.ctrl = q0.ctrl; f.nCR = p0.nCR; f.do_arcr = 1; f.out_ar = pp.out_ar; f.out_cr = pp.out_cr; f.ar_rows = pp.ar_rows;
        f.n_eval_inc = pending;
        hb_fin_counters(f);
        __threadfence();
        *reinterpret_cast<volatile unsigned*>(&pp.bar[1]) = round;
      } else {
        while (*reinterpret_cast<volatile unsigned*>(&pp.bar[1]) < round) {}
      }
    }
    if (threadIdx.x == 0) __threadfence();
    if (upd) pending = 0;
    __syncthreads();
    if (tr) printf("[persist trace] gen %lld cta %d: setup %lld | sweep %lld | xp %lld | density sums %lld + log1p %lld | accept+store %lld | block sync %lld | barrier+next head %lld cycles\n",
                   i, (int)blockIdx.x, tm[0] - tm[4], tm[1] - tm[0], tm[2] - tm[1], tm[7] - tm[2], tm[3] - tm[7], tm[5] - tm[3], tm[6] - tm[5], clock64() - tm[6]);
  }
  if (errbits) atomicOr(&q0.ctrl->err, errbits);
}

template <int ITER>
cudaError_t launch_persist(const hb_k1_params& p, const hb_k3_params& q, const hb_fused_tdist& td, const hb_persist_params& pp,
                           bool tape, int sm_count, cudaStream_t st) {
  const int threads = 256, warps = threads / 32;
  const long long blocks = (p.n_local + warps - 1) / warps;
  const int stage_rows = (1 + 2 * p.delta) > 4 ? (1 + 2 * p.delta) : 4;
  if (!td.rs || p.d > ITER * 32 || (p.ldx & 1)) return cudaErrorInvalidValue;
  const size_t smem = ((size_t)ITER * 32 * (1 + 4 * warps) + (size_t)warps * stage_rows * p.ldx) * sizeof(double);
  if (smem > 220 * 1024) return cudaErrorInvalidValue;
  const void* fn = tape ? (const void*)hb_persist_small_kernel<ITER, true> : (const void*)hb_persist_small_kernel<ITER, false>;
  cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, threads, smem);
  if (e != cudaSuccess) return e;
  if (blocks > (long long)per_sm * sm_count) return cudaErrorCooperativeLaunchTooLarge;   // the grid barrier needs every CTA resident
  void* args[] = {(void*)&p, (void*)&q, (void*)&td, (void*)&pp};
  return cudaLaunchCooperativeKernel(fn, dim3((unsigned)blocks), dim3(threads), args, smem, st);
}

// ------------------------------------------------------------------------------------------------------------------
// Wide variants (d > 16): a warp owns a batch of 32 chains.
//   phase A  lane-per-chain: the per-chain draws (snooker test, D, partner ids by partial Fisher-Yates, crossover id,
//            gamma choice) are computed once per chain by ONE lane -- 32 chains in parallel, no redundant lanes --
//            and parked in shared memory;
//   phase B  warp-per-chain: the rows the chain needs (its own state and the 2D partner rows) are fetched by 1-D bulk
//            TMA copies (cp.async.bulk -> shared memory, one mbarrier per stage) TWO chains ahead, so the HBM gather
//            latency is hidden behind the Philox work of the chains in between and costs no registers; the 32 lanes
//            then sweep the row out of shared memory.
// Two kernels share this frame:
//   hb_k1_pair_kernel  native Philox draws.  A lane owns PAIRS of adjacent dimensions (k0 = 2 (32 m + lane), k0 + 1):
//                      one Philox4x32-10 call serves both (slot map v2: 64 bits per dimension), rows are read with
//                      128-bit shared loads and xp leaves as 128-bit stores.  The kernel is instruction-issue bound
//                      (profiles/r01_final_*), so everything is sized in warp instructions per chain.
//   hb_k1_tape_kernel  replay mode: draws come from the decision tape, lane owns dimensions k = 32 m + lane.
// ------------------------------------------------------------------------------------------------------------------
constexpr int K1_STAGES = 2;
constexpr int K1_HEAD_WORDS = 12;   // full layout: packed, a[4], b[4], third snooker row, g_snooker (2 words)
// Without the archive (no snooker rows, no g_snooker) a head is packed | a[delta] | b[delta], rounded up to an even word
// count: 8 words for delta = 3 -- the 4 KB per CTA this saves is what lets a ninth warp fit next to the stage buffers.
__host__ __device__ inline int k1_head_words(int delta, int past_sample) {
  return past_sample ? K1_HEAD_WORDS : ((1 + 2 * delta + 1) & ~1);
}

__host__ __device__ inline int k1_rows_max(int delta) { return (1 + 2 * delta) > 4 ? (1 + 2 * delta) : 4; }
__host__ __device__ inline size_t k1_warp_smem_bytes(int delta, int ldx, int past_sample) {
  // [stage][row][ldx] doubles + 32 chain heads + mbarriers, 16-byte aligned pieces
  return (size_t)K1_STAGES * k1_rows_max(delta) * ldx * 8 + 32 * k1_head_words(delta, past_sample) * 4 + K1_STAGES * 8;
}

// phase A: the head of chain base + lane -> heads[lane][...].  Row ids are absolute rows of the source matrix
// (X: run offset added; Z: archive row), so that phase B turns them into addresses with one multiply.
template <bool TAPE>
__device__ __forceinline__ void k1_phase_a(const hb_k1_params& p, long long base, int lane, int n_src, int* heads,
                                           unsigned& errbits) {
  const int nCR = p.nCR, delta = p.delta;
  int hD = 1, hid = 0;
  int hpa[HB_MAX_DELTA], hpb[HB_MAX_DELTA], hps2 = 0;
  double hgsn = 0.0;
#pragma unroll
  for (int q = 0; q < HB_MAX_DELTA; ++q) { hpa[q] = 0; hpb[q] = 0; }
  const long long jl_raw = base + lane;
  const long long jl = jl_raw < p.n_local ? jl_raw : p.n_local - 1;
  const long long gl = p.chain0 + jl * p.stride;
  const long long run = p.n_runs > 1 ? gl / p.nc : 0;
  const long long j = gl - run * p.nc;
  const long long gch = p.gchain0 + gl;
  double u_sn;
  int g1;
  if (TAPE) {
    const long long rix = p.tape_g * p.tape.nct + gl;
    u_sn = p.tape.u_snooker ? p.tape.u_snooker[rix] : 1.0;          // R/internals.R:119
    hD = p.tape.D[rix];                                             // :123
    hid = p.tape.id[rix];                                           // :173
    g1 = p.tape.g_is_one[rix];                                      // :192
    if (p.tape.g_snooker) hgsn = p.tape.g_snooker[rix];             // :156
    if (hD < 1 || hD > delta || hid >= nCR) { errbits |= HB_FLAG_BAD_TAPE; hD = 1; hid = 0; }
    const int32_t* pr = p.tape.partners + rix * (2 * delta);
#pragma unroll
    for (int q = 0; q < HB_MAX_DELTA; ++q)
      if (q < hD) { hpa[q] = pr[q]; hpb[q] = pr[hD + q]; }
    if (p.past_sample && u_sn < p.psnooker) { hpa[0] = pr[0]; hpb[0] = pr[1]; hps2 = pr[2]; }
  } else {
    const hb_phx ph(p.seed, (uint64_t)gch, p.gen);
    const uint4 w0 = ph.slot(HB_SLOT_HEAD);
    const uint4 wc = ph.slot(HB_SLOT_CHOICE);
    u_sn = hb_u53(hb_pair64(w0, 0));
    hD = 1 + (int)hb_mulhi64(hb_pair64(w0, 1), (uint64_t)delta);
    const double u_id = hb_u53(hb_pair64(wc, 0));
    double cum = 0.0;
    hid = nCR - 1;
    bool found = false;
    for (int q = 0; q < nCR; ++q) {
      cum += p.ctrl[run].pCR[q];
      if (!found && u_id < cum) { hid = q; found = true; }
    }
    g1 = hb_u53(hb_pair64(wc, 1)) >= 1.0 - p.p_g;
    const bool sn = p.past_sample && (u_sn < p.psnooker);
    if (sn) hgsn = 1.2 + hb_u53(hb_pair64(ph.slot(HB_SLOT_ACCEPT), 1));
    const int n_part = sn ? 3 : 2 * hD;
    hb_fy fy;
    long long n_rem = p.past_sample ? p.m_arch : p.nc - 1;
    int y[2 * HB_MAX_DELTA];
    uint4 wp = make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int q = 0; q < 2 * HB_MAX_DELTA; ++q) {
      y[q] = 0;
      if (q < n_part) {
        if ((q & 1) == 0) wp = ph.slot(HB_SLOT_PARTNER0 + (q >> 1));
        long long v = fy.draw(hb_pair64(wp, q & 1), n_rem);
        if (!p.past_sample && v >= j) v += 1;                       // sample((1:nc)[-j], ...)  :142
        y[q] = (int)v;
      }
    }
#pragma unroll
    for (int q = 0; q < HB_MAX_DELTA; ++q) {
      hpa[q] = y[q];
      hpb[q] = (hD == 1) ? y[(1 + q) & 7] : (hD == 2) ? y[(2 + q) & 7] : (hD == 3) ? y[(3 + q) & 7] : y[(4 + q) & 7];
    }
    if (sn) { hpb[0] = y[1]; hps2 = y[2]; }
  }
  const bool sn = p.past_sample && (u_sn < p.psnooker);
#pragma unroll
  for (int q = 0; q < HB_MAX_DELTA; ++q) {
    const bool used = sn ? (q == 0) : (q < hD);
    if (used && (hpa[q] < 0 || hpa[q] >= n_src || hpb[q] < 0 || hpb[q] >= n_src)) {
      errbits |= HB_FLAG_BAD_TAPE; hpa[q] = 0; hpb[q] = 0;
    }
  }
  if (sn && (hps2 < 0 || hps2 >= n_src)) { errbits |= HB_FLAG_BAD_TAPE; hps2 = 0; }
  const int roff = p.past_sample ? 0 : (int)(gl - j);               // first chain of this chain's run
  // head = packed | the rows to fetch in issue order: a[0..D-1], b[0..D-1] (snooker: r1, r2, r3) | g_snooker at words
  // 10-11 of the full layout.  Lane r of the issue step reads word r.
  const int hw = k1_head_words(delta, p.past_sample);
  int* hd = heads + lane * hw;
  hd[0] = hD | (hid << 8) | ((g1 ? 1 : 0) << 16) | ((sn ? 1 : 0) << 17);
  if (sn) { hd[1] = hpa[0] + roff; hd[2] = hpb[0] + roff; hd[3] = hps2 + roff; }
  else {
#pragma unroll
    for (int q = 0; q < HB_MAX_DELTA; ++q) if (q < hD) { hd[1 + q] = hpa[q] + roff; hd[1 + hD + q] = hpb[q] + roff; }
  }
  if (p.past_sample) *reinterpret_cast<double*>(hd + 10) = hgsn;
}

// issue the bulk copies of the rows chain `c` of the current batch needs into stage buffer `dst`:
// stage rows 0 = x, 1..D = a, 1+delta.. = b (fixed offsets: no D-dependent address arithmetic in the sweep)
__device__ __forceinline__ void k1_issue(const hb_k1_params& p, const double* __restrict__ src, const int* heads, int c,
                                         long long base, double* dst, uint64_t* bar, int lane, uint32_t row_bytes) {
  const int* hd = heads + c * k1_head_words(p.delta, p.past_sample);
  const int pk = hd[0];
  const int Dn = pk & 0xff;
  const bool sn = (pk >> 17) & 1;
  const int n_rows = sn ? 4 : 1 + 2 * Dn;
  if (lane == 0) mbar_expect_tx(bar, (uint32_t)n_rows * row_bytes);
  __syncwarp();
  if (lane < n_rows) {
    const double* g;
    if (lane == 0) g = p.X + (p.chain0 + (base + c) * p.stride) * (long long)p.ldx;      // own state
    else {
      const int row = hd[lane];                                     // a[0..D-1], b[0..D-1] (snooker: r1, r2, r3)
      if (p.n_peers > 1) {                        // the row lives in rank rk's shard: NVLink peer pointer
        const int rk = (int)(row / p.peer_nl);
        g = p.peer[rk] + (long long)(row - rk * p.peer_nl) * p.ldx;
      } else g = src + (long long)row * p.ldx;
    }
    const int srow = (sn || lane <= Dn) ? lane : (1 + p.delta + (lane - 1 - Dn));
    tma_load_1d(dst + (size_t)srow * p.ldx, g, row_bytes, bar);
  }
}

__device__ __forceinline__ double2 lds2(const double* q) { return *reinterpret_cast<const double2*>(q); }

constexpr int K1_PAIR_WARPS = 8;     // 2 CTAs x 8 warps per SM at 126 registers.  Measured: 9 warps (launch bounds 288 -> 96
                                     // registers, spills) is 9 % SLOWER (1.174 vs 1.075 ms); 14 warps/SM is 13 % slower
// TD: the handle's target is the built-in t_distribution in its structured evaluation -- the kernel also accumulates
// sum u, sum u^2 (u_k = xp_k / sqrt(k)) of every proposal while it holds it in registers and writes the log-density, so the
// plugin launch (a second pass over the proposals: 808 bytes per chain) is not needed.
template <int ITER2, bool TD>
__global__ void __launch_bounds__(K1_PAIR_WARPS * 32, 2) hb_k1_pair_kernel(const hb_k1_params p) {
  constexpr unsigned FULL = 0xffffffffu;
  constexpr bool KEEP = ITER2 <= 4;     // keep the Philox words of the subspace pass in registers (else recompute)
  extern __shared__ __align__(16) unsigned char k1_smem[];
  __shared__ __align__(16) double td_rs_s[TD ? 2 * 32 * ITER2 : 2];   // 1 / sqrt(k + 1), zero past d
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const long long warp_id = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int d = p.d, ldx = p.ldx, delta = p.delta;
  const double* __restrict__ src = p.past_sample ? p.Z : p.X;
  const int n_src = (int)(p.past_sample ? p.m_arch : p.nc);
  const int rows_max = k1_rows_max(delta);
  const uint32_t row_bytes = (uint32_t)ldx * 8u;
  const int hw = k1_head_words(delta, p.past_sample);
  unsigned char* wbase = k1_smem + (size_t)wib * k1_warp_smem_bytes(delta, ldx, p.past_sample);
  double* stage_buf = reinterpret_cast<double*>(wbase);                                  // [STAGES][rows_max][ldx]
  int* heads = reinterpret_cast<int*>(wbase + (size_t)K1_STAGES * rows_max * ldx * 8);   // [32][HEAD_WORDS]
  uint64_t* bars = reinterpret_cast<uint64_t*>(heads + 32 * hw);                         // [STAGES]
  unsigned errbits = 0;
  const bool bound_or_prior = (p.bound != 0) || (p.prior == 1);

  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < K1_STAGES; ++s) mbar_init(&bars[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (TD) {
    for (int k = threadIdx.x; k < 2 * 32 * ITER2; k += blockDim.x) td_rs_s[k] = (k < d) ? p.td_rs[k] : 0.0;
    __syncthreads();
  }
  __syncwarp();
  unsigned it = 0;   // chains consumed by this warp so far: stage = it % STAGES, parity = (it / STAGES) & 1

  const int B = p.batch;   // chains per batch: 32, or fewer when the launch has fewer 32-chain batches than resident warps
  for (long long base = warp_id * B; base < p.n_local; base += n_warps * B) {
    k1_phase_a<false>(p, base, lane, n_src, heads, errbits);
    __syncwarp();
    double lp_mine = 0.0;
    double td_s1_mine = 0.0, td_s2_mine = 0.0;     // TD: lane c keeps chain c's sums
    double td_p1 = 0.0, td_p2 = 0.0;               // TD: the previous chain's per-lane partial sums -- their warp reduction (a chain
                                                   // of dependent shuffles and adds) is issued inside the NEXT chain's Philox pass

    const int n_in_batch = (int)((p.n_local - base) < B ? (p.n_local - base) : B);
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
#pragma unroll
    for (int s = 0; s < K1_STAGES; ++s)
      if (s < n_in_batch) {
        const int sg = (int)((it + s) % K1_STAGES);
        k1_issue(p, src, heads, s, base, stage_buf + (size_t)sg * rows_max * ldx, &bars[sg], lane, row_bytes);
      }

    for (int c = 0; c < n_in_batch; ++c, ++it) {
      const int stg = (int)(it % K1_STAGES);
      const int* hd = heads + c * hw;
      const int pk = hd[0];
      const int D = pk & 0xff, id = (pk >> 8) & 0xff;
      const bool g_is_one = (pk >> 16) & 1, snooker = (pk >> 17) & 1;
      const long long jl = base + c;
      const long long gch = p.gchain0 + p.chain0 + jl * p.stride;
      const uint32_t c0 = (uint32_t)gch, c1 = (uint32_t)((uint64_t)gch >> 32);
      double* rows = stage_buf + (size_t)stg * rows_max * ldx;          // row 0: x; 1..: a; 1+delta..: b
      const double* __restrict__ rowa = rows + ldx;
      const double* __restrict__ rowb = rows + (1 + delta) * ldx;
      double2 xp2[ITER2];
      if (TD && c > 0) {                       // chain c - 1's density sums: independent of everything this chain does
        hb_warp_sum2(td_p1, td_p2, lane);
        if (lane == c - 1) { td_s1_mine = td_p1; td_s2_mine = td_p2; }
      }

      if (snooker) {
        // snooker update, R/internals.R:150-163 with orth_proj :250-263 (rare path: psnooker is 0.1 at most)
        mbar_wait(&bars[stg], (it / K1_STAGES) & 1);
        const double g_sn = *reinterpret_cast<const double*>(hd + 10);
        const double* r1 = rows + 1 * ldx;
        const double* r2 = rows + 2 * ldx;
        const double* r3 = rows + 3 * ldx;
        double xk[ITER2][2], uu[ITER2][2], a2[ITER2][2], a3[ITER2][2];
        double s2 = 0, s3 = 0, su = 0, l2 = 0, l3 = 0, lu = 0;
        int differs = 0;
#pragma unroll
        for (int m = 0; m < ITER2; ++m) {
          const int k0 = 2 * (m * 32 + lane);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int k = k0 + h;
            xk[m][h] = 0; uu[m][h] = 0; a2[m][h] = 0; a3[m][h] = 0;
            if (k < d) {
              xk[m][h] = rows[k];
              const double b = r1[k];
              a2[m][h] = r2[k]; a3[m][h] = r3[k];
              uu[m][h] = __dadd_rn(xk[m][h], -b);
              differs |= (xk[m][h] != b);
              comp_add(s2, l2, __dmul_rn(__dadd_rn(a2[m][h], -xk[m][h]), uu[m][h]));
              comp_add(s3, l3, __dmul_rn(__dadd_rn(a3[m][h], -xk[m][h]), uu[m][h]));
              comp_add(su, lu, __dmul_rn(uu[m][h], uu[m][h]));
            }
          }
        }
        s2 = comp_group_sum<32>(s2, l2, FULL); s3 = comp_group_sum<32>(s3, l3, FULL); su = comp_group_sum<32>(su, lu, FULL);
        differs = hb_group_sum_int<32>(differs);
#pragma unroll
        for (int m = 0; m < ITER2; ++m) {
          const int k0 = 2 * (m * 32 + lane);
          uint4 ww = make_uint4(0, 0, 0, 0);
          if (k0 < d) ww = hb_philox4x32_10_ks(c0, c1, p.gen, HB_SLOT_DIM0 + (uint32_t)(k0 >> 1), p.ks);
          double o[2] = {0.0, 0.0};
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int k = k0 + h;
            if (k < d) {
              const double zeta = __dmul_rn(p.c_star, hb_rng_normal32(h ? ww.w : ww.y));
              double rp2 = a2[m][h], rp3 = a3[m][h];
              if (differs) {
                rp2 = __dadd_rn(xk[m][h], __ddiv_rn(__dmul_rn(uu[m][h], s2), su));
                rp3 = __dadd_rn(xk[m][h], __ddiv_rn(__dmul_rn(uu[m][h], s3), su));
              }
              if (!isfinite(rp2) || !isfinite(rp3)) errbits |= HB_FLAG_PROP_NONFINITE;
              o[h] = __dadd_rn(xk[m][h], __dadd_rn(zeta, __dmul_rn(g_sn, __dadd_rn(rp2, -rp3))));
            }
          }
          xp2[m] = make_double2(o[0], o[1]);
        }
      } else {
        // parallel direction update, R/internals.R:166-200
        const uint32_t thr = p.thr16[id];     // zt_k < CR[id]  <=>  (lo & 0xffff) < thr  (hb_rng_zt_threshold16)
        uint4 wk[KEEP ? ITER2 : 1];
        unsigned s0[ITER2], s1[ITER2];
        int d_star = 0;
#pragma unroll
        for (int m = 0; m < ITER2; ++m) {
          const int k0 = 2 * (m * 32 + lane);
          bool a = false, b = false;
          uint4 ww = make_uint4(0, 0, 0, 0);
          if (k0 < d) {
            ww = hb_philox4x32_10_ks(c0, c1, p.gen, HB_SLOT_DIM0 + (uint32_t)(k0 >> 1), p.ks);
            a = (ww.x & 0xffffu) < thr;                                   // :175-177
            b = (k0 + 1 < d) && ((ww.z & 0xffffu) < thr);
          }
          if (KEEP) wk[KEEP ? m : 0] = ww;
          s0[m] = __ballot_sync(FULL, a);
          s1[m] = __ballot_sync(FULL, b);
          d_star += __popc(s0[m]) + __popc(s1[m]);
        }
        int kmin = -1;
        if (d_star == 0) {                                               // :181-184  A <- which.min(zt), first minimum
          unsigned best = 0xffffffffu;
#pragma unroll
          for (int m = 0; m < ITER2; ++m) {
            const int k0 = 2 * (m * 32 + lane);
            if (k0 < d) {
              const uint4 ww = KEEP ? wk[KEEP ? m : 0] : hb_philox4x32_10_ks(c0, c1, p.gen, HB_SLOT_DIM0 + (uint32_t)(k0 >> 1), p.ks);
              best = min(best, ((ww.x & 0xffffu) << 16) | (unsigned)k0);
              if (k0 + 1 < d) best = min(best, ((ww.z & 0xffffu) << 16) | (unsigned)(k0 + 1));
            }
          }
          best = __reduce_min_sync(FULL, best);
          kmin = (int)(best & 0xffffu); d_star = 1;
        }
        // gamma_d = 2.38/sqrt(2 D d*) (:190) from the host-built table (same IEEE sqrt and division, bit-identical)
        const double gam = g_is_one ? 1.0 : p.gamma_tab[(D - 1) * d + (d_star - 1)];   // :192
        mbar_wait(&bars[stg], (it / K1_STAGES) & 1);                     // the rows have landed in shared memory
        // Branch-free sweep: every lane computes its pair unconditionally (out-of-range lanes on a clamped address) and
        // the crossover subspace enters as a select at the end -- unselected lanes would idle through the same warp
        // instructions anyway, and without the divergent branches the iterations form one basic block whose
        // independent fp64 chains the scheduler can interleave.  Only the warp-uniform branches on D remain.
#pragma unroll
        for (int m = 0; m < ITER2; ++m) {
          const int k0 = 2 * (m * 32 + lane);
          const int kc = (k0 < d) ? k0 : 0;
          const uint4 ww = KEEP ? wk[KEEP ? m : 0] : hb_philox4x32_10_ks(c0, c1, p.gen, HB_SLOT_DIM0 + (uint32_t)(kc >> 1), p.ks);
          const bool a = (kmin >= 0) ? (k0 == kmin) : ((s0[m] >> lane) & 1u);      // masks are 0 for out-of-range lanes
          const bool b = (kmin >= 0) ? (k0 + 1 == kmin) : ((s1[m] >> lane) & 1u);
          double2 x = lds2(rows + kc);
          // jump_diff <- colSums(r1[,A] - r2[,A])  :196.  One or two terms: the plain sum is already the correctly
          // rounded one; from the third term on, carry the rounding errors (R accumulates in long double).
          double2 tq[HB_MAX_DELTA];
#pragma unroll
          for (int q = 0; q < HB_MAX_DELTA; ++q) {
            tq[q] = make_double2(0.0, 0.0);
            if (q < D) {
              const double2 ra = lds2(rowa + q * ldx + kc), rb = lds2(rowb + q * ldx + kc);
              tq[q] = make_double2(__dadd_rn(ra.x, -rb.x), __dadd_rn(ra.y, -rb.y));
            }
          }
          double jd0, jd1;                        // (making these D-branches unconditional was measured 11 % slower)
          if (D <= 2) { jd0 = __dadd_rn(tq[0].x, tq[1].x); jd1 = __dadd_rn(tq[0].y, tq[1].y); }
          else {
            double sa, ea, ca, sb, eb, cb;
            two_sum(tq[0].x, tq[1].x, sa, ca);
            two_sum(tq[0].y, tq[1].y, sb, cb);
#pragma unroll
            for (int q = 2; q < HB_MAX_DELTA; ++q)
              if (q < D) {
                two_sum(sa, tq[q].x, sa, ea); ca = __dadd_rn(ca, ea);
                two_sum(sb, tq[q].y, sb, eb); cb = __dadd_rn(cb, eb);
              }
            jd0 = __dadd_rn(sa, ca); jd1 = __dadd_rn(sb, cb);
          }
          {
            const double lambda = hb_rng_lambda16(ww.x >> 16, p.c_val);                              // :188
            const double zeta = __dmul_rn(p.c_star, hb_rng_normal32(ww.y));                          // :194
            const double v = __dadd_rn(zeta, __dmul_rn(__dmul_rn(__dadd_rn(1.0, lambda), gam), jd0));   // :198
            const double nx = __dadd_rn(x.x, __dmul_rn(v, p.beta0));                                 // :200, :206
            x.x = a ? nx : x.x;
          }
          {
            const double lambda = hb_rng_lambda16(ww.z >> 16, p.c_val);
            const double zeta = __dmul_rn(p.c_star, hb_rng_normal32(ww.w));
            const double v = __dadd_rn(zeta, __dmul_rn(__dmul_rn(__dadd_rn(1.0, lambda), gam), jd1));
            const double ny = __dadd_rn(x.y, __dmul_rn(v, p.beta0));
            x.y = b ? ny : x.y;
          }
          xp2[m] = x;
        }
      }
      // the stage is free again: fetch the rows of the chain K1_STAGES ahead
      __syncwarp();
      if (c + K1_STAGES < n_in_batch) {
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        k1_issue(p, src, heads, c + K1_STAGES, base, rows, &bars[stg], lane, row_bytes);
      }

      // finite check ; bounds ; prior ; store
      double lp_part = 0.0;
      double td_s1 = 0.0, td_s2 = 0.0;
      double* __restrict__ xp_row = p.XP + jl * (long long)ldx;
#pragma unroll
      for (int m = 0; m < ITER2; ++m) {
        const int k0 = 2 * (m * 32 + lane);
        if (k0 < d) {
          double2 x = xp2[m];
          const bool has1 = k0 + 1 < d;
          if (!(fabs(x.x) <= 1.7976931348623157e308) || !(fabs(x.y) <= 1.7976931348623157e308)) errbits |= HB_FLAG_PROP_NONFINITE;   // :209
          if (bound_or_prior) {
            if (p.bound) {                                                       // :210-214
              x.x = hb_bound_par(x.x, p.pmin[k0], p.pmax[k0], p.bound);
              if (has1) x.y = hb_bound_par(x.y, p.pmin[k0 + 1], p.pmax[k0 + 1], p.bound);
            }
            if (p.prior == 1) {                                                  // dunif(log = TRUE), :57
              { const double a = p.pmin[k0], b = p.pmax[k0]; lp_part += (a <= x.x && x.x <= b) ? -log(b - a) : -INFINITY; }
              if (has1) { const double a = p.pmin[k0 + 1], b = p.pmax[k0 + 1]; lp_part += (a <= x.y && x.y <= b) ? -log(b - a) : -INFINITY; }
            }
          }
          *reinterpret_cast<double2*>(xp_row + k0) = x;
          if (TD) {                                      // u = xp / sqrt(k) (the scale of a coordinate past d is zero)
            const double2 r = lds2(td_rs_s + k0);
            hb_tdist_acc(td_s1, td_s2, x.x, has1 ? x.y : 0.0, r.x, r.y);
          }
        }
      }
      if (p.prior == 1) { lp_part = hb_group_sum<32>(lp_part); if (lane == c) lp_mine = lp_part; }   // lane c keeps chain c's log prior
      if (TD) { td_p1 = td_s1; td_p2 = td_s2; }
    }
    if (TD && n_in_batch > 0) {                // the last chain of the batch
      hb_warp_sum2(td_p1, td_p2, lane);
      if (lane == n_in_batch - 1) { td_s1_mine = td_p1; td_s2_mine = td_p2; }
    }
    // the scalars of the batch leave as coalesced stores: lane c owns chain base + c
    if (lane < n_in_batch) {
      double lp = lp_mine;                                               // flat prior: 0  :55
      if (!(lp >= HB_FLOOR)) lp = (lp != lp) ? lp : HB_FLOOR;            // :65
      if (!isfinite(lp)) errbits |= HB_FLAG_LP_NONFINITE;                // :67
      p.lp_xp[base + lane] = lp;
      p.id_out[base + lane] = (heads[lane * hw] >> 8) & 0xff;
      if (TD) {                                                          // dmvt, R/test_t_distribution.R:43 (hb_tdist_struct_kernel)
        const double v = hb_tdist_logdens(td_s1_mine, td_s2_mine, d, p.td_konst, p.td_df);
        if (p.td_fx) p.td_fx[base + lane] = v;
        p.td_ll[base + lane] = (p.td_lik == 1) ? log(v) : v;
      }
    }
    __syncwarp();
  }
  if (errbits) atomicOr(p.err, errbits);
}

// replay mode: the draws come from the decision tape; lane owns dimensions k = 32 m + lane
template <int ITER>
__global__ void __launch_bounds__(256, 2) hb_k1_tape_kernel(const hb_k1_params p) {
  constexpr unsigned FULL = 0xffffffffu;
  extern __shared__ __align__(16) unsigned char k1_smem[];
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const long long warp_id = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int d = p.d, ldx = p.ldx, nCR = p.nCR, delta = p.delta;
  const double* __restrict__ src = p.past_sample ? p.Z : p.X;
  const int n_src = (int)(p.past_sample ? p.m_arch : p.nc);
  const int rows_max = k1_rows_max(delta);
  const uint32_t row_bytes = (uint32_t)ldx * 8u;
  const int hw = k1_head_words(delta, p.past_sample);
  unsigned char* wbase = k1_smem + (size_t)wib * k1_warp_smem_bytes(delta, ldx, p.past_sample);
  double* stage_buf = reinterpret_cast<double*>(wbase);                                  // [STAGES][rows_max][ldx]
  int* heads = reinterpret_cast<int*>(wbase + (size_t)K1_STAGES * rows_max * ldx * 8);   // [32][HEAD_WORDS]
  uint64_t* bars = reinterpret_cast<uint64_t*>(heads + 32 * hw);                         // [STAGES]
  unsigned errbits = 0;
  const bool bound_or_prior = (p.bound != 0) || (p.prior == 1);

  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < K1_STAGES; ++s) mbar_init(&bars[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncwarp();
  unsigned it = 0;

  const int B = p.batch;
  for (long long base = warp_id * B; base < p.n_local; base += n_warps * B) {
    k1_phase_a<true>(p, base, lane, n_src, heads, errbits);
    __syncwarp();

    const int n_in_batch = (int)((p.n_local - base) < B ? (p.n_local - base) : B);
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
#pragma unroll
    for (int s = 0; s < K1_STAGES; ++s)
      if (s < n_in_batch) {
        const int sg = (int)((it + s) % K1_STAGES);
        k1_issue(p, src, heads, s, base, stage_buf + (size_t)sg * rows_max * ldx, &bars[sg], lane, row_bytes);
      }

    for (int c = 0; c < n_in_batch; ++c, ++it) {
      const int stg = (int)(it % K1_STAGES);
      const int* hd = heads + c * hw;
      const int pk = hd[0];
      const int D = pk & 0xff, id = (pk >> 8) & 0xff;
      const bool g_is_one = (pk >> 16) & 1, snooker = (pk >> 17) & 1;
      const long long jl = base + c;
      const long long gl = p.chain0 + jl * p.stride;
      const long long rix = p.tape_g * p.tape.nct + gl;
      double* rows = stage_buf + (size_t)stg * rows_max * ldx;          // row 0: x; 1..: a; 1+delta..: b
      const double* __restrict__ rowa = rows + ldx;
      const double* __restrict__ rowb = rows + (1 + delta) * ldx;
      double xk[ITER], dxk[ITER];

      if (snooker) {
        // snooker update, R/internals.R:150-163 with orth_proj :250-263
        mbar_wait(&bars[stg], (it / K1_STAGES) & 1);
        const double g_sn = *reinterpret_cast<const double*>(hd + 10);
        const double* r1 = rows + 1 * ldx;
        const double* r2 = rows + 2 * ldx;
        const double* r3 = rows + 3 * ldx;
        double uu[ITER], a2[ITER], a3[ITER];
        double s2 = 0, s3 = 0, su = 0, l2 = 0, l3 = 0, lu = 0;
        int differs = 0;
#pragma unroll
        for (int m = 0; m < ITER; ++m) {
          const int k = m * 32 + lane;
          xk[m] = 0; uu[m] = 0; a2[m] = 0; a3[m] = 0;
          if (k < d) {
            xk[m] = rows[k];
            const double b = r1[k];
            a2[m] = r2[k]; a3[m] = r3[k];
            uu[m] = __dadd_rn(xk[m], -b);
            differs |= (xk[m] != b);
            comp_add(s2, l2, __dmul_rn(__dadd_rn(a2[m], -xk[m]), uu[m]));
            comp_add(s3, l3, __dmul_rn(__dadd_rn(a3[m], -xk[m]), uu[m]));
            comp_add(su, lu, __dmul_rn(uu[m], uu[m]));
          }
        }
        s2 = comp_group_sum<32>(s2, l2, FULL); s3 = comp_group_sum<32>(s3, l3, FULL); su = comp_group_sum<32>(su, lu, FULL);
        differs = hb_group_sum_int<32>(differs);
#pragma unroll
        for (int m = 0; m < ITER; ++m) {
          const int k = m * 32 + lane;
          dxk[m] = 0;
          if (k < d) {
            const double zeta = p.tape.zeta[rix * d + k];
            double rp2 = a2[m], rp3 = a3[m];
            if (differs) {
              rp2 = __dadd_rn(xk[m], __ddiv_rn(__dmul_rn(uu[m], s2), su));
              rp3 = __dadd_rn(xk[m], __ddiv_rn(__dmul_rn(uu[m], s3), su));
            }
            if (!isfinite(rp2) || !isfinite(rp3)) errbits |= HB_FLAG_PROP_NONFINITE;
            dxk[m] = __dadd_rn(zeta, __dmul_rn(g_sn, __dadd_rn(rp2, -rp3)));
          }
        }
      } else {
        // parallel direction update, R/internals.R:166-200
        const double CRv = (double)(id + 1) / (double)nCR;               // CR <- (1:nCR)/nCR
        double ztv[ITER];
        unsigned selmask[ITER];
        int d_star = 0;
#pragma unroll
        for (int m = 0; m < ITER; ++m) {
          const int k = m * 32 + lane;
          bool sel = false;
          ztv[m] = 2.0;
          if (k < d) { ztv[m] = p.tape.zt[rix * d + k]; sel = ztv[m] < CRv; }   // :175-177
          selmask[m] = __ballot_sync(FULL, sel);
          d_star += __popc(selmask[m]);
        }
        int kmin = -1;
        if (d_star == 0) {                                               // :181-184  A <- which.min(zt)
          int bk = 0x7fffffff;
          double bv = 3.0;
#pragma unroll
          for (int m = 0; m < ITER; ++m) { const int k = m * 32 + lane; if (k < d && ztv[m] < bv) { bv = ztv[m]; bk = k; } }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_xor_sync(FULL, bv, o); const int ok = __shfl_xor_sync(FULL, bk, o);
            if (ov < bv || (ov == bv && ok < bk)) { bv = ov; bk = ok; }
          }
          kmin = bk; d_star = 1;
        }
        const double gam = g_is_one ? 1.0 : p.gamma_tab[(D - 1) * d + (d_star - 1)];   // :190-192
        mbar_wait(&bars[stg], (it / K1_STAGES) & 1);                     // the rows have landed in shared memory
        int rank_base = 0;
#pragma unroll
        for (int m = 0; m < ITER; ++m) {
          const int k = m * 32 + lane;
          const bool sel = (kmin >= 0) ? (k == kmin) : ((selmask[m] >> lane) & 1u);
          xk[m] = 0; dxk[m] = 0;
          if (k < d) {
            xk[m] = rows[k];
            if (sel) {
              const int rank = (kmin >= 0) ? 0 : rank_base + __popc(selmask[m] & ((1u << lane) - 1u));
              const double lambda = p.tape.lambda[rix * d + rank];       // :188 (by rank within A)
              const double zeta = p.tape.zeta[rix * d + rank];           // :194
              double tq[HB_MAX_DELTA];
#pragma unroll
              for (int q = 0; q < HB_MAX_DELTA; ++q) {
                tq[q] = 0.0;
                if (q < D) tq[q] = __dadd_rn(rowa[q * ldx + k], -rowb[q * ldx + k]);
              }
              double jd;                                                 // colSums(r1[,A] - r2[,A])  :196
              if (D <= 2) jd = __dadd_rn(tq[0], tq[1]);
              else {
                double s, e, comp;
                two_sum(tq[0], tq[1], s, comp);
#pragma unroll
                for (int q = 2; q < HB_MAX_DELTA; ++q) if (q < D) { two_sum(s, tq[q], s, e); comp = __dadd_rn(comp, e); }
                jd = __dadd_rn(s, comp);
              }
              const double v = __dadd_rn(zeta, __dmul_rn(__dmul_rn(__dadd_rn(1.0, lambda), gam), jd));   // :198
              dxk[m] = __dmul_rn(v, p.beta0);                                                           // :200
            }
          }
          rank_base += __popc(selmask[m]);
        }
      }
      __syncwarp();
      if (c + K1_STAGES < n_in_batch) {
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        k1_issue(p, src, heads, c + K1_STAGES, base, rows, &bars[stg], lane, row_bytes);
      }

      double lp_part = 0.0;
      double* __restrict__ xp_row = p.XP + jl * (long long)ldx;
#pragma unroll
      for (int m = 0; m < ITER; ++m) {
        const int k = m * 32 + lane;
        if (k < d) {
          double xp = __dadd_rn(xk[m], dxk[m]);                          // :206
          if (!(fabs(xp) <= 1.7976931348623157e308)) errbits |= HB_FLAG_PROP_NONFINITE;   // :209
          if (bound_or_prior) {
            if (p.bound) xp = hb_bound_par(xp, p.pmin[k], p.pmax[k], p.bound);   // :210-214
            if (p.prior == 1) {                                          // dunif(log = TRUE), :57
              const double a = p.pmin[k], b = p.pmax[k];
              lp_part += (a <= xp && xp <= b) ? -log(b - a) : -INFINITY;
            }
          }
          xp_row[k] = xp;
        }
      }
      if (p.prior == 1) lp_part = hb_group_sum<32>(lp_part);
      if (lane == 0) {
        double lp = (p.prior == 1) ? lp_part : 0.0;                      // flat prior: 0  :55
        if (!(lp >= HB_FLOOR)) lp = (lp != lp) ? lp : HB_FLOOR;          // :65
        if (!isfinite(lp)) errbits |= HB_FLAG_LP_NONFINITE;              // :67
        p.lp_xp[jl] = lp;
        p.id_out[jl] = id;
      }
    }
    __syncwarp();
  }
  if (errbits) atomicOr(p.err, errbits);
}

inline void k1_wide_geometry(hb_k1_params& p, int sm_count, int max_warps, int* threads, long long* blocks, size_t* smem) {
  const size_t per_warp = (k1_warp_smem_bytes(p.delta, p.ldx, p.past_sample) + 15) & ~(size_t)15;
  int warps = max_warps;                                               // two CTAs per SM
  while (warps > 1 && per_warp * warps * 2 + 4096 > 227 * 1024) warps = warps > 8 ? 8 : warps >> 1;
  int ctas = 2;
  if (const char* e = getenv("HB_K1_WARPS")) warps = atoi(e);          // tuning experiments only
  if (const char* e = getenv("HB_K1_CTAS")) ctas = atoi(e);
  *smem = per_warp * warps;
  *threads = warps * 32;
  long long cap = (long long)sm_count * ctas - p.reserve_ctas;          // persistent: two CTAs per SM, grid-stride batches
  if (cap < 1) cap = 1;
  // a warp works through its batch chain by chain, so the launch takes ceil(batches / warps) batch times: size the batch
  // so that the last wave is full (a piece of a multi-GPU shard is only one or two waves, where a fixed 32 would leave up
  // to half of the warps idle for a whole batch time)
  const long long n_warps_total = cap * warps;
  const long long waves = (p.n_local + n_warps_total * 32 - 1) / (n_warps_total * 32);
  long long Bq = (p.n_local + n_warps_total * waves - 1) / (n_warps_total * waves);
  int B = (int)(Bq > 32 ? 32 : Bq < 4 ? 4 : Bq);
  if (const char* e = getenv("HB_K1_BATCH")) B = atoi(e);
  p.batch = B;
  const long long warps_needed = (p.n_local + B - 1) / B;
  long long b = (warps_needed + warps - 1) / warps;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  *blocks = b;
}

template <int ITER>
cudaError_t launch_wide_tape(hb_k1_params p, int sm_count, cudaStream_t st) {
  int threads; long long blocks; size_t smem;
  k1_wide_geometry(p, sm_count, 8, &threads, &blocks, &smem);
  if (smem > 220 * 1024) return cudaErrorInvalidValue;
  cudaError_t e = cudaFuncSetAttribute(hb_k1_tape_kernel<ITER>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  hb_k1_tape_kernel<ITER><<<(unsigned)blocks, threads, smem, st>>>(p);
  return cudaGetLastError();
}

template <int ITER2>
cudaError_t launch_wide_pair(hb_k1_params p, int sm_count, cudaStream_t st, bool* density_done) {
  int threads; long long blocks; size_t smem;
  k1_wide_geometry(p, sm_count, K1_PAIR_WARPS, &threads, &blocks, &smem);
  if (smem > 220 * 1024) return cudaErrorInvalidValue;
  const bool td = ITER2 <= 4 && p.td_rs != nullptr && p.td_ll != nullptr && density_done != nullptr && p.n_runs <= 1;
  cudaError_t e = td ? cudaFuncSetAttribute(hb_k1_pair_kernel<ITER2, (ITER2 <= 4)>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                     : cudaFuncSetAttribute(hb_k1_pair_kernel<ITER2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  if (td) hb_k1_pair_kernel<ITER2, (ITER2 <= 4)><<<(unsigned)blocks, threads, smem, st>>>(p);
  else hb_k1_pair_kernel<ITER2, false><<<(unsigned)blocks, threads, smem, st>>>(p);
  e = cudaGetLastError();
  if (e == cudaSuccess && td) *density_done = true;
  return e;
}

template <int G, int ITER>
cudaError_t launch_gi(const hb_k1_params& p, bool tape, int sm_count, cudaStream_t st) {
  constexpr int CPW = 32 / G;
  const int threads = 256;
  const long long warps_needed = (p.n_local + CPW - 1) / CPW;
  long long blocks = (warps_needed + (threads / 32) - 1) / (threads / 32);
  const long long cap = (long long)sm_count * 8;   // persistent-ish: at most 8 CTAs per SM, grid-stride beyond
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  if (tape) hb_k1_kernel<G, ITER, true><<<(unsigned)blocks, threads, 0, st>>>(p);
  else hb_k1_kernel<G, ITER, false><<<(unsigned)blocks, threads, 0, st>>>(p);
  return cudaGetLastError();
}

}  // namespace

int g_hb_force_wide = 0;

cudaError_t hb_launch_k1(const hb_k1_params& p, bool tape_mode, int sm_count, cudaStream_t st, bool* density_done) {
  const int d = p.d;
  if (density_done) *density_done = false;
  if (d <= 1) return launch_gi<1, 1>(p, tape_mode, sm_count, st);
  if (d <= 2) return launch_gi<2, 1>(p, tape_mode, sm_count, st);
  if (d <= 4) return launch_gi<4, 1>(p, tape_mode, sm_count, st);
  if (d <= 8) return launch_gi<8, 1>(p, tape_mode, sm_count, st);
  if (d <= 16) return launch_gi<16, 1>(p, tape_mode, sm_count, st);
  if (p.nc > 0x7fffffffLL || p.m_arch > 0x7fffffffLL) return cudaErrorInvalidValue;
  // small populations (fewer 32-chain batches than resident warps): one chain per warp spreads the work over more SMs
  if (p.n_peers > 1 && d <= 16) return cudaErrorInvalidValue;   // peer gather is implemented in the wide kernel
  if (p.n_peers <= 1 && !g_hb_force_wide && !p.force_wide && p.n_local < (long long)sm_count * 16 * 32 / 2) {
    if (d <= 32) return launch_gi<32, 1>(p, tape_mode, sm_count, st);
    if (d <= 64) return launch_gi<32, 2>(p, tape_mode, sm_count, st);
    if (d <= 128) return launch_gi<32, 4>(p, tape_mode, sm_count, st);
    if (d <= 256) return launch_gi<32, 8>(p, tape_mode, sm_count, st);
    if (d <= 512) return launch_gi<32, 16>(p, tape_mode, sm_count, st);
    return launch_gi<32, 32>(p, tape_mode, sm_count, st);
  }
  if (tape_mode) {
    if (d <= 32) return launch_wide_tape<1>(p, sm_count, st);
    if (d <= 64) return launch_wide_tape<2>(p, sm_count, st);
    if (d <= 128) return launch_wide_tape<4>(p, sm_count, st);
    if (d <= 256) return launch_wide_tape<8>(p, sm_count, st);
    if (d <= 512) return launch_wide_tape<16>(p, sm_count, st);
    if (d <= 1024) return launch_wide_tape<32>(p, sm_count, st);
    return cudaErrorInvalidValue;
  }
  if (d <= 64) return launch_wide_pair<1>(p, sm_count, st, density_done);      // ITER2 = ceil(ceil(d/2)/32) pairs per lane
  if (d <= 128) return launch_wide_pair<2>(p, sm_count, st, density_done);
  if (d <= 192) return launch_wide_pair<3>(p, sm_count, st, density_done);
  if (d <= 256) return launch_wide_pair<4>(p, sm_count, st, density_done);
  if (d <= 512) return launch_wide_pair<8>(p, sm_count, st, density_done);
  if (d <= 1024) return launch_wide_pair<16>(p, sm_count, st, density_done);
  return cudaErrorInvalidValue;
}

bool hb_fused_small_eligible(long long n_local, int d, int sm_count) {
  return d > 16 && d <= 128 && !g_hb_force_wide && n_local < (long long)sm_count * 16 * 32 / 2;   // where hb_launch_k1 picks <32, ITER>
}

cudaError_t hb_launch_persist_small(const hb_k1_params& p, const hb_k3_params& q, const hb_fused_tdist& td,
                                    const hb_persist_params& pp, bool tape_mode, int sm_count, cudaStream_t st) {
  if (!hb_fused_small_eligible(p.n_local, p.d, sm_count) || p.past_sample || p.n_peers > 1 || q.n_runs > 1 ||
      q.accept_in || q.lik22 || q.log_rows || pp.XA == pp.XB || pp.i1 <= pp.i0 || (pp.adapt && (!pp.partials || !pp.shift)))
    return cudaErrorInvalidValue;
  if (p.d <= 32) return launch_persist<1>(p, q, td, pp, tape_mode, sm_count, st);
  if (p.d <= 64) return launch_persist<2>(p, q, td, pp, tape_mode, sm_count, st);
  return launch_persist<4>(p, q, td, pp, tape_mode, sm_count, st);
}

cudaError_t hb_launch_fused_small(const hb_k1_params& p, const hb_k3_params& q, const hb_fused_tdist& td, bool tape_mode,
                                  int sm_count, cudaStream_t st) {
  if (!hb_fused_small_eligible(p.n_local, p.d, sm_count) || p.past_sample || p.n_peers > 1 || q.adapt || q.n_runs > 1 ||
      q.accept_in || q.lik22 || q.log_rows || q.Xout == q.X)
    return cudaErrorInvalidValue;
  if (p.d <= 32) return launch_fused<1>(p, q, td, tape_mode, sm_count, st);
  if (p.d <= 64) return launch_fused<2>(p, q, td, tape_mode, sm_count, st);
  return launch_fused<4>(p, q, td, tape_mode, sm_count, st);
}
