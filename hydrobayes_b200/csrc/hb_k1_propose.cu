// hb_k1_propose.cu -- K1: differential-evolution proposal for every chain of the population.
//
// Replaces calc_prop (R/internals.R:115-219) + bound_par (:89-113) + prior_pdf (:49-69) as called for all nc chains
// from the generation-start state by dream_parallel (R/dream_parallel.R:269-275).
//
// Mapping: G = min(32, pow2 >= d) consecutive lanes own one chain; lane lg owns dimensions k = m*G + lg,
// m < ITER.  A warp therefore reads a partner row as ITER fully coalesced 256-byte segments, and the rows of the
// 1 + 2D chains it touches are the only HBM traffic (48d + 16 bytes per chain-generation).  Loads of
// dimensions outside the crossover subspace are predicated off.  Random draws are either read from the decision
// tape or generated with Philox4x32-10 keyed by (seed; chain, generation, slot): one Philox call per dimension
// (zt, lambda, zeta) plus five per chain, the latter computed by five different lanes and broadcast.
#include "hb_kernels.h"

namespace {

// error-free sum of up to HB_MAX_DELTA pair differences: reproduces R's long-double colSums (R/internals.R:196)
// to the last bit except for (rare) double-rounding cases
__device__ __forceinline__ void two_sum(double a, double b, double& s, double& e) {
  s = __dadd_rn(a, b);
  const double bb = __dadd_rn(s, -a);
  e = __dadd_rn(__dadd_rn(a, -__dadd_rn(s, -bb)), __dadd_rn(b, -bb));
}

template <int G, int ITER, bool TAPE>
__global__ void __launch_bounds__(256) hb_k1_kernel(const hb_k1_params p) {
  constexpr unsigned FULL = 0xffffffffu;
  constexpr int CPW = 32 / G;
  const int lane = threadIdx.x & 31;
  const int lg = lane & (G - 1);
  const int gbase = lane & ~(G - 1);
  const unsigned gmask = hb_group_mask<G>(lane);
  const long long warp_id = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int d = p.d, ldx = p.ldx, nCR = p.nCR, delta = p.delta;
  const double* __restrict__ src = p.past_sample ? p.Z : p.X;
  unsigned errbits = 0;

  for (long long base = warp_id * CPW; base < p.n_local; base += n_warps * CPW) {
    const long long jl_raw = base + lane / G;
    const bool valid = jl_raw < p.n_local;
    const long long jl = valid ? jl_raw : p.n_local - 1;
    const long long j = p.chain0 + jl;           // chain within the run
    const long long gch = p.gchain0 + j;         // global chain id: Philox counter and tape column
    const long long rix = p.tape_g * p.tape.nct + gch;

    // ---------------- per-chain draws: snooker test, D, partners, crossover id, gamma choice ----------------
    double u_sn, g_sn = 0.0;
    int D, id, g_is_one;
    long long pa[HB_MAX_DELTA], pb[HB_MAX_DELTA];   // a = first D partners, b = next D (R/internals.R:143-144)
    long long ps2 = 0;                               // third snooker row (:127)
    const hb_phx ph(p.seed, (uint64_t)gch, p.gen);
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
        cum += p.ctrl->pCR[q];
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

    const double* __restrict__ xrow = p.X + j * (long long)ldx;
    double xk[ITER], dxk[ITER];
    double lp_part = 0.0;

    if (snooker) {
      // ---------------- snooker update, R/internals.R:150-163 with orth_proj :250-263 ----------------
      const double* r1 = src + pa[0] * (long long)ldx;
      const double* r2 = src + pb[0] * (long long)ldx;
      const double* r3 = src + ps2 * (long long)ldx;
      double uu[ITER], a2[ITER], a3[ITER];
      double s2 = 0, s3 = 0, su = 0;
      int differs = 0;
#pragma unroll
      for (int m = 0; m < ITER; ++m) {
        const int k = m * G + lg;
        xk[m] = 0; uu[m] = 0; a2[m] = 0; a3[m] = 0;
        if (k < d) {
          xk[m] = xrow[k];
          const double b = r1[k];
          a2[m] = r2[k]; a3[m] = r3[k];
          uu[m] = xk[m] - b;
          differs |= (xk[m] != b);
          s2 += (a2[m] - xk[m]) * uu[m];
          s3 += (a3[m] - xk[m]) * uu[m];
          su += uu[m] * uu[m];
        }
      }
      s2 = hb_group_sum<G>(s2, gmask); s3 = hb_group_sum<G>(s3, gmask); su = hb_group_sum<G>(su, gmask);
      differs = hb_group_sum_int<G>(differs, gmask);
#pragma unroll
      for (int m = 0; m < ITER; ++m) {
        const int k = m * G + lg;
        dxk[m] = 0;
        if (k < d) {
          double zeta;
          if (TAPE) zeta = p.tape.zeta[rix * d + k];
          else { const uint4 w = ph.slot(HB_SLOT_DIM0 + (uint32_t)k); zeta = hb_zeta(w.z, w.w, p.c_star); }
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
      const double CRv = (double)(id + 1) / (double)nCR;               // CR <- (1:nCR)/nCR
      double ztv[ITER];
      uint32_t wl[ITER], wz2[ITER], wz3[ITER];
      unsigned selmask[ITER];
      int d_star = 0;
#pragma unroll
      for (int m = 0; m < ITER; ++m) {
        const int k = m * G + lg;
        ztv[m] = 2.0; wl[m] = 0; wz2[m] = 0; wz3[m] = 0;
        if (k < d) {
          if (TAPE) ztv[m] = p.tape.zt[rix * d + k];                   // :175
          else {
            const uint4 w = ph.slot(HB_SLOT_DIM0 + (uint32_t)k);
            ztv[m] = hb_u32d(w.x); wl[m] = w.y; wz2[m] = w.z; wz3[m] = w.w;
          }
        }
        selmask[m] = hb_group_ballot<G>(k < d && ztv[m] < CRv, lane, gmask); // A <- which(zt < CR[id])  :177
        d_star += __popc(selmask[m]);
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
      const double gamma_d = 2.38 / sqrt(2.0 * (double)D * (double)d_star);   // :190
      const double gam = g_is_one ? 1.0 : gamma_d;                            // :192
      int rank_base = 0;
#pragma unroll
      for (int m = 0; m < ITER; ++m) {
        const int k = m * G + lg;
        const bool sel = (kmin >= 0) ? (k == kmin) : ((selmask[m] >> lg) & 1u);
        xk[m] = 0; dxk[m] = 0;
        if (k < d) {
          xk[m] = xrow[k];
          if (sel) {
            double lambda, zeta;
            if (TAPE) {
              const int rank = (kmin >= 0) ? 0 : rank_base + __popc(selmask[m] & ((1u << lg) - 1u));
              lambda = p.tape.lambda[rix * d + rank];                  // :188 (by rank within A)
              zeta = p.tape.zeta[rix * d + rank];                      // :194
            } else {
              lambda = hb_rng_lambda(wl[m], p.c_val);
              zeta = hb_zeta(wz2[m], wz3[m], p.c_star);
            }
            // jump_diff <- colSums(r1[,A] - r2[,A])   :196
            double s = 0.0, comp = 0.0;
#pragma unroll
            for (int q = 0; q < HB_MAX_DELTA; ++q) {
              if (q < D) {
                const double a = src[pa[q] * (long long)ldx + k];
                const double b = src[pb[q] * (long long)ldx + k];
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
        rank_base += __popc(selmask[m]);
      }
    }

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

cudaError_t hb_launch_k1(const hb_k1_params& p, bool tape_mode, int sm_count, cudaStream_t st) {
  const int d = p.d;
  if (d <= 1) return launch_gi<1, 1>(p, tape_mode, sm_count, st);
  if (d <= 2) return launch_gi<2, 1>(p, tape_mode, sm_count, st);
  if (d <= 4) return launch_gi<4, 1>(p, tape_mode, sm_count, st);
  if (d <= 8) return launch_gi<8, 1>(p, tape_mode, sm_count, st);
  if (d <= 16) return launch_gi<16, 1>(p, tape_mode, sm_count, st);
  if (d <= 32) return launch_gi<32, 1>(p, tape_mode, sm_count, st);
  if (d <= 64) return launch_gi<32, 2>(p, tape_mode, sm_count, st);
  if (d <= 128) return launch_gi<32, 4>(p, tape_mode, sm_count, st);
  if (d <= 256) return launch_gi<32, 8>(p, tape_mode, sm_count, st);
  if (d <= 512) return launch_gi<32, 16>(p, tape_mode, sm_count, st);
  if (d <= 1024) return launch_gi<32, 32>(p, tape_mode, sm_count, st);
  return cudaErrorInvalidValue;
}
