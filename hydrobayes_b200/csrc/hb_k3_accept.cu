// hb_k3_accept.cu -- K3: Metropolis accept/reject + state update + history store + adaptation statistics,
// the finalize step (J, n_id, pCR adaptation, AR/CR rows) and the outlier-reset pieces.
//
// Replaces metropolis_acceptance (R/internals.R:221-247), the update block R/dream_parallel.R:348-378, the store
// block :386-399, the pCR adaptation :403-411, check_outlier's copy (R/internals.R:82-83) and the AR/CR rows :425-433.
//
// Statistics: std_x (R/dream_parallel.R:376) is the post-update sample sd over chains.  K3 accumulates, per block and
// in a fixed order, the shifted column sums  sum(x - c), sum((x - c)^2)  (c = last generation's column mean) and
// Q[id][k] = sum over accepted chains with crossover index id of dx_k^2; the finalize kernel reduces the block
// partials in block order (deterministic), forms var_k and J[id] += sum_k Q[id][k] / var_k, which equals
// sum_j sum_k (dx_jk / std_k)^2 of :377-378.  All of this only runs inside the adaptation period: J and n_id are
// unobservable afterwards (:403).
#include "hb_kernels.h"

namespace {

template <int G, int ITER, bool TAPE>
__global__ void __launch_bounds__(256) hb_k3_kernel(const hb_k3_params p) {
  constexpr unsigned FULL = 0xffffffffu;
  constexpr int CPW = 32 / G;
  extern __shared__ __align__(16) double smem[];
  __shared__ unsigned long long s_cnt[2 * HB_MAX_NCR + 1];
  const int lane = threadIdx.x & 31;
  const int lg = lane & (G - 1);
  const int gbase = lane & ~(G - 1);
  const int warp_in_block = threadIdx.x >> 5;
  const int slot = warp_in_block * CPW + lane / G;          // accumulator slot of this group
  const int n_slots = (blockDim.x >> 5) * CPW;
  const long long warp_id = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int d = p.d, ldx = p.ldx, nCR = p.nCR;
  const bool adapt = p.adapt != 0;
  const int E = (nCR + 2) * d;                              // accumulator elements per slot
  unsigned errbits = 0;

  if (threadIdx.x <= 2 * HB_MAX_NCR) s_cnt[threadIdx.x] = 0ull;
  if (adapt && !p.write_dx) {
    for (int e = threadIdx.x; e < n_slots * E; e += blockDim.x) smem[e] = 0.0;
  }
  __syncthreads();
  double* acc = smem + (size_t)slot * E;                    // [ (nCR+2) ][ d ] : rows 0,1 = s1,s2 ; 2+q = Q[q]
  double s1[ITER], s2[ITER], sh[ITER];
#pragma unroll
  for (int m = 0; m < ITER; ++m) {
    s1[m] = 0; s2[m] = 0;
    const int k = m * G + lg;
    sh[m] = (adapt && k < d) ? p.shift[k] : 0.0;
  }

  for (long long base = warp_id * CPW; base < p.n_local; base += n_warps * CPW) {
    const long long jl = base + lane / G;
    const bool valid = jl < p.n_local;
    const long long j = p.chain0 + (valid ? jl : 0) * p.stride;   // chain index within the handle
    const long long run = p.n_runs > 1 ? j / p.nc : 0;
    const long long si = j - p.state0;                            // index into the state / history arrays
    const long long gch = p.gchain0 + j;
    int accf = 0, id = 0;
    double lpx = 0, lln = 0, lpostx = 0, lpost_old = 0;
    if (valid && lg == 0) {
      lpx = p.lp_xp[jl];
      double llr = p.ll_raw[jl];
      if (!(llr >= HB_FLOOR)) llr = (llr != llr) ? llr : HB_FLOOR;    // calc_ll floor, R/internals.R:19
      if (!isfinite(llr)) errbits |= HB_FLAG_LL_NONFINITE;            // :21
      lln = llr;
      lpostx = lpx + lln;                                             // R/dream_parallel.R:291
      lpost_old = p.lpost[si];
      double u;
      if (TAPE) u = p.u_accept[j];
      else { const hb_phx ph(p.seed, (uint64_t)gch, p.gen); u = hb_u53(hb_pair64(ph.slot(HB_SLOT_ACCEPT), 0)); }
      const double p_acc = fmin(fmax(-100.0, lpostx - lpost_old), 0.0); // R/internals.R:239
      accf = p.accept_in ? (int)p.accept_in[jl] : p.lik22 ? ((lpostx >= lpost_old) || (lpostx >= 0.0)) : (p_acc > log(u));   // :241 ; ABC rule :224 ; mt > 1: hb_mt.cu
      id = p.id[jl];
    }
    int log_slot = -1;                                                // delta log: slot of this chain's accepted row
    if (p.log_rows && valid && lg == 0 && accf) { log_slot = (int)atomicAdd(p.log_count, 1u); p.log_idx[log_slot] = (int)j; }
    if (G > 1) { accf = __shfl_sync(FULL, accf, gbase); id = __shfl_sync(FULL, id, gbase); log_slot = __shfl_sync(FULL, log_slot, gbase); }
    if (log_slot >= 0)                                                // the padding columns travel with the row: keep them finite
      for (int k = d + lg; k < ldx; k += G) p.log_rows[(long long)log_slot * ldx + k] = 0.0;

    const bool copy_all = p.Xout != p.X;
    const bool need_x = adapt || p.hist_x != nullptr || copy_all;
#pragma unroll
    for (int m = 0; m < ITER; ++m) {
      const int k = m * G + lg;
      if (valid && k < d) {
        double xn = 0.0;
        if (accf) {
          xn = p.XP[jl * (long long)ldx + k];
          if (adapt) {
            const double dx = xn - p.X[j * (long long)ldx + k];       // R/dream_parallel.R:357 (after bound handling)
            if (p.write_dx) p.XP_rw[jl * (long long)ldx + k] = dx;
            else acc[(2 + id) * d + k] += dx * dx;
          }
          if (log_slot >= 0) p.log_rows[(long long)log_slot * ldx + k] = xn;   // :358 deferred: hb_delta_apply_kernel
          else p.Xout[j * (long long)ldx + k] = xn;                   // :358
        } else {
          if (need_x) xn = p.X[j * (long long)ldx + k];
          if (copy_all) p.Xout[j * (long long)ldx + k] = xn;
          if (adapt && p.write_dx) p.XP_rw[jl * (long long)ldx + k] = 0.0;   // dx stays 0, R/dream_parallel.R:266
        }
        if (adapt && !p.write_dx) { const double c = xn - sh[m]; s1[m] += c; s2[m] += c * c; }
        if (p.hist_x) p.hist_x[si * (long long)d + k] = xn;           // chain[ind_out, 1:d, ] :388
      }
    }
    if (valid && lg == 0) {
      double lp_c, ll_c, lpost_c;
      if (accf) {
        p.lp[si] = lpx; p.ll[si] = lln; p.lpost[si] = lpostx;         // :359-361
        if (p.fx && p.fx_xp) p.fx[si] = p.fx_xp[jl];                  // :362
        lp_c = lpx; ll_c = lln; lpost_c = lpostx;
      } else {
        lpost_c = lpost_old;                                          // :365
        lp_c = 0; ll_c = 0;
        if (p.hist_l) { lp_c = p.lp[si]; ll_c = p.ll[si]; }
      }
      if (p.hist_l) {                                                 // :389-391
        p.hist_l[si * 3 + 0] = lp_c; p.hist_l[si * 3 + 1] = ll_c; p.hist_l[si * 3 + 2] = lpost_c;
      }
      if (p.hist_fx && p.fx) p.hist_fx[si] = p.fx[si];                // :394
      if (p.lpost_hist_row) p.lpost_hist_row[si] = lpost_c;
      if (p.accept_rec) p.accept_rec[si] = (uint8_t)accf;
      if (p.n_runs > 1) {                                             // batched runs: counters live in the run's block
        atomicAdd(&p.ctrl[run].n_id_gen[id], 1ull);
        if (accf) { atomicAdd(&p.ctrl[run].n_acc_gen, 1ull); atomicAdd(&p.ctrl[run].n_acc_id_gen[id], 1ull); }
      } else {
        atomicAdd(&s_cnt[1 + id], 1ull);                              // n_id  :368-369
        if (accf) { atomicAdd(&s_cnt[0], 1ull); atomicAdd(&s_cnt[1 + HB_MAX_NCR + id], 1ull); }   // n_acc :374
      }
    }
  }

  const bool block_stats = adapt && !p.write_dx;
  if (block_stats) {
#pragma unroll
    for (int m = 0; m < ITER; ++m) {
      const int k = m * G + lg;
      if (k < d) { acc[k] = s1[m]; acc[d + k] = s2[m]; }
    }
  }
  __syncthreads();
  if (block_stats) {
    // fixed-order reduction over the block's slots -> one partial row per block
    for (int e = threadIdx.x; e < E; e += blockDim.x) {
      double s = 0.0;
      for (int q = 0; q < n_slots; ++q) s += smem[(size_t)q * E + e];
      p.partials[(size_t)blockIdx.x * E + e] = s;
    }
  }
  if (threadIdx.x <= 2 * HB_MAX_NCR) {
    const unsigned long long v = s_cnt[threadIdx.x];
    if (v) {
      if (threadIdx.x == 0) atomicAdd(&p.ctrl->n_acc_gen, v);
      else if (threadIdx.x <= HB_MAX_NCR) atomicAdd(&p.ctrl->n_id_gen[threadIdx.x - 1], v);
      else atomicAdd(&p.ctrl->n_acc_id_gen[threadIdx.x - 1 - HB_MAX_NCR], v);
    }
  }
  if (errbits) atomicOr(&p.ctrl->err, errbits);
}

// ------------------------------------------------------------------------------------------------------------------
// Wide variant (d > 16): a warp owns a batch of 32 chains.
//   phase A  lane-per-chain: coalesced loads of lp_xp / ll_raw / lpost / id / u, the Metropolis decision, and the
//            coalesced scalar stores (lp, ll, lpost, history scalars, accept mask);
//   phase B  only for the rows that need touching -- accepted rows (16 % for the t-distribution) outside the adaptation
//            period, every row inside it (std_x needs the whole post-update population) or when the generation is
//            stored.  The rows move through a per-warp ring of shared-memory slots by 1-D bulk TMA: up to eight row
//            loads are in flight per warp (XP row of an accepted chain, X row otherwise), and an accepted row leaves
//            the slot it landed in as bulk stores into X and -- HB_EXCHANGE_DELTA -- into every peer's replica over
//            NVLink, so in the steady state the row data never passes through registers.  The lanes read the slots
//            (lane owns columns k = 32 m + lane) only for what needs arithmetic: the adaptation statistics, the
//            realised jump dx, the history copy.
// ------------------------------------------------------------------------------------------------------------------
constexpr int K3_SLOTS = 8;

__host__ __device__ inline size_t k3_warp_smem_bytes(int ldx) { return (size_t)K3_SLOTS * ldx * 8 + K3_SLOTS * 8; }

template <int ITER, bool TAPE>
__global__ void __launch_bounds__(256, 3) hb_k3_wide_kernel(const hb_k3_params p, int warp_smem, int acc_off) {
  constexpr unsigned FULL = 0xffffffffu;
  extern __shared__ __align__(16) unsigned char k3_smem[];
  __shared__ unsigned long long s_cnt[2 * HB_MAX_NCR + 1];
  const int lane = threadIdx.x & 31;
  const int slot = threadIdx.x >> 5;
  const int n_slots = blockDim.x >> 5;
  const long long warp_id = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int d = p.d, ldx = p.ldx, nCR = p.nCR;
  const bool adapt = p.adapt != 0;
  const bool stats = adapt && !p.write_dx;
  const int E = (nCR + 2) * d;
  const uint32_t row_bytes = (uint32_t)ldx * 8u;
  unsigned errbits = 0;

  double* ring = reinterpret_cast<double*>(k3_smem + (size_t)slot * warp_smem);          // [K3_SLOTS][ldx]
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + (size_t)K3_SLOTS * ldx);           // [K3_SLOTS]
  // jump-statistic accumulators Q[id][k] of this warp (stats only).  The column sums s1 / s2 live in registers and are parked
  // in the warp's first two ring rows at the end: 2.4 KB of accumulators per warp instead of 4 KB is what lets a third CTA
  // share the SM (the adaptation-period launch is bound by the rows it keeps in flight per SM, profiles/r02b_k3_adaptation_full.txt)
  const int EQ = nCR * d;
  double* acc_all = reinterpret_cast<double*>(k3_smem + acc_off);                        // [n_slots][EQ]
  double* acc = acc_all + (size_t)slot * EQ;

  if (threadIdx.x <= 2 * HB_MAX_NCR) s_cnt[threadIdx.x] = 0ull;
  if (stats) for (int e = threadIdx.x; e < n_slots * EQ; e += blockDim.x) acc_all[e] = 0.0;
  if (lane == 0) {
#pragma unroll
    for (int q = 0; q < K3_SLOTS; ++q) mbar_init(&bars[q], 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();
  unsigned phase_bits = 0;                              // parity of the next completion of barrier q
  bool stores_pending = false;                          // lane 0: bulk stores of the previous group may still read the ring
  double s1[ITER], s2[ITER], sh[ITER];
#pragma unroll
  for (int m = 0; m < ITER; ++m) {
    s1[m] = 0; s2[m] = 0;
    const int k = m * 32 + lane;
    sh[m] = (adapt && k < d) ? p.shift[k] : 0.0;
  }
  const bool copy_all = p.Xout != p.X;                  // double-buffered shard: rejected rows are carried over
  const bool all_rows = adapt || p.hist_x != nullptr || copy_all;   // (write_dx needs every row too: dx = 0)
  const bool need_old = adapt;                          // accepted rows also need the old state: dx = xp - x  (:357)
  const bool lanes_read = adapt || p.hist_x != nullptr || copy_all; // somebody needs the values in registers
  const int group = need_old ? K3_SLOTS / 2 : K3_SLOTS; // rows per group (accepted rows take two slots in adaptation)

  for (long long base = warp_id * 32; base < p.n_local; base += n_warps * 32) {
    const long long jl = base + lane;
    const bool valid = jl < p.n_local;
    int accf = 0, id = 0;
    const long long gj = p.chain0 + (valid ? jl : 0) * p.stride;   // chain index within the handle
    const long long run = p.n_runs > 1 ? gj / p.nc : 0;
    const long long si = gj - p.state0;
    if (valid) {
      const long long gch = p.gchain0 + gj;
      const double lpx = p.lp_xp[jl];
      double llr = p.ll_raw[jl];
      if (!(llr >= HB_FLOOR)) llr = (llr != llr) ? llr : HB_FLOOR;    // calc_ll floor, R/internals.R:19
      if (!isfinite(llr)) errbits |= HB_FLAG_LL_NONFINITE;            // :21
      const double lpostx = lpx + llr;                                // R/dream_parallel.R:291
      const double lpost_old = p.lpost[si];
      double u;
      if (TAPE) u = p.u_accept[gj];
      else { const hb_phx ph(p.seed, (uint64_t)gch, p.gen); u = hb_u53(hb_pair64(ph.slot(HB_SLOT_ACCEPT), 0)); }
      const double p_acc = fmin(fmax(-100.0, lpostx - lpost_old), 0.0); // R/internals.R:239
      accf = p.accept_in ? (int)p.accept_in[jl] : p.lik22 ? ((lpostx >= lpost_old) || (lpostx >= 0.0)) : (p_acc > log(u));   // :241 ; ABC rule :224 ; mt > 1: hb_mt.cu
      id = p.id[jl];
      double lp_c = 0, ll_c = 0, lpost_c = lpost_old;                 // :365
      if (accf) {
        p.lp[si] = lpx; p.ll[si] = llr; p.lpost[si] = lpostx;         // :359-361
        if (p.fx && p.fx_xp) p.fx[si] = p.fx_xp[jl];                  // :362
        lp_c = lpx; ll_c = llr; lpost_c = lpostx;
      } else if (p.hist_l) { lp_c = p.lp[si]; ll_c = p.ll[si]; }
      if (p.hist_l) { p.hist_l[si * 3 + 0] = lp_c; p.hist_l[si * 3 + 1] = ll_c; p.hist_l[si * 3 + 2] = lpost_c; }   // :389-391
      if (p.hist_fx && p.fx) p.hist_fx[si] = p.fx[si];                // :394
      if (p.lpost_hist_row) p.lpost_hist_row[si] = lpost_c;
      if (p.accept_rec) p.accept_rec[si] = (uint8_t)accf;
      if (p.n_runs > 1) {                                             // batched runs: counters live in the run's block
        atomicAdd(&p.ctrl[run].n_id_gen[id], 1ull);
        if (accf) { atomicAdd(&p.ctrl[run].n_acc_gen, 1ull); atomicAdd(&p.ctrl[run].n_acc_id_gen[id], 1ull); }
      }
    }
    const unsigned vmask = __ballot_sync(FULL, valid);
    const unsigned amask = __ballot_sync(FULL, valid && accf);
    int log_base = 0;                                                 // delta log: popc(amask) consecutive slots for this batch
    if (p.log_rows && amask) {
      if (lane == 0) log_base = (int)atomicAdd(p.log_count, (unsigned)__popc(amask));
      log_base = __shfl_sync(FULL, log_base, 0);
      if (valid && accf) p.log_idx[log_base + __popc(amask & ((1u << lane) - 1u))] = (int)gj;
    }
    if (p.n_runs == 1) {
      for (int q = 0; q < nCR; ++q) {                                 // n_id  :368-369
        const unsigned m_q = __ballot_sync(FULL, valid && id == q);
        const unsigned a_q = m_q & amask;
        if (lane == 0 && m_q) atomicAdd(&s_cnt[1 + q], (unsigned long long)__popc(m_q));
        if (lane == 0 && a_q) atomicAdd(&s_cnt[1 + HB_MAX_NCR + q], (unsigned long long)__popc(a_q));
      }
      if (lane == 0 && amask) atomicAdd(&s_cnt[0], (unsigned long long)__popc(amask));   // n_acc :374
    }

    unsigned rows = all_rows ? vmask : amask;
    while (rows) {
      // ---- this group's rows: lane q < n_g owns row c_q (ascending chain order) ----
      int my_c = -1, n_g = 0;
      {
        unsigned rr = rows;
#pragma unroll
        for (int q = 0; q < K3_SLOTS; ++q)
          if (q < group && rr) { const int c = __ffs(rr) - 1; rr &= rr - 1; if (lane == q) my_c = c; n_g = q + 1; }
        rows = rr;
      }
      if (stores_pending) { if (lane == 0) tma_store_wait_read<0>(); stores_pending = false; }
      __syncwarp();
      if (lane < n_g) {                                               // issue the row loads of the group
        const int a_c = (amask >> my_c) & 1;
        const long long jr = base + my_c;
        const long long j = p.chain0 + jr * p.stride;
        const int sl = need_old ? 2 * lane : lane;
        const bool two = a_c && need_old;
        mbar_expect_tx(&bars[lane], two ? 2 * row_bytes : row_bytes);
        tma_load_1d(ring + (size_t)sl * ldx, a_c ? p.XP + jr * (long long)ldx : p.X + j * (long long)ldx, row_bytes, &bars[lane]);
        if (two) tma_load_1d(ring + (size_t)(sl + 1) * ldx, p.X + j * (long long)ldx, row_bytes, &bars[lane]);
      }
      // ---- consume ----
#pragma unroll 1
      for (int q = 0; q < n_g; ++q) {
        const int c = __shfl_sync(FULL, my_c, q);
        const int a_c = (amask >> c) & 1;
        const int id_c = __shfl_sync(FULL, id, c);
        const long long jr = base + c;
        const long long j = p.chain0 + jr * p.stride;
        const long long sr = j - p.state0;
        const int sl = need_old ? 2 * q : q;
        const double* __restrict__ rowv = ring + (size_t)sl * ldx;
        mbar_wait(&bars[q], (phase_bits >> q) & 1u);
        phase_bits ^= 1u << q;
        if (lanes_read) {
#pragma unroll
          for (int m = 0; m < ITER; ++m) {
            const int k = m * 32 + lane;
            if (k < d) {
              const double xn = rowv[k];
              if (a_c) {
                if (adapt) {
                  const double dx = xn - rowv[ldx + k];                 // R/dream_parallel.R:357 (after bound handling)
                  if (p.write_dx) p.XP_rw[jr * (long long)ldx + k] = dx;
                  else acc[id_c * d + k] += dx * dx;
                }
              } else {
                if (copy_all) p.Xout[j * (long long)ldx + k] = xn;
                if (adapt && p.write_dx) p.XP_rw[jr * (long long)ldx + k] = 0.0;
              }
              if (stats) { const double cc = xn - sh[m]; s1[m] += cc; s2[m] += cc * cc; }
              if (p.hist_x) p.hist_x[sr * (long long)d + k] = xn;       // chain[ind_out, 1:d, ] :388
            }
          }
        }
        if (a_c && lane == 0) {                                         // :358  xt[accept,] <- xp[accept,]
          if (p.log_rows) tma_store_1d(p.log_rows + (long long)(log_base + __popc(amask & ((1u << c) - 1u))) * ldx, rowv, row_bytes);
          else tma_store_1d(p.Xout + j * (long long)ldx, rowv, row_bytes);
          tma_store_commit();
          stores_pending = true;
        }
      }
      stores_pending = __shfl_sync(FULL, (int)stores_pending, 0) != 0;
      __syncwarp();                                                     // every lane is done reading the ring
    }
  }
  if (lane == 0) tma_store_wait<0>();                                   // all rows have left before the block retires

  if (stats) {
    __syncwarp();                                                       // lane 0's wait: no bulk copy touches the ring any more
#pragma unroll
    for (int m = 0; m < ITER; ++m) {
      const int k = m * 32 + lane;
      if (k < d) { ring[k] = s1[m]; ring[ldx + k] = s2[m]; }
    }
  }
  __syncthreads();
  if (stats) {
    // fixed-order reduction over the block's warps -> one partial row per block: [s1 | s2 | Q[0] .. Q[nCR-1]]
    for (int e = threadIdx.x; e < E; e += blockDim.x) {
      double s = 0.0;
      if (e < 2 * d) {
        const int off = e < d ? e : ldx + (e - d);
        for (int q = 0; q < n_slots; ++q) s += reinterpret_cast<const double*>(k3_smem + (size_t)q * warp_smem)[off];
      } else {
        for (int q = 0; q < n_slots; ++q) s += acc_all[(size_t)q * EQ + (e - 2 * d)];
      }
      p.partials[(size_t)blockIdx.x * E + e] = s;
    }
  }
  if (threadIdx.x <= 2 * HB_MAX_NCR) {
    const unsigned long long v = s_cnt[threadIdx.x];
    if (v) {
      if (threadIdx.x == 0) atomicAdd(&p.ctrl->n_acc_gen, v);
      else if (threadIdx.x <= HB_MAX_NCR) atomicAdd(&p.ctrl->n_id_gen[threadIdx.x - 1], v);
      else atomicAdd(&p.ctrl->n_acc_id_gen[threadIdx.x - 1 - HB_MAX_NCR], v);
    }
  }
  if (errbits) atomicOr(&p.ctrl->err, errbits);
}

template <int ITER>
cudaError_t launch_wide(const hb_k3_params& p, int sm_count, int* nblocks_out, size_t* smem_out, cudaStream_t st) {
  const size_t per_warp_ring = (k3_warp_smem_bytes(p.ldx) + 15) & ~(size_t)15;
  const size_t per_warp_acc = (size_t)p.nCR * p.d * sizeof(double);       // Q[id][k]; s1 / s2 are parked in the ring
  int warps = 8;
  while (warps > 1 && (per_warp_ring + per_warp_acc) * warps > 100 * 1024) warps >>= 1;
  if ((per_warp_ring + per_warp_acc) * warps > 220 * 1024) return cudaErrorInvalidValue;
  const int threads = warps * 32;
  const long long warps_needed = (p.n_local + 31) / 32;
  long long blocks = (warps_needed + warps - 1) / warps;
  // one resident wave per launch.  The block count reported to the caller (the finalize kernel reduces that many
  // partial rows) is the adaptation-period one; launches outside the adaptation period write no partials and may use
  // the larger grid their smaller shared-memory footprint allows.
  auto wave = [&](size_t smem_block) {
    long long per_sm = (long long)(224 * 1024 / (smem_block + 1024));
    if (per_sm > 3) per_sm = 3;                              // __launch_bounds__(256, 3)
    if (per_sm < 1) per_sm = 1;
    long long b = blocks;
    if (b > (long long)sm_count * per_sm) b = (long long)sm_count * per_sm;
    return b < 1 ? 1 : b;
  };
  const size_t smem_full = (per_warp_ring + per_warp_acc) * warps;
  if (nblocks_out) *nblocks_out = (int)wave(smem_full);
  if (smem_out) *smem_out = smem_full;
  if (p.X == nullptr) return cudaSuccess;                   // geometry query only
  const bool st_acc = p.adapt && !p.write_dx;
  const size_t smem = per_warp_ring * warps + (st_acc ? per_warp_acc * warps : 0);
  blocks = wave(st_acc ? smem_full : smem);
  const bool tape = p.u_accept != nullptr;
  cudaError_t e = cudaSuccess;
  if (tape) {
    if (smem > 48 * 1024) e = cudaFuncSetAttribute(hb_k3_wide_kernel<ITER, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    hb_k3_wide_kernel<ITER, true><<<(unsigned)blocks, threads, smem, st>>>(p, (int)per_warp_ring, (int)(per_warp_ring * warps));
  } else {
    if (smem > 48 * 1024) e = cudaFuncSetAttribute(hb_k3_wide_kernel<ITER, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    hb_k3_wide_kernel<ITER, false><<<(unsigned)blocks, threads, smem, st>>>(p, (int)per_warp_ring, (int)(per_warp_ring * warps));
  }
  return cudaGetLastError();
}

template <int G>
struct k3_geom {
  static constexpr int CPW = 32 / G;
};

template <int G, int ITER>
cudaError_t launch_gi(const hb_k3_params& p, int sm_count, int* nblocks_out, size_t* smem_out, cudaStream_t st) {
  constexpr int CPW = 32 / G;
  const size_t per_warp = (size_t)CPW * (p.nCR + 2) * p.d * sizeof(double);
  int warps = 8;
  const size_t budget = 160 * 1024;
  while (warps > 1 && per_warp * warps > budget) warps >>= 1;
  const size_t smem_adapt = per_warp * warps;
  if (smem_adapt > 220 * 1024) return cudaErrorInvalidValue;
  const int threads = warps * 32;
  const long long warps_needed = (p.n_local + CPW - 1) / CPW;
  long long blocks = (warps_needed + warps - 1) / warps;
  const long long cap = (long long)sm_count * 4;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  if (nblocks_out) *nblocks_out = (int)blocks;
  if (smem_out) *smem_out = smem_adapt;
  if (p.X == nullptr) return cudaSuccess;                   // geometry query only
  const size_t smem = (p.adapt && !p.write_dx) ? smem_adapt : 0;
  const bool tape = p.u_accept != nullptr;
  cudaError_t e = cudaSuccess;
  if (tape) {
    if (smem > 48 * 1024) e = cudaFuncSetAttribute(hb_k3_kernel<G, ITER, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    hb_k3_kernel<G, ITER, true><<<(unsigned)blocks, threads, smem, st>>>(p);
  } else {
    if (smem > 48 * 1024) e = cudaFuncSetAttribute(hb_k3_kernel<G, ITER, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    hb_k3_kernel<G, ITER, false><<<(unsigned)blocks, threads, smem, st>>>(p);
  }
  return cudaGetLastError();
}

cudaError_t dispatch(const hb_k3_params& p, int sm_count, int* nb, size_t* sm, cudaStream_t st) {
  const int d = p.d;
  if (d <= 1) return launch_gi<1, 1>(p, sm_count, nb, sm, st);
  if (d <= 2) return launch_gi<2, 1>(p, sm_count, nb, sm, st);
  if (d <= 4) return launch_gi<4, 1>(p, sm_count, nb, sm, st);
  if (d <= 8) return launch_gi<8, 1>(p, sm_count, nb, sm, st);
  if (d <= 16) return launch_gi<16, 1>(p, sm_count, nb, sm, st);
  if (!g_hb_force_wide && !p.force_wide && p.n_local < (long long)sm_count * 16 * 32 / 2) {   // small populations: one chain per warp (see hb_launch_k1)
    if (d <= 32) return launch_gi<32, 1>(p, sm_count, nb, sm, st);
    if (d <= 64) return launch_gi<32, 2>(p, sm_count, nb, sm, st);
    if (d <= 128) return launch_gi<32, 4>(p, sm_count, nb, sm, st);
    if (d <= 256) return launch_gi<32, 8>(p, sm_count, nb, sm, st);
    if (d <= 512) return launch_gi<32, 16>(p, sm_count, nb, sm, st);
    return launch_gi<32, 32>(p, sm_count, nb, sm, st);
  }
  if (d <= 32) return launch_wide<1>(p, sm_count, nb, sm, st);
  if (d <= 64) return launch_wide<2>(p, sm_count, nb, sm, st);
  if (d <= 128) return launch_wide<4>(p, sm_count, nb, sm, st);
  if (d <= 256) return launch_wide<8>(p, sm_count, nb, sm, st);
  if (d <= 512) return launch_wide<16>(p, sm_count, nb, sm, st);
  if (d <= 1024) return launch_wide<32>(p, sm_count, nb, sm, st);
  return cudaErrorInvalidValue;
}

// ---- finalize -------------------------------------------------------------------------------------------
// block partials [nblocks][E] -> colsum / qsum, in a fixed order (deterministic): 32 columns x 8 row groups per CTA, the
// eight group sums of a column are added in group order
__global__ void __launch_bounds__(256) hb_finalize_reduce_kernel(const hb_fin_params p) {
  __shared__ double part[8][33];
  const int d = p.d, E = (p.nCR + 2) * d;
  const int c = threadIdx.x & 31, g = threadIdx.x >> 5;
  const int e = blockIdx.x * 32 + c;
  const double* partials = p.partials + (size_t)blockIdx.y * p.nblocks * E;     // blockIdx.y: slot (generation of an interval)
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  if (e < E) {
    int b = g;
    for (; b + 24 < p.nblocks; b += 32) {                   // four independent loads in flight
      s0 += partials[(size_t)b * E + e]; s1 += partials[(size_t)(b + 8) * E + e];
      s2 += partials[(size_t)(b + 16) * E + e]; s3 += partials[(size_t)(b + 24) * E + e];
    }
    for (; b < p.nblocks; b += 8) s0 += partials[(size_t)b * E + e];
  }
  part[g][c] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  if (g == 0 && e < E) {
    double s = 0.0;
#pragma unroll
    for (int q = 0; q < 8; ++q) s += part[q][c];
    const size_t so = (size_t)blockIdx.y * p.slot_stride;
    if (e < 2 * d) p.colsum[so + e] = s; else p.qsum[so + e - 2 * d] = s;
  }
}

__global__ void __launch_bounds__(256) hb_finalize_kernel(const hb_fin_params p) {
  __shared__ double red[256];
  const int d = p.d, nCR = p.nCR;
  hb_ctrl* c = p.ctrl;
  if (p.phase == 0) return;
  __syncthreads();
  // Multi-GPU: the per-generation sums of one updateInterval are parked in slots and all-reduced once (nothing between
  // two pCR updates consumes J, n_id or std_x: R/dream_parallel.R:403-411); they are folded here in generation order.
  const int n_slots = p.adapt ? (p.n_slots > 0 ? p.n_slots : 1) : 0;
  for (int sl = 0; sl < n_slots; ++sl) {
    const double* colsum = p.colsum + (size_t)sl * p.slot_stride;
    const double* qsum = p.qsum + (size_t)sl * p.slot_stride;
    // var_k of the post-update population (n-1 denominator), R/dream_parallel.R:376
    const double n = (double)p.nc;
    for (int q = 0; q < nCR; ++q) {
      double part = 0.0;
      for (int k = threadIdx.x; k < d; k += blockDim.x) {
        const double a = colsum[k], b = colsum[d + k];
        const double var = (b - a * a / n) / (n - 1.0);
        part += qsum[q * d + k] / var;                      // sum_j (dx_jk/std_k)^2 over chains with id == q
      }
      red[threadIdx.x] = part;
      __syncthreads();
      for (int o = blockDim.x / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
      }
      if (threadIdx.x == 0) c->J[q] += red[0];              // :377-378
      __syncthreads();
    }
    // next generation's shift (the sums of every slot of an interval were taken against the same, interval-start shift)
    if (sl == n_slots - 1) for (int k = threadIdx.x; k < d; k += blockDim.x) p.shift[k] += colsum[k] / n;
  }
  if (threadIdx.x == 0) hb_fin_counters(p);
}

// ---- outlier pieces -------------------------------------------------------------------------------------
// proxy <- colMeans(lpost[ceiling(i/2):i, ])  R/internals.R:73 ; Kahan-compensated column sum
__global__ void hb_proxy_kernel(const double* __restrict__ hist, long long n, long long row0, long long row1,
                                long long ld, double* __restrict__ proxy) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  double s = 0.0, c = 0.0;
  long long r = row0;
  for (; r + 8 <= row1 + 1; r += 8) {            // eight loads in flight; the compensated sum runs in row order as before
    double v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = hist[(r + u) * ld + j];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const double y = v[u] - c;
      const double t = s + y;
      c = (t - s) - y;
      s = t;
    }
  }
  for (; r <= row1; ++r) {
    const double y = hist[r * ld + j] - c;
    const double t = s + y;
    c = (t - s) - y;
    s = t;
  }
  proxy[j] = s / (double)(row1 - row0 + 1);
}

// x[outliers,] <- x[new_states,] ; dens[nrow, outliers] <- dens[nrow, new_states]  (R/internals.R:82-83).
// new_states never contains an outlier, so sources are not modified by this kernel.
// X is the full population replica (every rank applies the same copies); lpost / hist_row are the local shard
// [chain0, chain0 + n_local); lpost_src holds the pre-reset lpost of ALL chains (the local array on one GPU, the
// all-gathered copy on several).
__global__ void hb_outlier_reset_kernel(double* X, int ldx, int d, double* lpost, double* hist_row,
                                        const double* lpost_src, long long chain0, long long n_local, const int* outl,
                                        const int* news, int n_out) {
  const int q = blockIdx.x;
  if (q >= n_out) return;
  const long long dst = outl[q], srcc = news[q];
  for (int k = threadIdx.x; k < d; k += blockDim.x) X[dst * ldx + k] = X[srcc * ldx + k];
  if (threadIdx.x == 0 && dst >= chain0 && dst < chain0 + n_local) {
    const double v = lpost_src[srcc];
    lpost[dst - chain0] = v;
    if (hist_row) hist_row[dst - chain0] = v;
  }
}

// peer variant: every rank owns only its shard; the source row is read from the owning rank through its peer pointer
__global__ void hb_outlier_reset_peer_kernel(double* Xlocal, hb_peer_ptrs peers, long long peer_nl, int ldx, int d,
                                             double* lpost, double* hist_row, const double* lpost_src, long long chain0,
                                             long long n_local, const int* outl, const int* news, int n_out) {
  const int q = blockIdx.x;
  if (q >= n_out) return;
  const long long dst = outl[q], srcc = news[q];
  if (dst < chain0 || dst >= chain0 + n_local) return;
  const int rk = (int)(srcc / peer_nl);
  const double* srow = peers.p[rk] + (srcc - rk * peer_nl) * ldx;
  for (int k = threadIdx.x; k < d; k += blockDim.x) Xlocal[(dst - chain0) * ldx + k] = srow[k];
  if (threadIdx.x == 0) {
    const double v = lpost_src[srcc];
    lpost[dst - chain0] = v;
    if (hist_row) hist_row[dst - chain0] = v;
  }
}

// ---- misc ------------------------------------------------------------------------------------------------
__global__ void hb_init_state_kernel(const double* lp_in, const double* ll_raw, const double* fx_in, double* lp,
                                     double* ll, double* lpost, double* fx, double* hist_row, long long n,
                                     unsigned* err) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  double l = ll_raw[j];
  if (!(l >= HB_FLOOR)) l = (l != l) ? l : HB_FLOOR;
  if (!isfinite(l)) atomicOr(err, HB_FLAG_LL_NONFINITE);
  const double a = lp_in[j];
  lp[j] = a; ll[j] = l; lpost[j] = a + l;                    // R/dream_parallel.R:238
  if (fx && fx_in) fx[j] = fx_in[j];
  if (hist_row) hist_row[j] = a + l;
}

__global__ void hb_prior_kernel(const double* X, int ldx, int d, long long n, int prior, const double* pmin,
                                const double* pmax, double* lp, unsigned* err) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  double s = 0.0;
  if (prior == 1)
    for (int k = 0; k < d; ++k) {
      const double x = X[j * ldx + k], a = pmin[k], b = pmax[k];
      s += (a <= x && x <= b) ? -log(b - a) : -INFINITY;
    }
  if (!(s >= HB_FLOOR)) s = (s != s) ? s : HB_FLOOR;
  if (!isfinite(s)) atomicOr(err, HB_FLAG_LP_NONFINITE);
  lp[j] = s;
}

// mvtnorm::dmvnorm(x, mu, cov, log = TRUE) (R/internals.R:59): z = (x - mu) W with W = chol(cov)^-1 upper triangular,
// lp = -sum(log(diag(chol))) - d/2 log(2 pi) - 0.5 sum(z^2).  One warp per row; lane n owns output columns n, n+32, ...
__global__ void hb_prior_normal_kernel(const double* __restrict__ X, int ldx, int d, long long n,
                                       const double* __restrict__ W, const double* __restrict__ mu, double konst,
                                       double* __restrict__ lp, unsigned* err) {
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  const double* xr = X + row * (long long)ldx;
  double rss = 0.0;
  for (int nn = lane; nn < d; nn += 32) {
    double z = 0.0;
    for (int k = 0; k <= nn; ++k) z = fma(xr[k] - mu[k], W[k * d + nn], z);
    rss = fma(z, z, rss);
  }
  rss = hb_warp_sum(rss);
  if (lane == 0) {
    double s = konst - 0.5 * rss;
    if (!(s >= HB_FLOOR)) s = (s != s) ? s : HB_FLOOR;               // :65
    if (!isfinite(s)) atomicOr(err, HB_FLAG_LP_NONFINITE);           // :67
    lp[row] = s;
  }
}

__global__ void hb_store_row_kernel(const double* X, int ldx, int d, long long n, const double* lp, const double* ll,
                                    const double* lpost, const double* fx, double* hist_x, double* hist_l,
                                    double* hist_fx) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < n * d) { const long long j = e / d; const int k = (int)(e % d); hist_x[e] = X[j * ldx + k]; }
  if (e < n) {
    hist_l[e * 3 + 0] = lp[e]; hist_l[e * 3 + 1] = ll[e]; hist_l[e * 3 + 2] = lpost[e];
    if (hist_fx && fx) hist_fx[e] = fx[e];
  }
}

__global__ void hb_colmajor_to_rows_kernel(const double* in, long long n_rows_total, long long row0, long long n, int d,
                                           int ldx, double* out) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n * ldx) return;
  const long long j = e / ldx; const int k = (int)(e % ldx);
  out[e] = (k < d) ? in[(row0 + j) + n_rows_total * k] : 0.0;
}

__global__ void hb_rows_to_colmajor_kernel(const double* in, int ldx, int d, long long n, double* out) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n * d) return;
  const long long j = e % n; const int k = (int)(e / n);
  out[e] = in[j * ldx + k];
}

__global__ void hb_fill_kernel(double* p, long long n, double v) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < n) p[e] = v;
}

// history [rows][nct][d] (+ [rows][nct][3]) -> R's chain[nrows, d+3, ncs] for chains c0..c0+ncs-1 (32x32 smem tiles)
__global__ void hb_chain_transpose_kernel(const double* __restrict__ hist_x, const double* __restrict__ hist_l,
                                          long long row0, long long nrows, long long rows_filled, long long nct, int d,
                                          long long c0, long long ncs, double* __restrict__ out) {
  __shared__ double tile[32][33];
  const int dd = d + 3;
  const long long ncols = ncs * dd;                         // output columns: (c - c0)*(d+3) + k
  const long long colb = (long long)blockIdx.x * 32, rowb = (long long)blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const long long it = rowb + r, col = colb + threadIdx.x;
    double v = NAN;                                         // rows not yet written stay NA (R: array(NA, ...))
    if (it < nrows && col < ncols && row0 + it < rows_filled) {
      const long long c = c0 + col / dd; const int k = (int)(col % dd);
      v = (k < d) ? hist_x[((row0 + it) * nct + c) * d + k] : hist_l[((row0 + it) * nct + c) * 3 + (k - d)];
    }
    tile[r][threadIdx.x] = v;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const long long col = colb + r, it = rowb + threadIdx.x;
    if (col < ncols && it < nrows) out[it + nrows * col] = tile[threadIdx.x][r];
  }
}

}  // namespace

int hb_k3_grid(long long n_local, int d, int nCR, int sm_count, int force_wide, size_t* smem_out) {
  hb_k3_params p{};
  p.n_local = n_local; p.d = d; p.ldx = d == 1 ? 1 : (d + 3) & ~3; p.nCR = nCR; p.X = nullptr; p.force_wide = force_wide;
  int nb = 0;
  if (dispatch(p, sm_count, &nb, smem_out, nullptr) != cudaSuccess) return -1;
  return nb;
}

cudaError_t hb_launch_k3(const hb_k3_params& p, int sm_count, int* nblocks_out, size_t* smem_out, cudaStream_t st) {
  return dispatch(p, sm_count, nblocks_out, smem_out, st);
}

cudaError_t hb_launch_finalize(const hb_fin_params& p, cudaStream_t st) {
  if (p.adapt && (p.phase == 0 || p.phase == 2)) {
    const int E = (p.nCR + 2) * p.d;
    hb_finalize_reduce_kernel<<<dim3((E + 31) / 32, p.reduce_slots > 0 ? p.reduce_slots : 1), 256, 0, st>>>(p);
  }
  if (p.phase != 0) hb_finalize_kernel<<<1, 256, 0, st>>>(p);
  return cudaGetLastError();
}

cudaError_t hb_launch_proxy(const double* lpost_hist, long long nc_local, long long row0, long long row1, long long ld,
                            double* proxy, cudaStream_t st) {
  const int th = nc_local >= 256 * 64 ? 256 : 64;          // small populations: more, smaller CTAs (the loop is latency bound)
  hb_proxy_kernel<<<(unsigned)((nc_local + th - 1) / th), th, 0, st>>>(lpost_hist, nc_local, row0, row1, ld, proxy);
  return cudaGetLastError();
}

cudaError_t hb_launch_outlier_reset(double* X, int ldx, int d, double* lpost, double* lpost_hist_row,
                                    const double* lpost_src, long long chain0, long long n_local, const int* outl,
                                    const int* news, int n_out, cudaStream_t st) {
  if (n_out <= 0) return cudaSuccess;
  hb_outlier_reset_kernel<<<n_out, 128, 0, st>>>(X, ldx, d, lpost, lpost_hist_row, lpost_src, chain0, n_local, outl,
                                                 news, n_out);
  return cudaGetLastError();
}

cudaError_t hb_launch_outlier_reset_peer(double* Xlocal, hb_peer_ptrs peers, long long peer_nl, int ldx, int d,
                                         double* lpost, double* lpost_hist_row, const double* lpost_src,
                                         long long chain0, long long n_local, const int* outl, const int* news,
                                         int n_out, cudaStream_t st) {
  if (n_out <= 0) return cudaSuccess;
  hb_outlier_reset_peer_kernel<<<n_out, 128, 0, st>>>(Xlocal, peers, peer_nl, ldx, d, lpost, lpost_hist_row, lpost_src,
                                                      chain0, n_local, outl, news, n_out);
  return cudaGetLastError();
}

cudaError_t hb_launch_chain_transpose(const double* hist_x, const double* hist_l, long long out_t_alloc, long long row0,
                                      long long nrows, long long rows_filled, long long nct, int d, long long c0,
                                      long long ncs, double* out, cudaStream_t st) {
  (void)out_t_alloc;
  const long long ncols = ncs * (d + 3);
  dim3 grid((unsigned)((ncols + 31) / 32), (unsigned)((nrows + 31) / 32));
  dim3 block(32, 8);
  hb_chain_transpose_kernel<<<grid, block, 0, st>>>(hist_x, hist_l, row0, nrows, rows_filled, nct, d, c0, ncs, out);
  return cudaGetLastError();
}

cudaError_t hb_launch_init_state(const double* lp_in, const double* ll_raw, const double* fx_in, double* lp, double* ll,
                                 double* lpost, double* fx, double* lpost_hist_row, long long n, unsigned* err,
                                 cudaStream_t st) {
  hb_init_state_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(lp_in, ll_raw, fx_in, lp, ll, lpost, fx,
                                                                     lpost_hist_row, n, err);
  return cudaGetLastError();
}

cudaError_t hb_launch_prior(const double* X, int ldx, int d, long long n, int prior, const double* pmin,
                            const double* pmax, double* lp, unsigned* err, cudaStream_t st) {
  hb_prior_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(X, ldx, d, n, prior, pmin, pmax, lp, err);
  return cudaGetLastError();
}

cudaError_t hb_launch_prior_normal(const double* X, int ldx, int d, long long n, const double* W, const double* mu,
                                   double konst, double* lp, unsigned* err, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  hb_prior_normal_kernel<<<(unsigned)((n * 32 + 255) / 256), 256, 0, st>>>(X, ldx, d, n, W, mu, konst, lp, err);
  return cudaGetLastError();
}

cudaError_t hb_launch_store_row(const double* X, int ldx, int d, long long n, const double* lp, const double* ll,
                                const double* lpost, const double* fx, double* hist_x, double* hist_l, double* hist_fx,
                                cudaStream_t st) {
  const long long tot = n * (d > 1 ? d : 1);
  hb_store_row_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(X, ldx, d, n, lp, ll, lpost, fx, hist_x, hist_l,
                                                                      hist_fx);
  return cudaGetLastError();
}

cudaError_t hb_launch_colmajor_to_rows(const double* in_colmajor, long long n_rows_total, long long row0, long long n,
                                       int d, int ldx, double* out, cudaStream_t st) {
  const long long tot = n * ldx;
  hb_colmajor_to_rows_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(in_colmajor, n_rows_total, row0, n, d, ldx,
                                                                             out);
  return cudaGetLastError();
}

cudaError_t hb_launch_rows_to_colmajor(const double* in, int ldx, int d, long long n, double* out_colmajor,
                                       cudaStream_t st) {
  const long long tot = n * d;
  hb_rows_to_colmajor_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(in, ldx, d, n, out_colmajor);
  return cudaGetLastError();
}

cudaError_t hb_launch_fill(double* p, long long n, double v, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  hb_fill_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(p, n, v);
  return cudaGetLastError();
}
