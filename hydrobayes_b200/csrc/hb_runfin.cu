// hb_runfin.cu -- per-run bookkeeping kernel: one CTA per independent run.
//
// Used by the sequential sweep (dream(), R/dream.R:203-295) and by batched independent runs (BASELINE C5), where every
// run keeps its own J, n_id, pCR, AR/CR rows and outlier resets.  Everything the reference does between two proposal
// rounds happens here on the device, with no host round trip:
//   * fold the integer counters of the round into n_id / n_acc (R/dream.R:241,248 ; R/dream_parallel.R:368-375)
//   * std_x = apply(xt, 2, sd) of the run's post-update population (two-pass) and J[id] += sum((dx/std_x)^2)
//     (R/dream.R:249-250 ; R/dream_parallel.R:376-378)
//   * pCR adaptation (R/dream.R:278-283 ; R/dream_parallel.R:407-411)
//   * check_outlier (R/internals.R:71-87): window means, type-7 quartiles from a bitonic sort in shared memory,
//     without-replacement replacement draw (tape or Philox), row copies
//   * AR / CR output rows (R/dream.R:266-267 ; R/dream_parallel.R:425-433)
#include "hb_kernels.h"

namespace {

__device__ double blk_sum(double v, double* red) {
  v = hb_warp_sum(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  if (threadIdx.x == 0) { double s = 0; for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += red[i]; red[32] = s; }
  __syncthreads();
  return red[32];
}

__device__ double quantile7_dev(const double* s, int n, double p) {   // stats::quantile type 7 on sorted values
  const double index = 1 + (double)(n > 1 ? n - 1 : 0) * p;
  const double lo = floor(index), hi = ceil(index);
  double qs = s[(int)lo - 1];
  const double xh = s[(int)hi - 1];
  if (index > lo && xh != qs) { const double h = index - lo; qs = (1 - h) * qs + h * xh; }
  return qs;
}

__global__ void __launch_bounds__(256) hb_runfin_kernel(const hb_runfin_params p) {
  extern __shared__ double sm[];
  __shared__ double red[33];
  __shared__ int s_nout;
  const int run = blockIdx.x;
  const int d = p.d, nCR = p.nCR, nc = (int)p.nc, ldx = p.ldx;
  hb_ctrl* c = p.ctrl + run;
  double* var = sm;                       // [d]
  double* srt = sm + d;                   // [npow2] sorted proxies
  double* prox = srt + p.npow2;           // [nc]
  int* ilist = reinterpret_cast<int*>(prox + nc);   // [2*nc]: outliers, then candidates / replacements
  double* Xr = p.X + (long long)run * nc * ldx;
  const long long s0 = (long long)run * nc - p.state0;   // state index of chain 0 of the run

  // ---- counters ----
  if (threadIdx.x == 0) {
    for (int q = 0; q < nCR; ++q) {
      c->n_id[q] += (double)c->n_id_gen[q]; c->n_id_gen[q] = 0ull;
      c->n_acc_id[q] += (double)c->n_acc_id_gen[q]; c->n_acc_id_gen[q] = 0ull;
    }
    c->n_acc += (double)c->n_acc_gen; c->n_acc_gen = 0ull;
    c->n_eval += (double)p.n_eval_inc;
  }
  __syncthreads();

  // ---- J: std_x of the post-update population, then the jump distances of this round's items ----
  if (p.adapt) {
    for (int k = threadIdx.x; k < d; k += blockDim.x) {
      double s = 0.0;
      for (int j = 0; j < nc; ++j) s += Xr[(long long)j * ldx + k];
      double mu = s / nc, t = 0.0;
      for (int j = 0; j < nc; ++j) t += Xr[(long long)j * ldx + k] - mu;
      mu += t / nc;
      double q = 0.0;
      for (int j = 0; j < nc; ++j) { const double e = Xr[(long long)j * ldx + k] - mu; q += e * e; }
      var[k] = q / (nc - 1);
    }
    __syncthreads();
    // items of this run in the round: item = item0 + run*item_run_stride + i*1, i < items_per_run
    for (int i = 0; i < p.items_per_run; ++i) {
      const long long item = p.item0 + (long long)run * p.item_run_stride + i;
      double part = 0.0;
      for (int k = threadIdx.x; k < d; k += blockDim.x) {
        const double q = p.DX[item * ldx + k] / sqrt(var[k]);
        part += q * q;
      }
      const double tot = blk_sum(part, red);
      if (threadIdx.x == 0) c->J[p.id[item]] += tot;
      __syncthreads();
    }
  }

  // ---- AR/CR rows of the sequential sweep are written BEFORE the adaptation (R/dream.R:266-267 precede :275) ----
  if (p.seq_row >= 0 && threadIdx.x == 0) {
    double* ar = p.out_ar + (long long)run * p.ar_rows * p.ar_cols;
    double* cr = p.out_cr + (long long)run * p.ar_rows * p.cr_cols;
    for (int q = 0; q < nCR; ++q) {
      ar[p.seq_row + p.ar_rows * q] = c->n_acc_id[q] / c->n_id[q];
      cr[p.seq_row + p.ar_rows * q] = c->pCR[q];
    }
  }
  __syncthreads();

  // ---- pCR adaptation ----
  if (p.do_pcr && threadIdx.x == 0) {
    bool anypos = false, anynan = false;
    for (int q = 0; q < nCR; ++q) { if (c->J[q] != c->J[q]) anynan = true; if (c->J[q] > 0) anypos = true; }
    if (anynan) atomicOr(&p.ctrl[0].err, HB_FLAG_J_NAN);
    else if (anypos) {
      double s = 0.0;
      for (int q = 0; q < nCR; ++q) { double v = c->J[q] / c->n_id[q]; if (v != v) v = 1.0 / nCR; c->pCR[q] = v; s += v; }
      for (int q = 0; q < nCR; ++q) c->pCR[q] = c->pCR[q] / s;
    }
  }

  // ---- check_outlier ----
  if (p.do_outlier) {
    const long long r0 = (p.gen + 1) / 2;       // ceiling(i/2), 1-based
    const long long nrow = p.gen - r0 + 1;
    for (int j = threadIdx.x; j < p.npow2; j += blockDim.x) {
      double v = INFINITY;
      if (j < nc) {
        double s = 0.0, cc = 0.0;
        for (long long r = r0; r <= p.gen; ++r) {
          const double y = p.lpost_hist[(r - 1) * p.hist_ld + s0 + j] - cc;
          const double t = s + y; cc = (t - s) - y; s = t;
        }
        v = s / (double)nrow;
        prox[j] = v;
      }
      srt[j] = v;
    }
    __syncthreads();
    for (int k = 2; k <= p.npow2; k <<= 1)
      for (int jj = k >> 1; jj > 0; jj >>= 1) {
        for (int i = threadIdx.x; i < p.npow2; i += blockDim.x) {
          const int ixj = i ^ jj;
          if (ixj > i) {
            const bool up = ((i & k) == 0);
            const double a = srt[i], b = srt[ixj];
            if ((a > b) == up) { srt[i] = b; srt[ixj] = a; }
          }
        }
        __syncthreads();
      }
    int* outl = ilist; int* cand = ilist + nc; int* news = ilist + 2 * nc;
    if (threadIdx.x == 0) {
      const double q1 = quantile7_dev(srt, nc, 0.25), q3 = quantile7_dev(srt, nc, 0.75);   // R/internals.R:75
      const double thr = q1 - 2 * fabs(q3 - q1);                                            // :76-78
      int n_out = 0, n_rem = 0;
      for (int j = 0; j < nc; ++j) { if (prox[j] < thr) outl[n_out++] = j; else cand[n_rem++] = j; }
      if (n_out > 0) {                           // new_states <- sample((1:N)[-outliers], n_out, replace = FALSE)  :81
        if (p.tape_outlier) {
          const long long e = p.gen / p.update_interval - 1;
          for (int q = 0; q < n_out; ++q) {
            int v = p.tape_outlier[e * p.nct + (long long)run * nc + q];
            if (v < 0 || v >= nc) { atomicOr(&p.ctrl[0].err, HB_FLAG_BAD_TAPE); v = 0; }
            news[q] = v;
          }
        } else {
          const hb_phx ph(p.seed, ((uint64_t)0xFFFFFFFFu << 32) | (uint32_t)(p.run_offset + run), (uint32_t)p.gen);
          uint4 w = make_uint4(0, 0, 0, 0);
          for (int q = 0; q < n_out; ++q) {      // y = x[r]; x[r] = x[--n]  (R's sample without replacement)
            if ((q & 1) == 0) w = ph.slot((uint32_t)(q >> 1));
            const int r = (int)hb_mulhi64(hb_pair64(w, q & 1), (uint64_t)n_rem);
            news[q] = cand[r];
            cand[r] = cand[--n_rem];
          }
        }
        const unsigned long long at = atomicAdd(p.outlier_count, (unsigned long long)n_out);
        for (int q = 0; q < n_out; ++q)
          if (at + q < (unsigned long long)p.outlier_cap) {
            p.outlier_gen[at + q] = p.gen; p.outlier_chain[at + q] = (long long)run * nc + outl[q];
          }
      }
      s_nout = n_out;
    }
    __syncthreads();
    const int n_out = s_nout;
    // x[outliers,] <- x[new_states,] ; dens[nrow, outliers] <- dens[nrow, new_states]  (:82-83); sources are
    // non-outliers, hence untouched by these copies
    for (int q = 0; q < n_out; ++q) {
      for (int k = threadIdx.x; k < d; k += blockDim.x) Xr[(long long)outl[q] * ldx + k] = Xr[(long long)news[q] * ldx + k];
      if (threadIdx.x == 0) {
        const double v = p.lpost[s0 + news[q]];
        p.lpost[s0 + outl[q]] = v;
        if (p.gen <= p.hist_rows) p.lpost_hist[(p.gen - 1) * p.hist_ld + s0 + outl[q]] = v;
      }
    }
    __syncthreads();
  }

  // ---- AR/CR rows of the synchronous sweep (R/dream_parallel.R:425-433), after the adaptation ----
  if (p.do_arcr && threadIdx.x == 0) {
    double* ar = p.out_ar + (long long)run * p.ar_rows * p.ar_cols;
    double* cr = p.out_cr + (long long)run * p.ar_rows * p.cr_cols;
    c->i_ar += 1;
    const long long r = c->i_ar - 1;
    c->cum_eval += c->n_eval;
    if (r < p.ar_rows) {
      ar[r + p.ar_rows] = c->n_acc / c->n_eval;
      ar[r] = c->cum_eval;
      cr[r] = c->cum_eval;
      for (int q = 0; q < nCR; ++q) cr[r + p.ar_rows * (q + 1)] = c->pCR[q];
    }
    c->n_acc = 0; c->n_eval = 0;
  }
}

}  // namespace

size_t hb_runfin_smem(int d, long long nc, int* npow2_out) {
  int np2 = 1;
  while (np2 < nc) np2 <<= 1;
  if (npow2_out) *npow2_out = np2;
  return (size_t)(d + np2 + nc) * sizeof(double) + (size_t)3 * nc * sizeof(int) + 16;
}

cudaError_t hb_launch_runfin(const hb_runfin_params& p, long long n_runs, cudaStream_t st) {
  int np2 = 0;
  const size_t smem = hb_runfin_smem(p.d, p.nc, &np2);
  static size_t attr_set = 0;
  if (smem > 48 * 1024 && smem > attr_set) {
    cudaError_t e = cudaFuncSetAttribute(hb_runfin_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    attr_set = smem;
  }
  hb_runfin_params q = p;
  q.npow2 = np2;
  hb_runfin_kernel<<<(unsigned)n_runs, 256, smem, st>>>(q);
  return cudaGetLastError();
}
