// hb_seq.cu -- sequential sweep (dream(), R/dream.R:203-295) and batched independent runs (BASELINE C5).
//
// Both reuse the population kernels with a strided item -> chain map (hb_k1_params::chain0/stride):
//   * batched synchronous runs: one launch of K1 / K2 / K3 over all n_runs*nc chains (run-major), then the per-run
//     bookkeeping kernel hb_runfin (J, pCR, outliers, AR/CR) -- no host round trip in the generation loop;
//   * sequential sweep: chain j+1 must see chain j's update (Gauss-Seidel), so a generation is nc rounds;
//     round j launches K1 / K2 / K3 on chain j of EVERY run (the only parallelism the sweep leaves), then hb_runfin
//     recomputes std_x for that chain's jump distance exactly as R/dream.R:249-250 does after every chain.
#include <algorithm>
#include <numeric>
#include <string>
#include <vector>

#include "hb_internal.h"

#define SCK(call)                                                                                               \
  do {                                                                                                          \
    cudaError_t e_ = (call);                                                                                    \
    if (e_ != cudaSuccess) return hb_fail(h, HB_ERR_CUDA, (std::string(#call " failed: ") + cudaGetErrorString(e_)).c_str()); \
  } while (0)

int hb_seq_supported(const hb_config* cfg, std::string& why) {
  if (cfg->past_sample || cfg->psnooker > 0) { why = "archive sampling (DREAM-ZS) is only available for a single synchronous population"; return HB_ERR_UNSUPPORTED; }
  if (cfg->nc > 2048) { why = "sequential sweep / batched runs support nc <= 2048 chains per run"; return HB_ERR_UNSUPPORTED; }
  if (cfg->n_runs * cfg->nc > 0x7fffffffLL) { why = "too many chains in one handle"; return HB_ERR_UNSUPPORTED; }
  return HB_OK;
}

int hb_seq_allocate(hb_handle* h) {
  const long long events = h->t / h->cfg.update_interval + 1;
  long long cap = events * h->nct;
  if (cap > (1LL << 24)) cap = 1LL << 24;
  if (cap < 1024) cap = 1024;
  h->outl_cap = cap;
  SCK(cudaMalloc((void**)&h->outl_count_dev, sizeof(unsigned long long)));
  SCK(cudaMemset(h->outl_count_dev, 0, sizeof(unsigned long long)));
  SCK(cudaMalloc((void**)&h->outl_gen_dev, sizeof(long long) * cap));
  SCK(cudaMalloc((void**)&h->outl_chain_dev, sizeof(long long) * cap));
  return HB_OK;
}

int hb_seq_initialise(hb_handle*) { return HB_OK; }

void hb_seq_destroy(hb_handle* h) {
  if (h->outl_count_dev) cudaFree(h->outl_count_dev);
  if (h->outl_gen_dev) cudaFree(h->outl_gen_dev);
  if (h->outl_chain_dev) cudaFree(h->outl_chain_dev);
}

int hb_seq_pull_outliers(hb_handle* h) {
  unsigned long long n = 0;
  SCK(cudaStreamSynchronize(h->stream));
  SCK(cudaMemcpy(&n, h->outl_count_dev, sizeof n, cudaMemcpyDeviceToHost));
  if ((long long)n > h->outl_cap) n = (unsigned long long)h->outl_cap;
  std::vector<long long> g((size_t)n), c((size_t)n);
  if (n) {
    SCK(cudaMemcpy(g.data(), h->outl_gen_dev, sizeof(long long) * n, cudaMemcpyDeviceToHost));
    SCK(cudaMemcpy(c.data(), h->outl_chain_dev, sizeof(long long) * n, cudaMemcpyDeviceToHost));
  }
  std::vector<size_t> ord((size_t)n);
  std::iota(ord.begin(), ord.end(), 0);
  std::sort(ord.begin(), ord.end(), [&](size_t a, size_t b) { return g[a] != g[b] ? g[a] < g[b] : c[a] < c[b]; });
  h->outlier_gen.clear(); h->outlier_chain.clear();
  for (size_t q : ord) { h->outlier_gen.push_back(g[q]); h->outlier_chain.push_back(c[q]); }
  return HB_OK;
}

static hb_runfin_params make_runfin(hb_handle* h, long long i) {
  const hb_config& c = h->cfg;
  hb_runfin_params p{};
  p.ctrl = h->ctrl; p.X = h->X; p.ldx = h->ldx; p.d = h->d; p.nCR = c.nCR; p.nc = h->nc; p.nct = h->nct; p.state0 = 0;
  p.DX = h->XP; p.id = h->id;
  p.lpost = h->lpost; p.lpost_hist = h->lpost_hist; p.hist_ld = h->n_local; p.hist_rows = h->H;
  p.gen = i; p.update_interval = c.update_interval;
  p.tape_outlier = (c.rng_mode == HB_RNG_TAPE) ? h->tape_outlier_dev : nullptr;
  p.seed = c.seed; p.run_offset = c.run_offset;
  p.out_ar = h->out_ar; p.out_cr = h->out_cr; p.ar_rows = h->ar_rows; p.ar_cols = (int)h->ar_cols; p.cr_cols = (int)h->cr_cols;
  p.outlier_count = h->outl_count_dev; p.outlier_gen = h->outl_gen_dev; p.outlier_chain = h->outl_chain_dev;
  p.outlier_cap = h->outl_cap;
  p.seq_row = -1;
  return p;
}

int hb_seq_generation(hb_handle* h, long long i) {
  const hb_config& c = h->cfg;
  const long long nr = c.n_runs, nc = h->nc, nct = h->nct;
  const int d = h->d;
  const bool store = ((double)i > c.burnin * (double)h->t) && (i % c.thin == 0);    // ind_out was advanced by the caller
  const bool in_adapt = ((double)i <= c.adapt * (double)h->t);
  const bool upd = (i % c.update_interval == 0);
  const bool tape_mode = (c.rng_mode == HB_RNG_TAPE);
  if (tape_mode && in_adapt && upd && c.outlier_check && !h->tape_outlier_dev)
    return hb_fail(h, HB_ERR_INVALID, "tape has no outlier_new array");

  if (c.sweep == HB_SWEEP_SYNCHRONOUS) {
    const hb_k1_params k1 = hb_make_k1(h, i, 0, 1, nct);
    SCK(hb_launch_k1(k1, tape_mode, h->sm_count, h->stream));
    if (int rc = hb_prior_after_k1(h, nct)) return rc;
    if (int rc = hb_target_launch(h, h->XP, nct, 0, 1)) return hb_fail(h, rc, "target launch failed");
    const hb_k3_params k3 = hb_make_k3(h, i, 0, 1, nct, store, in_adapt, in_adapt);
    SCK(hb_launch_k3(k3, h->sm_count, nullptr, nullptr, h->stream));
    h->launches += 3;
    h->pending_evals += nc;
    if (in_adapt || upd) {
      hb_runfin_params f = make_runfin(h, i);
      f.item0 = 0; f.item_run_stride = nc; f.items_per_run = (int)nc;
      f.n_eval_inc = h->pending_evals;
      f.adapt = in_adapt; f.do_pcr = in_adapt && upd; f.do_outlier = in_adapt && upd && c.outlier_check; f.do_arcr = upd;
      SCK(hb_launch_runfin(f, nr, h->stream));
      h->launches++;
      h->pending_evals = 0;
    }
  } else {
    for (long long j = 0; j < nc; ++j) {                                           // R/dream.R:213
      const hb_k1_params k1 = hb_make_k1(h, i, j, nc, nr);
      SCK(hb_launch_k1(k1, tape_mode, h->sm_count, h->stream));
      if (int rc = hb_prior_after_k1(h, nr)) return rc;
      if (int rc = hb_target_launch(h, h->XP, nr, j, nc)) return hb_fail(h, rc, "target launch failed");
      const hb_k3_params k3 = hb_make_k3(h, i, j, nc, nr, false, in_adapt, in_adapt);
      SCK(hb_launch_k3(k3, h->sm_count, nullptr, nullptr, h->stream));
      h->launches += 3;
      if (in_adapt) {                                                              // :248-250 after every chain
        hb_runfin_params f = make_runfin(h, i);
        f.item0 = 0; f.item_run_stride = 1; f.items_per_run = 1; f.adapt = 1;
        SCK(hb_launch_runfin(f, nr, h->stream));
        h->launches++;
      }
    }
    if (store) {                                                                   // :255-263
      const size_t slot = (size_t)(h->ind_out - 1);
      SCK(hb_launch_store_row(h->X, h->ldx, d, nct, h->lp, h->ll, h->lpost, c.store_fx ? h->fx : nullptr,
                              h->hist_x + slot * nct * d, h->hist_l + slot * nct * 3,
                              h->hist_fx ? h->hist_fx + slot * nct : nullptr, h->stream));
      h->launches++;
    }
    if (store || (in_adapt && upd)) {                                              // :266-267, :275-293
      hb_runfin_params f = make_runfin(h, i);
      f.items_per_run = 0; f.adapt = 0;
      f.seq_row = store ? h->ind_out - 1 : -1;
      f.do_pcr = in_adapt && upd; f.do_outlier = in_adapt && upd && c.outlier_check;
      SCK(hb_launch_runfin(f, nr, h->stream));
      h->launches++;
    }
  }
  if (c.check_convergence && store && h->ind_out > 50) {
    // out_rstat[ind_out - 50, ] <- R_stat(chain[1:ind_out, 1:d, ]) of every run (R/dream.R:270-271 with out_x := the stored
    // chain as R/dream_test.R:344); the history rows are not touched by the outlier reset, so the order against
    // the bookkeeping kernel does not matter.  rstat_dev is [row][run][d].
    SCK(hb_launch_rstat_runs(h->hist_x, h->ind_out, nct, nc, nr, d, h->rstat_xm, h->rstat_ss,
                             h->rstat_dev + (size_t)(h->ind_out - 51) * nr * d, h->stream));
    h->launches += 2;
  }
  h->gen = i;
  return HB_OK;
}
