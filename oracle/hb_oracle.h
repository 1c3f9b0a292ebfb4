/*
 * hb_oracle.h -- CPU restatement of the HydroBayes DREAM hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under hydrobayes_b200/ may include, link
 * or call this; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, and only as the checker or as
 * the reported CPU baseline.
 *
 * PARITY UNPINNED: the reference ships no tests, fixtures or golden vectors
 *  and R is not installed here, so this restatement is
 * checked against the reference by code reading, against known answers derived
 * from the formulas and against scipy for the third-party
 * mvtnorm::dmvt arithmetic.  tools/record_reference.R lets anyone with R pin it.
 *
 * Each function cites the reference file:line it follows (paths relative to
 * the reference checkout, tpilz/HydroBayes).
 */
#ifndef HB_ORACLE_H
#define HB_ORACLE_H

#include <stdint.h>
#include "../include/hydrobayes_b200.h" /* hb_config, hb_tape: shared plain-C structs only */

#ifdef __cplusplus
extern "C" {
#endif

enum { HBO_TARGET_MIXTURE = 0, HBO_TARGET_TDIST = 1, HBO_TARGET_HYDROMODEL = 2 };

typedef struct hbo_target {
  int32_t kind;
  int32_t reserved;
  const double* forcing; /* HydroModel: [T] */
  int64_t T;
  const double* obs;     /* lik in {21, 22, 31, 99}: [T] */
  int64_t n_obs;
  /* lik 21 / 22 (ABC, R/internals.R:8-12): abc_rho is the registered distance "abc_distfun" = abs(sim - obs)
   * (examples/script_examples.R:494); abc_e has n_abc_e entries (1 = recycled, or T).  lik 99: lik_fun is the registered
   * "fun_lik" = -n/2 * log(sum((sim-obs)^2)) (examples/script_examples.R:503-507). */
  const double* abc_e;
  int64_t n_abc_e;
} hbo_target;

/* caller-allocated outputs (numpy arrays in the tests); any pointer may be NULL to skip */
typedef struct hbo_io {
  double* chain;      /* [out_t, d+3, nc] column-major, unwritten rows NaN */
  double* fx;         /* [out_t, nc] (targets with m == 1) */
  double* AR;         /* sync: [floor(t/upd)+1, 2] ; seq: [out_t, nCR]   column-major */
  double* CR;         /* sync: [floor(t/upd)+1, nCR+1] ; seq: [out_t, nCR] */
  uint8_t* accept;    /* [t-1][nc] */
  int64_t* outlier_gen;   /* pairs (generation i, 0-based chain) */
  int64_t* outlier_chain;
  int64_t outlier_cap;
  int64_t n_outlier_pairs; /* out */
  double* rstat;      /* [out_t-50, d] column-major when cfg->check_convergence */
  double* xt;         /* final state [nc, d] column-major */
  double* lp;
  double* ll;
  double* lpost;
  double* pCR;        /* final pCR [nCR] */
  double* J;          /* final J [nCR] */
  double* n_id;       /* final n_id [nCR] */
  hb_tape* record;    /* philox mode: write the decisions taken into these (writable) arrays */
  int64_t run;        /* which run of a batched config this call evaluates (chain ids run*nc + j) */
  char err[256];      /* message on failure */
} hbo_io;

/* constants and scalar helpers */
double hbo_log_xmin(void);                                  /* log(.Machine$double.xmin)  R/internals.R:19,65 */
void hbo_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
double hbo_rng_u16(uint32_t h);                             /* include/hydrobayes_b200_rng.h, slot map v2 */
double hbo_rng_lambda16(uint32_t h, double c_val);
double hbo_rng_normal32(uint32_t w);
uint32_t hbo_rng_zt_threshold16(double cr);
double hbo_dnorm(double x, double mu, double sigma);        /* R nmath dnorm4 (recollection) */
double hbo_mixture(double x);                               /* R/test_mixture.R:16 */
int hbo_tdist_logpdf(const double* x_colmajor, int64_t n, int32_t d, double* out); /* R/test_t_distribution.R:19-46 */
int hbo_hydromodel(const double* forcing, int64_t T, double param, double* Y);     /* R/test_HydroModel.R:17-36 */
double hbo_calc_ll_ex(const double* res, int64_t m, int32_t lik, const double* obs, double glue_shape,
                      const double* abc_e, int64_t n_abc_e, int* err);
double hbo_calc_ll(const double* res, int64_t m, int32_t lik, const double* obs, double glue_shape,
                   int* err);                               /* R/internals.R:3-23 */
double hbo_bound_par(double par, double mn, double mx, int32_t handle); /* R/internals.R:89-113 */
double hbo_quantile7(const double* x, int64_t n, double p); /* stats::quantile type 7 (recollection) */
int hbo_rstat(const double* x, int64_t t, int32_t d, int64_t n, double* out); /* R/gelman_rubin.R:21-56 */
/* log density (after calc_ll and floor) of a batch of points, column-major x [n,d] */
int hbo_logdensity(const hbo_target* tgt, int32_t lik, double glue_shape, const double* x_colmajor, int64_t n,
                   int32_t d, double* ll_out, double* fx_out);

/* output sizes, as the reference computes them (R/dream_parallel.R:195-205, R/dream.R:152-162) */
void hbo_out_dims(const hb_config* cfg, int64_t* out_t, int64_t* ar_rows, int64_t* ar_cols, int64_t* cr_rows,
                  int64_t* cr_cols);

/* The generation loop.  cfg->sweep selects R/dream_parallel.R (synchronous) or R/dream.R (sequential).
 * x0 [n0, d] column-major, n0 = m0 (archive) if past_sample else nc.  par_min/par_max may be NULL, n_bounds 1 or d.
 * tape: required when cfg->rng_mode == HB_RNG_TAPE.  Returns 0 or an HB_ERR_* code with io->err set. */
int hbo_run(const hb_config* cfg, const hbo_target* tgt, const double* par_min, const double* par_max, int n_bounds,
            const double* prior_mu, const double* prior_cov, const double* x0, const hb_tape* tape, hbo_io* io);

/* OpenMP threads used for the per-chain loop of the synchronous sweep (bench.py's CPU baseline) and the
 * monotonic-clock seconds the last hbo_run spent in its generation loop. */
void hbo_set_threads(int n);
double hbo_last_loop_seconds(void);

#ifdef __cplusplus
}
#endif
#endif
