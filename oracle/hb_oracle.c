/*
 * hb_oracle.c -- CPU restatement of the HydroBayes DREAM hot path (plain C99 + optional OpenMP).
 *
 * TEST INFRASTRUCTURE ONLY -- see hb_oracle.h.  PARITY UNPINNED (no reference tests exist, R absent).
 *
 * It follows the reference text line by line, including its quirks:
 * floors at log(.Machine$double.xmin), store-before-outlier-reset, stale lp/ll after a reset,
 * lambda per selected dimension, zeta scaled by beta0, post-update std_x, biased mean in R_stat.
 * Where R accumulates in long double (sum, colSums, colMeans, mean, sd -- recollection of R's
 * summary.c / array.c / cov.c) this file uses `long double` (80-bit on x86-64) too.
 *
 * Random numbers come from one of two sources, chosen by cfg->rng_mode:
 *   HB_RNG_TAPE    decisions and continuous draws recorded from a run (hb_tape)
 *   HB_RNG_PHILOX  Philox4x32-10, counter = (chain_lo, chain_hi, generation, slot), key = seed;
 *                  the slot map is the one documented in DESIGN.md "Philox slot map" and is what
 *                  the device kernels use, so a device native-RNG run can be compared draw by draw.
 */
#include "hb_oracle.h"
#include "../include/hydrobayes_b200_rng.h" /* native-RNG draw transforms: a spec shared with the device */

#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define HBO_PI 3.141592653589793238462643383279502884

static int g_threads = 1;

/* ------------------------------------------------------------------------------------------------ */
/* constants                                                                                         */
/* ------------------------------------------------------------------------------------------------ */
double hbo_log_xmin(void) { return log(DBL_MIN); } /* R/internals.R:19,65 : log(.Machine$double.xmin) */

/* ------------------------------------------------------------------------------------------------ */
/* Philox4x32-10 (Salmon et al. 2011, Random123); third-party algorithm restated from the paper     */
/* ------------------------------------------------------------------------------------------------ */
void hbo_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
  uint32_t k0 = key[0], k1 = key[1];
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

typedef struct { uint32_t key[2]; uint32_t c0, c1, gen; } phx_t;

static void phx_init(phx_t* p, uint64_t seed, uint64_t gchain, uint32_t gen) {
  p->key[0] = (uint32_t)seed; p->key[1] = (uint32_t)(seed >> 32);
  p->c0 = (uint32_t)gchain; p->c1 = (uint32_t)(gchain >> 32); p->gen = gen;
}
static void phx_slot(const phx_t* p, uint32_t slot, uint32_t w[4]) {
  uint32_t c[4] = {p->c0, p->c1, p->gen, slot};
  hbo_philox4x32_10(c, p->key, w);
}
static inline uint64_t pair64(const uint32_t w[4], int which) {
  return which ? (((uint64_t)w[3] << 32) | w[2]) : (((uint64_t)w[1] << 32) | w[0]);
}
/* uniform in the open interval (0,1) with 53 random bits, like R's runif never returns 0 or 1 */
static inline double u53(uint64_t x) { return hb_rng_u53(x); }
static inline double u32d(uint32_t x) { return hb_rng_u32(x); }
static inline uint64_t mulhi64(uint64_t a, uint64_t b) { return (uint64_t)(((unsigned __int128)a * b) >> 64); }

/* Philox slot map (must match hydrobayes_b200/csrc/hb_common.cuh and include/hydrobayes_b200_rng.h, slot map v2) */
enum { SLOT_HEAD = 0, SLOT_PARTNER0 = 1, SLOT_CHOICE = 5, SLOT_ACCEPT = 6, SLOT_DIM0 = 8 };

/* ------------------------------------------------------------------------------------------------ */
/* targets                                                                                           */
/* ------------------------------------------------------------------------------------------------ */
/* R nmath/dnorm.c dnorm4(x, mu, sigma, give_log = FALSE) -- recollection of R >= 3.3 sources */
double hbo_dnorm(double x, double mu, double sigma) {
  const double M_1_SQRT_2PI_ = 0.398942280401432677939946059934;
  double z = (x - mu) / sigma;
  if (!isfinite(z)) return 0.0;
  z = fabs(z);
  if (z >= 2 * sqrt(DBL_MAX)) return 0.0;
  if (z < 5) return M_1_SQRT_2PI_ * exp(-0.5 * z * z) / sigma;
  if (z > sqrt(-2 * 0.693147180559945309417232121458 * (DBL_MIN_EXP + 1 - DBL_MANT_DIG))) return 0.0;
  double x1 = ldexp(nearbyint(ldexp(z, 16)), -16);
  double x2 = z - x1;
  return M_1_SQRT_2PI_ / sigma * (exp(-0.5 * x1 * x1) * exp((-0.5 * x2 - x1) * x2));
}

/* R/test_mixture.R:16 : 1/6*dnorm(x,-8,1) + 5/6*dnorm(x,10,1) */
double hbo_mixture(double x) { return 1.0 / 6.0 * hbo_dnorm(x, -8, 1) + 5.0 / 6.0 * hbo_dnorm(x, 10, 1); }

/* upper Cholesky factor R (R^T R = C), row-major d x d; LAPACK dpotrf('U') semantics, plain loops */
static int chol_upper(const double* C, int d, double* R) {
  memset(R, 0, sizeof(double) * (size_t)d * d);
  for (int j = 0; j < d; ++j) {
    long double s = C[j * d + j];
    for (int k = 0; k < j; ++k) s -= (long double)R[k * d + j] * R[k * d + j];
    if (!(s > 0)) return -1;
    double rjj = sqrt((double)s);
    R[j * d + j] = rjj;
    for (int i = j + 1; i < d; ++i) {
      long double v = C[j * d + i];
      for (int k = 0; k < j; ++k) v -= (long double)R[k * d + j] * R[k * d + i];
      R[j * d + i] = (double)(v / rjj);
    }
  }
  return 0;
}

typedef struct { int d; double* R; double konst; } tdist_t;

/* R/test_t_distribution.R:27-40 : df = 60, A = 0.5 I + 0.5, C[i,j] = A[i,j]*sqrt(i*j) (1-based) */
static int tdist_setup(tdist_t* td, int d) {
  const double df = 60;
  td->d = d;
  double* C = (double*)calloc((size_t)d * d, sizeof(double));
  td->R = (double*)malloc(sizeof(double) * (size_t)d * d);
  for (int i = 1; i <= d; ++i)
    for (int j = 1; j <= d; ++j) {
      double a = (i == j) ? 1.0 : 0.5;
      C[(i - 1) * d + (j - 1)] = a * sqrt((double)i * (double)j);
    }
  int rc = chol_upper(C, d, td->R);
  free(C);
  if (rc) return rc;
  /* mvtnorm::dmvt (recollection): lgamma((p+df)/2) - (lgamma(df/2) + sum(log(diag(dec))) + p/2*log(pi*df)) */
  long double sl = 0;
  for (int i = 0; i < d; ++i) sl += logl((long double)td->R[i * d + i]);
  td->konst = lgamma((d + df) / 2) - (lgamma(df / 2) + (double)sl + d / 2.0 * log(HBO_PI * df));
  return 0;
}
static void tdist_free(tdist_t* td) { free(td->R); td->R = NULL; }

/* dmvt: z = backsolve(dec, x, transpose = TRUE) (solves R^T z = x); rss = colSums(z^2);
 * const - 0.5*(df+p)*log1p(rss/df)     (R/test_t_distribution.R:43) */
static double tdist_eval(const tdist_t* td, const double* x, double* z) {
  const double df = 60;
  int d = td->d;
  const double* R = td->R;
  for (int i = 0; i < d; ++i) {
    double s = x[i];
    for (int k = 0; k < i; ++k) s -= R[k * d + i] * z[k];
    z[i] = s / R[i * d + i];
  }
  long double rss = 0;
  for (int i = 0; i < d; ++i) rss += (long double)(z[i] * z[i]);
  return td->konst - 0.5 * (df + d) * log1p((double)rss / df);
}

int hbo_tdist_logpdf(const double* x, int64_t n, int32_t d, double* out) {
  tdist_t td;
  if (tdist_setup(&td, d)) return HB_ERR_INVALID;
  double* xi = (double*)malloc(sizeof(double) * d);
  double* z = (double*)malloc(sizeof(double) * d);
  for (int64_t i = 0; i < n; ++i) {
    for (int k = 0; k < d; ++k) xi[k] = x[i + n * k];
    out[i] = tdist_eval(&td, xi, z);
  }
  free(xi); free(z); tdist_free(&td);
  return 0;
}

/* R/test_HydroModel.R:17-36 : linear store, backward Euler, Dt = 1, S0 = 1 */
int hbo_hydromodel(const double* forcing, int64_t T, double param, double* Y) {
  if (param < 0) return HB_ERR_INVALID;                       /* :18-19 */
  for (int64_t i = 0; i < T; ++i) if (forcing[i] < 0) return HB_ERR_INVALID; /* :20-21 */
  const double Dt = 1;
  double S0 = 1;
  for (int64_t i = 0; i < T; ++i) {
    double S_Dt = (forcing[i] * Dt + S0) / (1 + param * Dt);  /* :30 */
    Y[i] = param * S_Dt;                                      /* :31 */
    S0 = S_Dt;
  }
  return 0;
}

/* R's mean(): long double sum / n plus one refinement pass (summary.c, recollection) */
static double r_mean(const double* x, int64_t n) {
  long double s = 0;
  for (int64_t i = 0; i < n; ++i) s += x[i];
  s /= n;
  if (isfinite((double)s)) {
    long double t = 0;
    for (int64_t i = 0; i < n; ++i) t += (x[i] - s);
    s += t / n;
  }
  return (double)s;
}

/* R/internals.R:3-23 */
double hbo_calc_ll(const double* res, int64_t m, int32_t lik, const double* obs, double glue_shape, int* err) {
  return hbo_calc_ll_ex(res, m, lik, obs, glue_shape, NULL, 0, err);
}

/* R/internals.R:3-23.  abc_rho = "abc_distfun" = abs(sim - obs) and lik_fun = "fun_lik" (SLS) are the two closures the
 * reference's own example script defines (examples/script_examples.R:494, 503-507); R recycles abc_e over the m distances */
double hbo_calc_ll_ex(const double* res, int64_t m, int32_t lik, const double* obs, double glue_shape,
                      const double* abc_e, int64_t n_abc_e, int* err) {
  double out;
  *err = 0;
  if (lik == 1) {
    out = log(res[0]);                                        /* :4-5 */
  } else if (lik == 2) {
    out = res[0];                                             /* :6-7 */
  } else if (lik == 31) {                                     /* :13-14 */
    long double sse = 0, sst = 0;
    double mo = r_mean(obs, m);
    for (int64_t i = 0; i < m; ++i) { double e = res[i] - obs[i]; sse += (long double)(e * e); }
    for (int64_t i = 0; i < m; ++i) { double e = obs[i] - mo; sst += (long double)(e * e); }
    double nse = 1 - (double)sse / (double)sst;
    out = glue_shape * log(nse > 0 ? nse : 0);
  } else if (lik == 21) {                                     /* :8-10 */
    if (!abc_e || n_abc_e < 1 || !obs) { *err = HB_ERR_INVALID; return NAN; }
    long double sl = 0, q = 0;
    for (int64_t i = 0; i < n_abc_e; ++i) sl += logl((long double)abc_e[i]);   /* sum(log(abc_e)): over abc_e AS GIVEN */
    for (int64_t i = 0; i < m; ++i) { double r = fabs(res[i] - obs[i]) / abc_e[i % n_abc_e]; q += (long double)(r * r); }
    out = -(double)m / 2 * log(2 * HBO_PI) - (double)sl - 0.5 * (double)q;
  } else if (lik == 22) {                                     /* :11-12  min(abc_e - rho) */
    if (!abc_e || n_abc_e < 1 || !obs) { *err = HB_ERR_INVALID; return NAN; }
    out = INFINITY;
    for (int64_t i = 0; i < m; ++i) { double v = abc_e[i % n_abc_e] - fabs(res[i] - obs[i]); if (v < out || isnan(v)) out = v; }
  } else if (lik == 99) {                                     /* :15-16  fun_lik: -n/2 * log(sum((sim-obs)^2)) */
    if (!obs) { *err = HB_ERR_INVALID; return NAN; }
    long double sse = 0;
    for (int64_t i = 0; i < m; ++i) { double e = res[i] - obs[i]; sse += (long double)(e * e); }
    out = -(double)m / 2 * log((double)sse);
  } else {
    *err = HB_ERR_UNSUPPORTED;
    return NAN;
  }
  double fl = hbo_log_xmin();
  if (!(out >= fl)) out = isnan(out) ? out : fl;              /* :19  max(out, log(xmin)); max() propagates NaN */
  if (!isfinite(out)) *err = HB_ERR_NONFINITE;                /* :21 */
  return out;
}

/* R/internals.R:89-113 ; handle: HB_BOUND_* ; caller guarantees min/max non-NULL */
double hbo_bound_par(double par, double mn, double mx, int32_t handle) {
  double o = par;
  if (handle == HB_BOUND_NONE) return o;
  if (par < mn || par > mx) {
    if (handle == HB_BOUND_BOUND) {
      o = fmax(fmin(par, mx), mn);                            /* :94 */
    } else if (handle == HB_BOUND_REFLECT) {
      while (o < mn || o > mx) {                              /* :97-100 */
        if (o < mn) o = mn + fabs(mn - o);
        if (o > mx) o = mx - fabs(mx - o);
      }
    } else if (handle == HB_BOUND_FOLD) {
      while (o < mn || o > mx) {                              /* :103-106 */
        if (o < mn) o = mx - fabs(mn - o);
        if (o > mx) o = mn + fabs(mx - o);
      }
    }
  }
  return o;
}

static int cmp_double(const void* a, const void* b) {
  double x = *(const double*)a, y = *(const double*)b;
  return (x > y) - (x < y);
}

/* stats::quantile(x, p) default type 7 (recollection): index = 1+(n-1)p; (1-h)*x[lo] + h*x[hi] */
double hbo_quantile7(const double* x, int64_t n, double p) {
  double* s = (double*)malloc(sizeof(double) * n);
  memcpy(s, x, sizeof(double) * n);
  qsort(s, n, sizeof(double), cmp_double);
  double index = 1 + (double)(n > 1 ? n - 1 : 0) * p;
  double lo = floor(index), hi = ceil(index);
  double qs = s[(int64_t)lo - 1];
  double xh = s[(int64_t)hi - 1];
  if (index > lo && xh != qs) {
    double h = index - lo;
    qs = (1 - h) * qs + h * xh;
  }
  free(s);
  return qs;
}

/* R/gelman_rubin.R:21-56 ; x [t, d, n] column-major */
int hbo_rstat(const double* x, int64_t t, int32_t d, int64_t n, double* out) {
  int64_t r0 = (t + 1) / 2; /* ceiling(t/2), 1-based first row :37 */
  double* xm = (double*)malloc(sizeof(double) * n);
  for (int k = 0; k < d; ++k) {
    /* x_mean <- 2/(t-2) * colSums(y)   :31,42 */
    for (int64_t c = 0; c < n; ++c) {
      long double s = 0;
      for (int64_t it = r0 - 1; it < t; ++it) s += x[it + t * (k + (int64_t)d * c)];
      xm[c] = 2.0 / (double)(t - 2) * (double)s;
    }
    /* W_r: sum(apply(y,1,function(z) (z - x_mean)^2))   :33 */
    long double w = 0;
    for (int64_t it = r0 - 1; it < t; ++it)
      for (int64_t c = 0; c < n; ++c) {
        double e = x[it + t * (k + (int64_t)d * c)] - xm[c];
        w += (long double)(e * e);
      }
    double W = 2.0 / ((double)n * (double)(t - 2)) * (double)w; /* :37 */
    double xmm = r_mean(xm, n);                                /* :44 */
    long double b = 0;
    for (int64_t c = 0; c < n; ++c) { double e = xm[c] - xmm; b += (long double)(e * e); }
    double B = 1.0 / (2.0 * (double)(n - 1)) * (double)b;      /* :46 */
    double sigma = (double)(t - 2) / (double)t * W + 2 * B;    /* :51 */
    out[k] = sqrt((double)(n + 1) / (double)n * sigma / W - (double)(t - 2) / ((double)n * (double)t)); /* :54 */
  }
  free(xm);
  return 0;
}

/* ------------------------------------------------------------------------------------------------ */
/* sampler state                                                                                     */
/* ------------------------------------------------------------------------------------------------ */
typedef struct {
  const hb_config* cfg;
  const hbo_target* tgt;
  tdist_t td;
  int d; int64_t nc;
  const double* pmin; const double* pmax; int n_bounds;
  int have_bounds;
  /* normal prior */
  double* prior_R; double prior_const; const double* prior_mu;
  double sst; /* lik 31: sum((obs-mean(obs))^2) */
  const hb_tape* tape;
  hb_tape* rec;
  int64_t run;
  int64_t nct;
  int n_sets, set;   /* MT-DREAM: proposal sets per generation and the set being drawn (tape row g*n_sets + set; Philox
                        substream = set + 1 in the top byte of the generation word, 0 when mt == 1) */
} ctx_t;

static inline double bmin(const ctx_t* c, int k) { return c->n_bounds > 1 ? c->pmin[k] : c->pmin[0]; }
static inline double bmax(const ctx_t* c, int k) { return c->n_bounds > 1 ? c->pmax[k] : c->pmax[0]; }

/* R/internals.R:49-69 */
static double prior_pdf(const ctx_t* c, const double* x, int* err) {
  double lp;
  *err = 0;
  if (c->cfg->lik == 22) {
    lp = 0;
  } else if (c->cfg->prior == HB_PRIOR_FLAT) {
    lp = 0;                                                   /* :55 */
  } else if (c->cfg->prior == HB_PRIOR_UNIFORM) {             /* :57 sum(dunif(x,min,max,log=T)) */
    long double s = 0;
    for (int k = 0; k < c->d; ++k) {
      double a = bmin(c, k), b = bmax(c, k);
      s += (a <= x[k] && x[k] <= b) ? -log(b - a) : -INFINITY;
    }
    lp = (double)s;
  } else { /* normal: mvtnorm::dmvnorm(log=T) (recollection): -sum(log(diag(dec))) - p/2 log(2pi) - rss/2 */
    int d = c->d;
    double* z = (double*)malloc(sizeof(double) * d);
    for (int i = 0; i < d; ++i) {
      double s = x[i] - c->prior_mu[i];
      for (int k = 0; k < i; ++k) s -= c->prior_R[k * d + i] * z[k];
      z[i] = s / c->prior_R[i * d + i];
    }
    long double rss = 0;
    for (int i = 0; i < d; ++i) rss += (long double)(z[i] * z[i]);
    free(z);
    lp = c->prior_const - 0.5 * (double)rss;
  }
  double fl = hbo_log_xmin();
  if (!(lp >= fl)) lp = isnan(lp) ? lp : fl;                  /* :65 */
  if (!isfinite(lp)) *err = HB_ERR_NONFINITE;                 /* :67 */
  return lp;
}

/* get(fun)(x, ...) followed by calc_ll: returns ll, *fx the raw scalar output when m == 1 */
static double eval_target(const ctx_t* c, const double* x, double* scratch, double* fx, int* err) {
  *err = 0;
  int lik = c->cfg->lik;
  switch (c->tgt->kind) {
    case HBO_TARGET_MIXTURE: {
      double r = hbo_mixture(x[0]);
      if (fx) *fx = r;
      return hbo_calc_ll(&r, 1, lik, NULL, 0, err);
    }
    case HBO_TARGET_TDIST: {
      double r = tdist_eval(&c->td, x, scratch);
      if (fx) *fx = r;
      return hbo_calc_ll(&r, 1, lik, NULL, 0, err);
    }
    case HBO_TARGET_HYDROMODEL: {
      int64_t T = c->tgt->T;
      if (hbo_hydromodel(c->tgt->forcing, T, x[0], scratch)) { *err = HB_ERR_INVALID; return NAN; }
      if (fx) *fx = NAN;
      if (lik == 31) { /* same arithmetic as hbo_calc_ll with the constant denominator hoisted */
        long double sse = 0;
        for (int64_t i = 0; i < T; ++i) { double e = scratch[i] - c->tgt->obs[i]; sse += (long double)(e * e); }
        double nse = 1 - (double)sse / c->sst;
        double out = c->cfg->glue_shape * log(nse > 0 ? nse : 0);
        double fl = hbo_log_xmin();
        if (!(out >= fl)) out = isnan(out) ? out : fl;
        if (!isfinite(out)) *err = HB_ERR_NONFINITE;
        return out;
      }
      return hbo_calc_ll_ex(scratch, T, lik, c->tgt->obs, c->cfg->glue_shape, c->tgt->abc_e, c->tgt->n_abc_e, err);
    }
  }
  *err = HB_ERR_UNKNOWN_TARGET;
  return NAN;
}

int hbo_logdensity(const hbo_target* tgt, int32_t lik, double glue_shape, const double* x, int64_t n, int32_t d,
                   double* ll_out, double* fx_out) {
  hb_config cfg; memset(&cfg, 0, sizeof cfg);
  cfg.lik = lik; cfg.glue_shape = glue_shape;
  ctx_t c; memset(&c, 0, sizeof c);
  c.cfg = &cfg; c.tgt = tgt; c.d = d;
  if (tgt->kind == HBO_TARGET_TDIST && tdist_setup(&c.td, d)) return HB_ERR_INVALID;
  if (tgt->kind == HBO_TARGET_HYDROMODEL && lik == 31) {
    double mo = r_mean(tgt->obs, tgt->n_obs);
    long double sst = 0;
    for (int64_t i = 0; i < tgt->n_obs; ++i) { double e = tgt->obs[i] - mo; sst += (long double)(e * e); }
    c.sst = (double)sst;
  }
  size_t ns = (size_t)(tgt->kind == HBO_TARGET_HYDROMODEL ? tgt->T : d) + 1;
  double* scratch = (double*)malloc(sizeof(double) * ns);
  double* xi = (double*)malloc(sizeof(double) * d);
  int rc = 0;
  for (int64_t i = 0; i < n; ++i) {
    for (int k = 0; k < d; ++k) xi[k] = x[i + n * k];
    int err;
    double fx;
    ll_out[i] = eval_target(&c, xi, scratch, &fx, &err);
    if (fx_out) fx_out[i] = fx;
    if (err && !rc) rc = err;
  }
  free(scratch); free(xi);
  if (tgt->kind == HBO_TARGET_TDIST) tdist_free(&c.td);
  return rc;
}

/* ------------------------------------------------------------------------------------------------ */
/* calc_prop  (R/internals.R:115-219)                                                                */
/* ------------------------------------------------------------------------------------------------ */
/* partial Fisher-Yates over 0..n-1 with a sparse swap list, like R's sample(n, k) does on its index vector:
 *   y[i] = x[r]; x[r] = x[--n]   (recollection of R's do_sample / walker-free branch) */
typedef struct { int64_t pos[8]; int64_t val[8]; int cnt; } fy_t;
static int64_t fy_get(const fy_t* f, int64_t p) {
  for (int i = f->cnt - 1; i >= 0; --i) if (f->pos[i] == p) return f->val[i];
  return p;
}
static int64_t fy_draw(fy_t* f, uint64_t rnd, int64_t* n_rem) {
  int64_t r = (int64_t)mulhi64(rnd, (uint64_t)*n_rem);
  int64_t y = fy_get(f, r);
  int64_t last = fy_get(f, *n_rem - 1);
  f->pos[f->cnt] = r; f->val[f->cnt] = last; f->cnt++;
  (*n_rem)--;
  return y;
}

typedef struct {
  int id;         /* 0-based crossover index */
  int snooker;
} prop_info_t;

/* x: current population [nc][d] row-major (this run); z: archive [m][d] row-major; xp out [d] */
static int calc_prop(const ctx_t* c, int64_t i_gen, int64_t j, const double* x, const double* z, int64_t m_arch,
                     const double* pCR, double* xp, prop_info_t* info, double* ztbuf) {
  const hb_config* cfg = c->cfg;
  const int d = c->d, nCR = cfg->nCR, delta = cfg->delta;
  const int64_t nc = c->nc;
  const int64_t gch = c->run * nc + j;       /* global chain id */
  const int64_t g = i_gen - 2;               /* tape generation index */
  const int tape_mode = (cfg->rng_mode == HB_RNG_TAPE);
  const hb_tape* tp = c->tape;
  hb_tape* rec = c->rec;
  const int n_sets = c->n_sets > 1 ? c->n_sets : 1;
  const int64_t rix = tape_mode || rec ? ((g * n_sets + c->set) * c->nct + gch) : 0; /* record index */
  phx_t ph; uint32_t w[4];
  if (!tape_mode) phx_init(&ph, cfg->seed, (uint64_t)gch, (uint32_t)i_gen | (n_sets > 1 ? (uint32_t)(c->set + 1) << 24 : 0u));

  /* :119 snooker test uniform is ALWAYS drawn */
  double u_sn; int D;
  if (tape_mode) {
    u_sn = tp->u_snooker ? tp->u_snooker[rix] : 1.0;
    D = tp->D[rix];                                           /* :123 */
  } else {
    phx_slot(&ph, SLOT_HEAD, w);
    u_sn = u53(pair64(w, 0));
    D = 1 + (int)mulhi64(pair64(w, 1), (uint64_t)delta);
  }
  int snooker = (u_sn < cfg->psnooker);
  if (D < 1 || D > delta) return HB_ERR_INVALID;

  /* :124-147 partner rows */
  int n_part = (cfg->past_sample && snooker) ? 3 : 2 * D;
  int64_t part[8];
  if (tape_mode) {
    for (int p = 0; p < n_part; ++p) part[p] = tp->partners[rix * (2 * delta) + p];
  } else {
    fy_t f; f.cnt = 0;
    int64_t n_rem = cfg->past_sample ? m_arch : nc - 1;
    for (int p = 0; p < n_part; ++p) {
      if ((p & 1) == 0) phx_slot(&ph, SLOT_PARTNER0 + (uint32_t)(p >> 1), w);
      int64_t y = fy_draw(&f, pair64(w, p & 1), &n_rem);
      if (!cfg->past_sample && y >= j) y += 1;              /* (1:nc)[-j] :142 */
      part[p] = y;
    }
  }
  const double* src = cfg->past_sample ? z : x;
  const int64_t n_src = cfg->past_sample ? m_arch : nc;
  for (int p = 0; p < n_part; ++p) if (part[p] < 0 || part[p] >= n_src) return HB_ERR_INVALID;

  /* :153 / :173 crossover index (drawn in both branches) ; :192 gamma choice */
  int id, g_is_one; double g_sn = 0;
  if (tape_mode) {
    id = tp->id[rix];
    g_is_one = tp->g_is_one ? tp->g_is_one[rix] : 0;
    g_sn = tp->g_snooker ? tp->g_snooker[rix] : 1.7;
  } else {
    phx_slot(&ph, SLOT_CHOICE, w);
    double u_id = u53(pair64(w, 0));
    double cum = 0; id = nCR - 1;
    for (int q = 0; q < nCR; ++q) { cum += pCR[q]; if (u_id < cum) { id = q; break; } }
    g_is_one = (u53(pair64(w, 1)) >= 1.0 - cfg->p_g);
    uint32_t w5[4]; phx_slot(&ph, SLOT_ACCEPT, w5);
    g_sn = 1.2 + u53(pair64(w5, 1));                        /* runif(1, 1.2, 2.2) :156 */
  }
  if (id < 0 || id >= nCR) return HB_ERR_INVALID;
  info->id = id; info->snooker = snooker;

  const double* xj = x + j * d;
  double* dx = ztbuf + d; /* scratch [d] */
  for (int k = 0; k < d; ++k) dx[k] = 0;

  if (rec) {
    if (rec->u_snooker) ((double*)rec->u_snooker)[rix] = u_sn;
    ((uint8_t*)rec->D)[rix] = (uint8_t)D;
    for (int p = 0; p < 2 * delta; ++p) ((int32_t*)rec->partners)[rix * (2 * delta) + p] = p < n_part ? (int32_t)part[p] : -1;
    ((uint8_t*)rec->id)[rix] = (uint8_t)id;
    ((uint8_t*)rec->g_is_one)[rix] = (uint8_t)g_is_one;
    if (rec->g_snooker) ((double*)rec->g_snooker)[rix] = g_sn;
  }

  if (cfg->past_sample && snooker) {
    /* :150-163 snooker update; zeta has length d, no beta0 */
    const double* r1 = z + part[0] * d; const double* r2 = z + part[1] * d; const double* r3 = z + part[2] * d;
    /* orth_proj(a = x_j, b = r1, p)  :250-263 */
    int same = 1;
    for (int k = 0; k < d; ++k) if (xj[k] != r1[k]) { same = 0; break; }
    long double s2 = 0, s3 = 0, su = 0;
    if (!same) {
      for (int k = 0; k < d; ++k) {
        double u = xj[k] - r1[k];
        s2 += (long double)((r2[k] - xj[k]) * u);
        s3 += (long double)((r3[k] - xj[k]) * u);
        su += (long double)(u * u);
      }
    }
    for (int k = 0; k < d; ++k) {
      double zeta;
      if (tape_mode) zeta = tp->zeta[rix * d + k];
      else {
        if ((k & 1) == 0) phx_slot(&ph, SLOT_DIM0 + (uint32_t)(k >> 1), w);   /* one call per two dimensions */
        const uint32_t lo = w[2 * (k & 1)], hi = w[2 * (k & 1) + 1];
        zeta = cfg->c_star * hb_rng_normal32(hi);
        if (rec) { ((double*)rec->zt)[rix * d + k] = hb_rng_u16(lo); ((double*)rec->lambda)[rix * d + k] = 0; }
      }
      if (rec) ((double*)rec->zeta)[rix * d + k] = zeta;
      double rp2, rp3;
      if (same) { rp2 = r2[k]; rp3 = r3[k]; }
      else {
        double u = xj[k] - r1[k];
        rp2 = xj[k] + u * (double)s2 / (double)su;          /* q <- a + u*sum((p-a)*u)/sum(u*u)  :257 */
        rp3 = xj[k] + u * (double)s3 / (double)su;
      }
      if (!isfinite(rp2) || !isfinite(rp3)) return HB_ERR_NONFINITE; /* :260 */
      dx[k] = zeta + g_sn * (rp2 - rp3);                    /* :163 */
    }
  } else {
    /* :166-200 parallel direction update */
    const double CRv = (double)(id + 1) / (double)nCR;        /* CR <- (1:nCR)/nCR */
    double* zt = ztbuf;
    double* lam_k = ztbuf + 2 * d; double* zeta_k = ztbuf + 3 * d; /* philox: values keyed by dimension */
    for (int k = 0; k < d; ++k) {
      if (tape_mode) zt[k] = tp->zt[rix * d + k];
      else {
        if ((k & 1) == 0) phx_slot(&ph, SLOT_DIM0 + (uint32_t)(k >> 1), w);   /* one call per two dimensions */
        const uint32_t lo = w[2 * (k & 1)], hi = w[2 * (k & 1) + 1];
        zt[k] = hb_rng_u16(lo);
        lam_k[k] = hb_rng_lambda16(lo >> 16, cfg->c_val);
        zeta_k[k] = cfg->c_star * hb_rng_normal32(hi);
      }
      if (rec) ((double*)rec->zt)[rix * d + k] = zt[k];
    }
    int d_star = 0;
    for (int k = 0; k < d; ++k) if (zt[k] < CRv) d_star++;    /* :177-179 */
    int kmin = 0, use_min = 0;
    if (d_star == 0) {                                        /* :181-184 which.min: first minimum */
      for (int k = 1; k < d; ++k) if (zt[k] < zt[kmin]) kmin = k;
      d_star = 1; use_min = 1;
    }
    double gamma_d = 2.38 / sqrt(2.0 * D * d_star);           /* :190 */
    double gam = g_is_one ? 1.0 : gamma_d;                    /* :192 */
    int rank = 0;
    for (int k = 0; k < d; ++k) {
      int sel = use_min ? (k == kmin) : (zt[k] < CRv);
      if (!sel) continue;
      double lambda, zeta;
      if (tape_mode) { lambda = tp->lambda[rix * d + rank]; zeta = tp->zeta[rix * d + rank]; } /* :188,:194 */
      else { lambda = lam_k[k]; zeta = zeta_k[k]; }
      if (rec) { ((double*)rec->lambda)[rix * d + rank] = lambda; ((double*)rec->zeta)[rix * d + rank] = zeta; }
      /* jump_diff <- colSums(r1[,A]-r2[,A])   :196  (long double accumulation) */
      long double jd = 0;
      for (int p = 0; p < D; ++p) jd += (long double)(src[part[p] * d + k] - src[part[D + p] * d + k]);
      double v = zeta + (1 + lambda) * gam * (double)jd;     /* :198 */
      dx[k] = v * cfg->beta0;                                 /* :200 */
      rank++;
    }
    if (rec) for (int r = rank; r < d; ++r) { ((double*)rec->lambda)[rix * d + r] = 0; ((double*)rec->zeta)[rix * d + r] = 0; }
  }

  for (int k = 0; k < d; ++k) {
    xp[k] = xj[k] + dx[k];                                    /* :206 */
    if (!isfinite(xp[k])) return HB_ERR_NONFINITE;            /* :209 */
  }
  if (c->have_bounds)                                         /* :210-214 */
    for (int k = 0; k < d; ++k) xp[k] = hbo_bound_par(xp[k], bmin(c, k), bmax(c, k), cfg->bound);
  return 0;
}

/* R/internals.R:221-247, mt == 1, lik != 22 */
/* R/internals.R:222-225, lik == 22: modified rule of Sadegh and Vrugt (2014), deterministic -- no uniform is drawn */
static int metropolis_accept_abc(double p_xp, double p_x) { return (p_xp >= p_x) || (p_xp >= 0); }

static int metropolis_accept(double p_xp, double p_x, double u) {
  double p_acc = fmin(fmax(-100, p_xp - p_x), 0);             /* :239 */
  return p_acc > log(u);                                      /* :241 */
}

static double accept_uniform(const ctx_t* c, int64_t i_gen, int64_t j) {
  const int64_t gch = c->run * c->nc + j;
  double u;
  if (c->cfg->rng_mode == HB_RNG_TAPE) u = c->tape->u_accept[(i_gen - 2) * c->nct + gch];
  else {
    phx_t ph; uint32_t w[4];
    phx_init(&ph, c->cfg->seed, (uint64_t)gch, (uint32_t)i_gen);
    phx_slot(&ph, SLOT_ACCEPT, w);
    u = u53(pair64(w, 0));
  }
  if (c->rec) ((double*)c->rec->u_accept)[(i_gen - 2) * c->nct + gch] = u;
  return u;
}

/* apply(xt, 2, sd): var via cov.c two-pass with long double and mean refinement (recollection) */
static void col_sd(const double* x, int64_t nc, int d, double* sd) {
  for (int k = 0; k < d; ++k) {
    long double s = 0;
    for (int64_t j = 0; j < nc; ++j) s += x[j * d + k];
    long double mu = s / nc;
    long double t = 0;
    for (int64_t j = 0; j < nc; ++j) t += (x[j * d + k] - mu);
    mu += t / nc;
    double mud = (double)mu;
    long double q = 0;
    for (int64_t j = 0; j < nc; ++j) { long double e = x[j * d + k] - mud; q += e * e; }
    sd[k] = sqrt((double)(q / (nc - 1)));
  }
}

/* R/internals.R:71-87 on window rows r0..r1 (inclusive, 1-based generations) of the lpost history.
 * Returns the number of outliers; outl[] ascending 0-based chains; news[] the replacement chains. */
static int64_t check_outlier(const ctx_t* c, int64_t i_gen, const double* hist /* [rows][nc], row = i-1 */,
                             int64_t r0, int64_t r1, double* x, double* lpost_row, int64_t* outl, int64_t* news,
                             double* proxy, int64_t* cand) {
  const int64_t nc = c->nc; const int d = c->d;
  const int64_t nrow = r1 - r0 + 1;
  for (int64_t j = 0; j < nc; ++j) {                         /* proxy <- colMeans(dens) :73 */
    long double s = 0;
    for (int64_t r = r0; r <= r1; ++r) s += hist[(r - 1) * nc + j];
    proxy[j] = (double)(s / nrow);
  }
  double q1 = hbo_quantile7(proxy, nc, 0.25), q3 = hbo_quantile7(proxy, nc, 0.75); /* :75 */
  double iqr = fabs(q3 - q1);                                 /* :76 */
  int64_t n_out = 0;
  for (int64_t j = 0; j < nc; ++j) if (proxy[j] < q1 - 2 * iqr) outl[n_out++] = j; /* :78 */
  if (n_out == 0) return 0;
  /* new_states <- sample((1:N)[-outliers], length(outliers), replace = FALSE)  :81 */
  if (c->cfg->rng_mode == HB_RNG_TAPE) {
    int64_t e = i_gen / c->cfg->update_interval - 1;
    for (int64_t q = 0; q < n_out; ++q) news[q] = c->tape->outlier_new[e * c->nct + c->run * nc + q];
  } else {
    int64_t n_rem = 0, o = 0;
    for (int64_t j = 0; j < nc; ++j) { if (o < n_out && outl[o] == j) { o++; continue; } cand[n_rem++] = j; }
    phx_t ph; uint32_t w[4];
    phx_init(&ph, c->cfg->seed, ((uint64_t)0xFFFFFFFFu << 32) | (uint32_t)c->run, (uint32_t)i_gen);
    for (int64_t q = 0; q < n_out; ++q) {
      if ((q & 1) == 0) phx_slot(&ph, (uint32_t)(q >> 1), w);
      int64_t r = (int64_t)mulhi64(pair64(w, (int)(q & 1)), (uint64_t)n_rem);
      news[q] = cand[r];
      cand[r] = cand[--n_rem];
    }
  }
  if (c->rec && c->rec->outlier_new) {
    int64_t e = i_gen / c->cfg->update_interval - 1;
    for (int64_t q = 0; q < n_out; ++q) ((int32_t*)c->rec->outlier_new)[e * c->nct + c->run * nc + q] = (int32_t)news[q];
  }
  /* x[outliers,] <- x[new_states,] ; dens[nrow,outliers] <- dens[nrow,new_states]  :82-83
   * (R evaluates the right-hand side first: copies come from the pre-reset state) */
  double* xs = (double*)malloc(sizeof(double) * (size_t)n_out * d);
  double* ls = (double*)malloc(sizeof(double) * (size_t)n_out);
  for (int64_t q = 0; q < n_out; ++q) {
    memcpy(xs + q * d, x + news[q] * d, sizeof(double) * d);
    ls[q] = lpost_row[news[q]];
  }
  for (int64_t q = 0; q < n_out; ++q) {
    memcpy(x + outl[q] * d, xs + q * d, sizeof(double) * d);
    lpost_row[outl[q]] = ls[q];
  }
  free(xs); free(ls);
  return n_out;
}

void hbo_out_dims(const hb_config* cfg, int64_t* out_t, int64_t* ar_rows, int64_t* ar_cols, int64_t* cr_rows,
                  int64_t* cr_cols) {
  int64_t ot;
  if (cfg->burnin > 0 || cfg->thin > 1) ot = (int64_t)((1 - cfg->burnin) * (double)cfg->t / cfg->thin + 1);
  else ot = cfg->t;                                           /* R/dream_parallel.R:195-198 */
  *out_t = ot;
  if (cfg->sweep == HB_SWEEP_SYNCHRONOUS) {
    *ar_rows = cfg->t / cfg->update_interval + 1; *ar_cols = 2; /* :205 */
    *cr_rows = *ar_rows; *cr_cols = cfg->nCR + 1;               /* :204 */
  } else {
    *ar_rows = ot; *ar_cols = cfg->nCR; *cr_rows = ot; *cr_cols = cfg->nCR; /* R/dream.R:162 */
  }
}

#define FAIL(code, msg) do { snprintf(io->err, sizeof io->err, "%s", msg); rc = (code); goto done; } while (0)

static const char* nonfinite_msg(int what) {
  switch (what) {
    case 0: return "Calculated proposal is not finite!";
    case 1: return "Calculated log-prior is not finite";
    case 2: return "Calculated log-likelihood is not finite!";
    default: return "Calculate Metropolis acceptance is not finite";
  }
}

static double now_s(void) {
  struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

static double g_last_loop_seconds = 0;
double hbo_last_loop_seconds(void) { return g_last_loop_seconds; }
void hbo_set_threads(int n) { g_threads = n < 1 ? 1 : n; }

int hbo_run(const hb_config* cfg, const hbo_target* tgt, const double* par_min, const double* par_max, int n_bounds,
            const double* prior_mu, const double* prior_cov, const double* x0, const hb_tape* tape, hbo_io* io) {
  int rc = 0;
  const int d = cfg->d, nCR = cfg->nCR;
  const int64_t nc = cfg->nc, t = cfg->t;
  const int seq = (cfg->sweep == HB_SWEEP_SEQUENTIAL);
  io->err[0] = 0; io->n_outlier_pairs = 0;

  double *xt = NULL, *xp = NULL, *lp = NULL, *ll = NULL, *lpost = NULL, *lp_xp = NULL, *ll_xp = NULL, *fxv = NULL,
         *fx_xp = NULL, *hist = NULL, *z = NULL, *dxm = NULL, *stdx = NULL, *proxy = NULL, *scr = NULL, *ztb = NULL;
  int64_t *outl = NULL, *news = NULL, *cand = NULL; int* ids = NULL; uint8_t* acc = NULL;
  const int mt = cfg->mt > 1 ? cfg->mt : 1;
  double *mt_xp = NULL, *mt_lp = NULL, *mt_ll = NULL, *mt_fx = NULL, *mt_zt = NULL, *mt_lpz = NULL, *mt_llz = NULL, *mt_fxz = NULL, *mt_lpostzp = NULL;
  int* mt_id = NULL;
  ctx_t c; memset(&c, 0, sizeof c);

  /* argument checks: R/dream.R:136-137, R/dream_parallel.R:166-175 */
  if (!cfg->past_sample && nc <= 2 * (int64_t)cfg->delta) { snprintf(io->err, sizeof io->err, "Argument 'nc' must be > 'delta' * 2!"); return HB_ERR_INVALID; }
  if (cfg->mt > 1 && (cfg->sweep == HB_SWEEP_SEQUENTIAL || cfg->n_runs > 1)) { snprintf(io->err, sizeof io->err, "mt > 1 (MT-DREAM) exists in dream_parallel only (R/dream_parallel.R:295)"); return HB_ERR_UNSUPPORTED; }
  if (cfg->mt > 64) { snprintf(io->err, sizeof io->err, "mt must be <= 64"); return HB_ERR_UNSUPPORTED; }
  if (cfg->lik != 1 && cfg->lik != 2 && cfg->lik != 31 && cfg->lik != 21 && cfg->lik != 22 && cfg->lik != 99) { snprintf(io->err, sizeof io->err, "Argument 'lik' has to be one of {1,2,21,22}!"); return HB_ERR_INVALID; }
  if ((cfg->lik == 21 || cfg->lik == 22 || cfg->lik == 99) && tgt->kind != HBO_TARGET_HYDROMODEL) { snprintf(io->err, sizeof io->err, "lik 21/22/99 are restated for HydroModel only"); return HB_ERR_UNSUPPORTED; }
  if (seq && (cfg->past_sample || cfg->psnooker > 0)) { snprintf(io->err, sizeof io->err, "dream() has no archive sampling"); return HB_ERR_INVALID; }
  if (cfg->rng_mode == HB_RNG_TAPE && !tape) { snprintf(io->err, sizeof io->err, "tape mode needs a tape"); return HB_ERR_INVALID; }
  if (cfg->delta > 4 || cfg->delta < 1) { snprintf(io->err, sizeof io->err, "delta must be in 1..4"); return HB_ERR_INVALID; }

  int64_t m0 = cfg->m0 > 0 ? cfg->m0 : (cfg->past_sample ? 10 * (int64_t)d : nc);      /* :168-171 */
  int au = cfg->archive_update > 0 ? cfg->archive_update : (cfg->past_sample ? 10 : 0);  /* :172-173 */
  if (m0 < nc) { snprintf(io->err, sizeof io->err, "m0 must be >= nc"); return HB_ERR_INVALID; }

  c.cfg = cfg; c.tgt = tgt; c.d = d; c.nc = nc;
  c.pmin = par_min; c.pmax = par_max; c.n_bounds = n_bounds;
  c.have_bounds = (par_min && par_max && cfg->bound != HB_BOUND_NONE);
  c.tape = tape; c.rec = io->record; c.run = io->run;
  c.nct = (cfg->n_runs > 0 ? cfg->n_runs : 1) * nc;
  c.prior_mu = prior_mu;
  if (cfg->prior == HB_PRIOR_UNIFORM && !(par_min && par_max)) { snprintf(io->err, sizeof io->err, "uniform prior needs min and max"); return HB_ERR_INVALID; }
  if (cfg->prior == HB_PRIOR_NORMAL) {
    if (!prior_mu || !prior_cov) { snprintf(io->err, sizeof io->err, "normal prior needs mu and cov"); return HB_ERR_INVALID; }
    c.prior_R = (double*)malloc(sizeof(double) * (size_t)d * d);
    double* Cm = (double*)calloc((size_t)d * d, sizeof(double));
    for (int i = 0; i < d; ++i) for (int j = 0; j < d; ++j) Cm[i * d + j] = prior_cov[i + d * j];
    int ch = chol_upper(Cm, d, c.prior_R);
    free(Cm);
    if (ch) { snprintf(io->err, sizeof io->err, "cov is not positive definite"); free(c.prior_R); return HB_ERR_INVALID; }
    long double sl = 0;
    for (int i = 0; i < d; ++i) sl += logl((long double)c.prior_R[i * d + i]);
    c.prior_const = -(double)sl - 0.5 * d * log(2 * HBO_PI);
  }
  if (tgt->kind == HBO_TARGET_TDIST && tdist_setup(&c.td, d)) { snprintf(io->err, sizeof io->err, "chol failed"); return HB_ERR_INVALID; }
  if (tgt->kind == HBO_TARGET_HYDROMODEL && cfg->lik == 31) {
    double mo = r_mean(tgt->obs, tgt->n_obs);
    long double sst = 0;
    for (int64_t i = 0; i < tgt->n_obs; ++i) { double e = tgt->obs[i] - mo; sst += (long double)(e * e); }
    c.sst = (double)sst;
  }

  int64_t out_t, ar_rows, ar_cols, cr_rows, cr_cols;
  hbo_out_dims(cfg, &out_t, &ar_rows, &ar_cols, &cr_rows, &cr_cols);
  const int64_t H = (int64_t)floor(cfg->adapt * (double)t) + 1; /* lpost rows needed by check_outlier */
  const size_t nscr = (size_t)(tgt->kind == HBO_TARGET_HYDROMODEL ? tgt->T : d) + 1;
  const int nthr = g_threads;

  xt = (double*)malloc(sizeof(double) * nc * d); xp = (double*)malloc(sizeof(double) * nc * d);
  lp = (double*)malloc(sizeof(double) * nc); ll = (double*)malloc(sizeof(double) * nc);
  lpost = (double*)malloc(sizeof(double) * nc); lp_xp = (double*)malloc(sizeof(double) * nc);
  ll_xp = (double*)malloc(sizeof(double) * nc); fxv = (double*)malloc(sizeof(double) * nc);
  fx_xp = (double*)malloc(sizeof(double) * nc);
  hist = (double*)malloc(sizeof(double) * (size_t)H * nc);
  dxm = (double*)calloc((size_t)nc * d, sizeof(double)); stdx = (double*)malloc(sizeof(double) * d);
  proxy = (double*)malloc(sizeof(double) * nc);
  scr = (double*)malloc(sizeof(double) * nscr * nthr); ztb = (double*)malloc(sizeof(double) * 4 * d * nthr);
  outl = (int64_t*)malloc(sizeof(int64_t) * nc); news = (int64_t*)malloc(sizeof(int64_t) * nc);
  cand = (int64_t*)malloc(sizeof(int64_t) * nc); ids = (int*)malloc(sizeof(int) * nc);
  acc = (uint8_t*)malloc(nc);
  if (mt > 1) {
    mt_xp = (double*)malloc(sizeof(double) * (size_t)mt * nc * d); mt_lp = (double*)malloc(sizeof(double) * (size_t)mt * nc);
    mt_ll = (double*)malloc(sizeof(double) * (size_t)mt * nc); mt_fx = (double*)malloc(sizeof(double) * (size_t)mt * nc);
    mt_id = (int*)malloc(sizeof(int) * (size_t)mt * nc); mt_zt = (double*)malloc(sizeof(double) * (size_t)nc * d);
    mt_lpz = (double*)malloc(sizeof(double) * nc); mt_llz = (double*)malloc(sizeof(double) * nc); mt_fxz = (double*)malloc(sizeof(double) * nc);
    mt_lpostzp = (double*)malloc(sizeof(double) * (size_t)mt * nc);
    c.n_sets = 2 * mt - 1;
  }
  int64_t m_arch = m0, m_cap = m0 + (au ? nc * (t / au + 1) : 0);
  if (cfg->past_sample) z = (double*)malloc(sizeof(double) * (size_t)m_cap * d);

  double J[16] = {0}, n_id[16] = {0}, n_acc_v[16] = {0}, pCR[16];
  double n_acc = 0, n_eval = 0;
  if (nCR > 16 || nCR < 1) FAIL(HB_ERR_INVALID, "nCR must be in 1..16");
  for (int q = 0; q < nCR; ++q) pCR[q] = 1.0 / nCR;           /* R/dream_parallel.R:192 */

  /* archive and initial states: last nc rows of z (R/dream_parallel.R:211-213) */
  for (int64_t j = 0; j < nc; ++j)
    for (int k = 0; k < d; ++k) xt[j * d + k] = x0[(m0 - nc + j) + m0 * k];
  if (z) for (int64_t r = 0; r < m0; ++r) for (int k = 0; k < d; ++k) z[r * d + k] = x0[r + m0 * k];

  /* initial evaluation :215-238 */
  for (int64_t j = 0; j < nc; ++j) {
    int e1, e2;
    lp[j] = prior_pdf(&c, xt + j * d, &e1);
    if (e1) FAIL(e1, nonfinite_msg(1));
    ll[j] = eval_target(&c, xt + j * d, scr, &fxv[j], &e2);
    if (e2) FAIL(e2, e2 == HB_ERR_NONFINITE ? nonfinite_msg(2) : "target evaluation failed");
    lpost[j] = lp[j] + ll[j];
    hist[j] = lpost[j];
  }

  if (io->chain) for (int64_t q = 0; q < out_t * (d + 3) * nc; ++q) io->chain[q] = NAN;
  if (io->fx) for (int64_t q = 0; q < out_t * nc; ++q) io->fx[q] = NAN;
  if (io->AR) for (int64_t q = 0; q < ar_rows * ar_cols; ++q) io->AR[q] = NAN;
  if (io->CR) for (int64_t q = 0; q < cr_rows * cr_cols; ++q) io->CR[q] = NAN;
  if (io->rstat && out_t > 50) for (int64_t q = 0; q < (out_t - 50) * d; ++q) io->rstat[q] = NAN;

#define STORE_ROW(row)                                                                                   \
  do {                                                                                                   \
    int64_t r_ = (row);                                                                                  \
    if (r_ > out_t) FAIL(HB_ERR_INVALID, "subscript out of bounds (non-integer out_t, see dream_parallel.R:195)"); \
    if (io->chain) for (int64_t j_ = 0; j_ < nc; ++j_) {                                                 \
      for (int k_ = 0; k_ < d; ++k_) io->chain[(r_ - 1) + out_t * (k_ + (int64_t)(d + 3) * j_)] = xt[j_ * d + k_]; \
      io->chain[(r_ - 1) + out_t * (d + 0 + (int64_t)(d + 3) * j_)] = lp[j_];                           \
      io->chain[(r_ - 1) + out_t * (d + 1 + (int64_t)(d + 3) * j_)] = ll[j_];                           \
      io->chain[(r_ - 1) + out_t * (d + 2 + (int64_t)(d + 3) * j_)] = lpost[j_];                        \
    }                                                                                                    \
    if (io->fx) for (int64_t j_ = 0; j_ < nc; ++j_) io->fx[(r_ - 1) + out_t * j_] = fxv[j_];             \
  } while (0)

  STORE_ROW(1);                                               /* :241-248 */
  int64_t ind_out = 1, i_ar = 1;
  if (!seq) {
    if (io->AR) { io->AR[0] = (double)nc; io->AR[ar_rows] = NAN; } /* :245-246 */
    if (io->CR) { io->CR[0] = (double)nc; for (int q = 0; q < nCR; ++q) io->CR[cr_rows * (q + 1)] = pCR[q]; }
  } else {
    if (io->AR) for (int q = 0; q < nCR; ++q) io->AR[ar_rows * q] = 0; /* R/dream.R:193 (rep(0,3) -> rep(0,nCR)) */
    if (io->CR) for (int q = 0; q < nCR; ++q) io->CR[cr_rows * q] = pCR[q];
  }
  double cum_eval = (double)nc;

  double t_loop0 = now_s();
  for (int64_t i = 2; i <= t; ++i) {
    const int store = ((double)i > cfg->burnin * (double)t) && (i % cfg->thin == 0);
    if (store) ind_out++;                                     /* :263-264 */
    const int in_adapt = ((double)i <= cfg->adapt * (double)t);
    double* hrow = (i <= H) ? hist + (i - 1) * nc : NULL;

    if (!seq) {
      /* ---- synchronous sweep: R/dream_parallel.R:269-378 ---- */
      int perr = 0;
      int64_t nacc = 0;
      if (mt > 1) {
        /* ---- MT-DREAM, R/dream_parallel.R:269-336 + metropolis_acceptance R/internals.R:227-236 ---- */
        /* stage 1: mt proposal sets around xt, try-major (replicate(mt, lapply(1:nc, ...)) :269) */
        for (int m = 0; m < mt && !perr; ++m) {
          c.set = m;
          for (int64_t j = 0; j < nc; ++j) {
            prop_info_t inf; int e;
            double* xpm = mt_xp + ((size_t)m * nc + j) * d;
            e = calc_prop(&c, i, j, xt, z, m_arch, pCR, xpm, &inf, ztb);
            if (e) { perr = e == HB_ERR_NONFINITE ? -1 : -5; break; }
            mt_id[(size_t)m * nc + j] = inf.id;
            mt_lp[(size_t)m * nc + j] = prior_pdf(&c, xpm, &e);                 /* :275 */
            if (e) { perr = -2; break; }
            mt_ll[(size_t)m * nc + j] = eval_target(&c, xpm, scr, &mt_fx[(size_t)m * nc + j], &e);   /* :278-289 */
            if (e) { perr = e == HB_ERR_NONFINITE ? -3 : -6; break; }
          }
        }
        /* candidate selection per chain, :297-307 */
        for (int64_t j = 0; j < nc && !perr; ++j) {
          int pick;
          const int64_t gix = (i - 2) * c.nct + c.run * nc + j;
          if (cfg->rng_mode == HB_RNG_TAPE) {
            if (!tape->mt_pick) { perr = -5; break; }
            pick = tape->mt_pick[gix];
            if (pick < 0 || pick >= mt) { perr = -5; break; }
          } else {
            long double psum = 0;
            double lps[64];
            for (int m = 0; m < mt; ++m) { lps[m] = exp(mt_lp[(size_t)m * nc + j] + mt_ll[(size_t)m * nc + j]); psum += (long double)lps[m]; }
            phx_t ph; uint32_t w[4];
            phx_init(&ph, cfg->seed, (uint64_t)(c.run * nc + j), (uint32_t)i);
            phx_slot(&ph, 7, w);                                                  /* slot 7: the multi-try selection uniform */
            const double u = u53(pair64(w, 0));
            double cum = 0; pick = mt - 1;
            for (int m = 0; m < mt; ++m) {
              cum += (psum == 0) ? 1.0 / mt : lps[m] / (double)psum;              /* :301-305 */
              if (u < cum) { pick = m; break; }
            }
          }
          if (c.rec && c.rec->mt_pick) ((uint8_t*)c.rec->mt_pick)[gix] = (uint8_t)pick;
          const size_t sel = (size_t)pick * nc + j;                               /* :308 */
          for (int k = 0; k < d; ++k) mt_zt[j * d + k] = mt_xp[sel * d + k];
          mt_lpz[j] = mt_lp[sel]; mt_llz[j] = mt_ll[sel]; mt_fxz[j] = mt_fx[sel]; ids[j] = mt_id[sel];   /* :309-314 */
        }
        /* stage 2: mt - 1 proposal sets around the candidates zt (calc_prop(j, zt, ...) :316), only their density is used */
        for (int m = 0; m < mt - 1 && !perr; ++m) {
          c.set = mt + m;
          for (int64_t j = 0; j < nc; ++j) {
            prop_info_t inf; int e; double fxd;
            e = calc_prop(&c, i, j, mt_zt, z, m_arch, pCR, xp + j * d, &inf, ztb);
            if (e) { perr = e == HB_ERR_NONFINITE ? -1 : -5; break; }
            const double lpz = prior_pdf(&c, xp + j * d, &e);                     /* :321 */
            if (e) { perr = -2; break; }
            const double llz = eval_target(&c, xp + j * d, scr, &fxd, &e);        /* :323-331 */
            if (e) { perr = e == HB_ERR_NONFINITE ? -3 : -6; break; }
            mt_lpostzp[(size_t)m * nc + j] = lpz + llz;                           /* :333 */
          }
        }
        c.set = 0;
        if (perr == -1) FAIL(HB_ERR_NONFINITE, nonfinite_msg(0));
        if (perr == -2) FAIL(HB_ERR_NONFINITE, nonfinite_msg(1));
        if (perr == -3) FAIL(HB_ERR_NONFINITE, nonfinite_msg(2));
        if (perr == -5) FAIL(HB_ERR_INVALID, "invalid tape record (D, partner, id or mt_pick out of range)");
        if (perr == -6) FAIL(HB_ERR_INVALID, "target evaluation failed (HydroModel: Argument 'param' has to be >= zero!)");
        for (int64_t j = 0; j < nc; ++j) {                                        /* :340-341, R/internals.R:229-236 */
          long double p1 = 0, p2 = 0;
          for (int m = 0; m < mt; ++m) p1 += (long double)exp(mt_lp[(size_t)m * nc + j] + mt_ll[(size_t)m * nc + j]);
          for (int m = 0; m < mt - 1; ++m) p2 += (long double)exp(mt_lpostzp[(size_t)m * nc + j]);
          p2 += (long double)exp(lpost[j]);                                       /* :335  mt-th value: pdf(xt) */
          int a = 0;
          if ((double)p2 != 0) {
            const double r = (double)p1 / (double)p2;
            const double p_acc = r < 1 ? r : 1;
            const double u = accept_uniform(&c, i, j);
            a = p_acc > u;
          }
          acc[j] = (uint8_t)a;
          if (a) {                                                                /* :349-355 */
            for (int k = 0; k < d; ++k) { dxm[j * d + k] = mt_zt[j * d + k] - xt[j * d + k]; xt[j * d + k] = mt_zt[j * d + k]; }
            lp[j] = mt_lpz[j]; ll[j] = mt_llz[j]; lpost[j] = mt_lpz[j] + mt_llz[j]; fxv[j] = mt_fxz[j];
            nacc++;
          } else {
            for (int k = 0; k < d; ++k) dxm[j * d + k] = 0;
          }
          n_id[ids[j]] += 1;
        }
      } else {
#pragma omp parallel for schedule(static) num_threads(nthr) if (nthr > 1)
      for (int64_t j = 0; j < nc; ++j) {
#ifdef _OPENMP
        int tid = omp_get_thread_num();
#else
        int tid = 0;
#endif
        prop_info_t inf; int e;
        e = calc_prop(&c, i, j, xt, z, m_arch, pCR, xp + j * d, &inf, ztb + (size_t)tid * 4 * d);
        if (e) { perr = e == HB_ERR_NONFINITE ? -1 : -5; continue; }
        ids[j] = inf.id;
        lp_xp[j] = prior_pdf(&c, xp + j * d, &e);             /* :275 */
        if (e) { perr = -2; continue; }
        ll_xp[j] = eval_target(&c, xp + j * d, scr + (size_t)tid * nscr, &fx_xp[j], &e); /* :278-289 */
        if (e) { perr = e == HB_ERR_NONFINITE ? -3 : -6; continue; }
      }
      if (perr == -1) FAIL(HB_ERR_NONFINITE, nonfinite_msg(0));
      if (perr == -2) FAIL(HB_ERR_NONFINITE, nonfinite_msg(1));
      if (perr == -3) FAIL(HB_ERR_NONFINITE, nonfinite_msg(2));
      if (perr == -5) FAIL(HB_ERR_INVALID, "invalid tape record (D, partner or id out of range)");
      if (perr == -6) FAIL(HB_ERR_INVALID, "target evaluation failed (HydroModel: Argument 'param' has to be >= zero!)");
      for (int64_t j = 0; j < nc; ++j) {                      /* :343 */
        double lpx = lp_xp[j] + ll_xp[j];                     /* :291 */
        int a;
        if (cfg->lik == 22) a = metropolis_accept_abc(lpx, lpost[j]);
        else { double u = accept_uniform(&c, i, j); a = metropolis_accept(lpx, lpost[j], u); }
        acc[j] = (uint8_t)a;
        if (a) {                                              /* :356-361 */
          for (int k = 0; k < d; ++k) { dxm[j * d + k] = xp[j * d + k] - xt[j * d + k]; xt[j * d + k] = xp[j * d + k]; }
          lp[j] = lp_xp[j]; ll[j] = ll_xp[j]; lpost[j] = lpx; fxv[j] = fx_xp[j];
          nacc++;
        } else {
          for (int k = 0; k < d; ++k) dxm[j * d + k] = 0;     /* :266 dx initialised to 0 */
        }
        n_id[ids[j]] += 1;                                    /* :368-369 */
      }
      }   /* mt == 1 */
      n_acc += (double)nacc; n_eval += (double)nc;            /* :374-375 */
      if (in_adapt) {                                         /* J only feeds the pCR update inside the adaptation period */
        col_sd(xt, nc, d, stdx);                              /* :376 post-update state */
        for (int64_t j = 0; j < nc; ++j) {
          long double s = 0;
          for (int k = 0; k < d; ++k) { double q = dxm[j * d + k] / stdx[k]; s += (long double)(q * q); }
          J[ids[j]] += (double)s;                             /* :377-378 */
        }
      }
    } else {
      /* ---- sequential sweep: R/dream.R:213-252 ---- */
      for (int64_t j = 0; j < nc; ++j) {
        prop_info_t inf; int e;
        double* xpj = xp + j * d;
        e = calc_prop(&c, i, j, xt, NULL, 0, pCR, xpj, &inf, ztb);
        if (e) FAIL(e, e == HB_ERR_NONFINITE ? nonfinite_msg(0) : "invalid tape record (D, partner or id out of range)");
        double lpx = prior_pdf(&c, xpj, &e);
        if (e) FAIL(e, nonfinite_msg(1));
        double fxx;
        double llx = eval_target(&c, xpj, scr, &fxx, &e);
        if (e) FAIL(e, e == HB_ERR_NONFINITE ? nonfinite_msg(2) : "target evaluation failed");
        double lpostx = lpx + llx;
        int a;                                                /* :232 */
        if (cfg->lik == 22) a = metropolis_accept_abc(lpostx, lpost[j]);
        else { double u = accept_uniform(&c, i, j); a = metropolis_accept(lpostx, lpost[j], u); }
        acc[j] = (uint8_t)a;
        double* dxj = dxm; /* reuse first row */
        if (a) {
          for (int k = 0; k < d; ++k) { dxj[k] = xpj[k] - xt[j * d + k]; xt[j * d + k] = xpj[k]; }
          lp[j] = lpx; ll[j] = llx; lpost[j] = lpostx; fxv[j] = fxx;
          n_acc_v[inf.id] += 1;                               /* :241 */
        } else {
          for (int k = 0; k < d; ++k) dxj[k] = 0;             /* :244 */
        }
        n_id[inf.id] += 1;                                    /* :248 */
        if (in_adapt) {
          col_sd(xt, nc, d, stdx);                            /* :249 recomputed after every chain */
          long double s = 0;
          for (int k = 0; k < d; ++k) { double q = dxj[k] / stdx[k]; s += (long double)(q * q); }
          J[inf.id] += (double)s;                             /* :250 */
        }
      }
    }
    if (io->accept) memcpy(io->accept + (i - 2) * nc, acc, nc);
    if (hrow) memcpy(hrow, lpost, sizeof(double) * nc);

    /* archive append: R/dream_parallel.R:382-383 */
    if (z && au && (i % au == 0)) { memcpy(z + (size_t)m_arch * d, xt, sizeof(double) * nc * d); m_arch += nc; }

    /* store: R/dream_parallel.R:386-399, R/dream.R:255-272 (before the outlier reset) */
    if (store) {
      STORE_ROW(ind_out);
      if (seq) {
        if (io->AR) for (int q = 0; q < nCR; ++q) io->AR[(ind_out - 1) + ar_rows * q] = n_acc_v[q] / n_id[q]; /* R/dream.R:266 */
        if (io->CR) for (int q = 0; q < nCR; ++q) io->CR[(ind_out - 1) + cr_rows * q] = pCR[q];
      }
      if (cfg->check_convergence && ind_out > 50 && io->rstat && io->chain) {
        /* R_stat(chain[1:ind_out, 1:d, ]) as R/dream_test.R:344 does (the call in dream*.R is broken) */
        double* sub = (double*)malloc(sizeof(double) * (size_t)ind_out * d * nc);
        double* r = (double*)malloc(sizeof(double) * d);
        for (int64_t cc = 0; cc < nc; ++cc) for (int k = 0; k < d; ++k) for (int64_t it = 0; it < ind_out; ++it)
          sub[it + ind_out * (k + (int64_t)d * cc)] = io->chain[it + out_t * (k + (int64_t)(d + 3) * cc)];
        hbo_rstat(sub, ind_out, d, nc, r);
        for (int k = 0; k < d; ++k) io->rstat[(ind_out - 51) + (out_t - 50) * k] = r[k];
        free(sub); free(r);
      }
    }

    /* adaptation: R/dream_parallel.R:403-421, R/dream.R:275-293 */
    if (in_adapt && (i % cfg->update_interval == 0)) {
      int anypos = 0;
      for (int q = 0; q < nCR; ++q) { if (isnan(J[q])) FAIL(HB_ERR_NONFINITE, "missing value where TRUE/FALSE needed (J is NaN: zero std_x)"); if (J[q] > 0) anypos = 1; }
      if (anypos) {                                           /* :407-411 */
        long double s = 0;
        for (int q = 0; q < nCR; ++q) { pCR[q] = J[q] / n_id[q]; if (isnan(pCR[q])) pCR[q] = 1.0 / nCR; }
        for (int q = 0; q < nCR; ++q) s += pCR[q];
        for (int q = 0; q < nCR; ++q) pCR[q] = pCR[q] / (double)s;
      }
      if (!cfg->past_sample && cfg->outlier_check) {          /* :414-419 */
        int64_t r0 = (i + 1) / 2;                             /* ceiling(i/2) */
        int64_t n_out = check_outlier(&c, i, hist, r0, i, xt, lpost, outl, news, proxy, cand);
        if (hrow) memcpy(hrow, lpost, sizeof(double) * nc);   /* lpost[i,] <- check_out$p_x */
        for (int64_t q = 0; q < n_out; ++q) {
          if (io->outlier_gen && io->n_outlier_pairs < io->outlier_cap) {
            io->outlier_gen[io->n_outlier_pairs] = i; io->outlier_chain[io->n_outlier_pairs] = outl[q];
          }
          io->n_outlier_pairs++;
        }
      }
    }

    /* AR and CR: R/dream_parallel.R:425-433 */
    if (!seq && (i % cfg->update_interval == 0)) {
      i_ar++;
      cum_eval += n_eval;
      if (io->AR) { io->AR[(i_ar - 1) + ar_rows] = n_acc / n_eval; io->AR[i_ar - 1] = cum_eval; }
      if (io->CR) { io->CR[i_ar - 1] = cum_eval; for (int q = 0; q < nCR; ++q) io->CR[(i_ar - 1) + cr_rows * (q + 1)] = pCR[q]; }
      n_acc = 0; n_eval = 0;
    }
  }
  g_last_loop_seconds = now_s() - t_loop0;

  if (io->xt) for (int64_t j = 0; j < nc; ++j) for (int k = 0; k < d; ++k) io->xt[j + nc * k] = xt[j * d + k];
  if (io->lp) memcpy(io->lp, lp, sizeof(double) * nc);
  if (io->ll) memcpy(io->ll, ll, sizeof(double) * nc);
  if (io->lpost) memcpy(io->lpost, lpost, sizeof(double) * nc);
  if (io->pCR) memcpy(io->pCR, pCR, sizeof(double) * nCR);
  if (io->J) memcpy(io->J, J, sizeof(double) * nCR);
  if (io->n_id) memcpy(io->n_id, n_id, sizeof(double) * nCR);

done:
  free(xt); free(xp); free(lp); free(ll); free(lpost); free(lp_xp); free(ll_xp); free(fxv); free(fx_xp); free(hist);
  free(z); free(dxm); free(stdx); free(proxy); free(scr); free(ztb); free(outl); free(news); free(cand); free(ids);
  free(acc); free(c.prior_R);
  free(mt_xp); free(mt_lp); free(mt_ll); free(mt_fx); free(mt_zt); free(mt_lpz); free(mt_llz); free(mt_fxz); free(mt_lpostzp); free(mt_id);
  if (tgt->kind == HBO_TARGET_TDIST) tdist_free(&c.td);
  return rc;
}


/* native-RNG transforms of include/hydrobayes_b200_rng.h (slot map v2), exported for the known-answer tests */
double hbo_rng_u16(uint32_t h) { return hb_rng_u16(h); }
double hbo_rng_lambda16(uint32_t h, double c_val) { return hb_rng_lambda16(h, c_val); }
double hbo_rng_normal32(uint32_t w) { return hb_rng_normal32(w); }
uint32_t hbo_rng_zt_threshold16(double cr) { return hb_rng_zt_threshold16(cr); }
