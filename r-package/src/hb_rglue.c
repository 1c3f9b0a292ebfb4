/*
 * hb_rglue.c -- .Call glue between the HydroBayes R drivers and libhydrobayes_b200.so.
 *
 * NOT COMPILED IN THE BUILD IMAGE (Rinternals.h is absent there); shown as the reference-side binding a HydroBayes
 * maintainer would add (src/ + useDynLib(HydroBayes) in NAMESPACE, PKG_LIBS = -lhydrobayes_b200 in src/Makevars).
 * Every entry point below wraps exactly one function of include/hydrobayes_b200.h.
 */
#include <R.h>
#include <Rinternals.h>
#include <string.h>
#include "hydrobayes_b200.h"

static void hb_finalizer(SEXP ptr) {
  hb_handle* h = (hb_handle*)R_ExternalPtrAddr(ptr);
  if (h) { hb_destroy(h); R_ClearExternalPtr(ptr); }
}
static hb_handle* H(SEXP ptr) {
  hb_handle* h = (hb_handle*)R_ExternalPtrAddr(ptr);
  if (!h) error("hydrobayes_b200 handle was already destroyed");
  return h;
}
static void chk(hb_handle* h, int rc) { if (rc != HB_OK) error("%s", hb_last_error(h)); }  /* -> stop(<message>) */

/* cfg: named list built by the R driver from the formals of dream()/dream_parallel() */
static double num(SEXP l, const char* n, double dflt) {
  SEXP names = getAttrib(l, R_NamesSymbol);
  for (int i = 0; i < length(l); ++i)
    if (strcmp(CHAR(STRING_ELT(names, i)), n) == 0 && !isNull(VECTOR_ELT(l, i))) return asReal(VECTOR_ELT(l, i));
  return dflt;
}

SEXP hbR_create(SEXP cfg_list, SEXP device) {
  hb_config c;
  hb_config_default(&c);
  c.sweep = (int)num(cfg_list, "sweep", 0);
  c.nc = (int64_t)num(cfg_list, "nc", 0); c.t = (int64_t)num(cfg_list, "t", 0); c.d = (int)num(cfg_list, "d", 0);
  c.lik = (int)num(cfg_list, "lik", 0);
  c.burnin = num(cfg_list, "burnin", 0); c.adapt = num(cfg_list, "adapt", 0.1);
  c.update_interval = (int)num(cfg_list, "updateInterval", 10); c.delta = (int)num(cfg_list, "delta", 3);
  c.c_val = num(cfg_list, "c_val", 0.1); c.c_star = num(cfg_list, "c_star", 1e-12); c.nCR = (int)num(cfg_list, "nCR", 3);
  c.p_g = num(cfg_list, "p_g", 0.2); c.beta0 = num(cfg_list, "beta0", 1); c.thin = (int)num(cfg_list, "thin", 1);
  c.outlier_check = (int)num(cfg_list, "outlier_check", 1);
  c.prior = (int)num(cfg_list, "prior", HB_PRIOR_FLAT); c.bound = (int)num(cfg_list, "bound", HB_BOUND_NONE);
  c.past_sample = (int)num(cfg_list, "past_sample", 0); c.m0 = (int64_t)num(cfg_list, "m0", 0);
  c.archive_update = (int)num(cfg_list, "archive_update", 0); c.psnooker = num(cfg_list, "psnooker", 0);
  c.mt = (int)num(cfg_list, "mt", 1); c.glue_shape = num(cfg_list, "glue_shape", 0);
  c.seed = (uint64_t)num(cfg_list, "seed", 1312); c.check_convergence = (int)num(cfg_list, "checkConvergence", 0);
  c.store_fx = (int)num(cfg_list, "store_fx", 1);
  hb_handle* h = NULL;
  if (hb_create(&c, asInteger(device), &h) != HB_OK) error("%s", hb_last_error(NULL));
  SEXP ptr = PROTECT(R_MakeExternalPtr(h, R_NilValue, R_NilValue));
  R_RegisterCFinalizerEx(ptr, hb_finalizer, TRUE);
  UNPROTECT(1);
  return ptr;
}

SEXP hbR_set_bounds(SEXP p, SEXP mn, SEXP mx) { chk(H(p), hb_set_bounds(H(p), REAL(mn), REAL(mx), length(mn))); return R_NilValue; }
SEXP hbR_set_obs(SEXP p, SEXP obs) { chk(H(p), hb_set_obs(H(p), REAL(obs), (int64_t)XLENGTH(obs))); return R_NilValue; }
SEXP hbR_set_prior_normal(SEXP p, SEXP mu, SEXP cov) {   /* par.info$mu, par.info$cov (prior = "normal", R/internals.R:58-59) */
  chk(H(p), hb_set_prior_normal(H(p), REAL(mu), REAL(cov)));
  return R_NilValue;
}
/* abc_rho=, abc_e=, lik_fun= : function NAMES, resolved in the device registry (R/internals.R:10,12,16) */
SEXP hbR_set_lik_args(SEXP p, SEXP abc_rho, SEXP abc_e, SEXP lik_fun) {
  chk(H(p), hb_set_lik_args(H(p), isNull(abc_rho) ? NULL : CHAR(STRING_ELT(abc_rho, 0)), isNull(abc_e) ? NULL : REAL(abc_e),
                            isNull(abc_e) ? 0 : (int64_t)XLENGTH(abc_e), isNull(lik_fun) ? NULL : CHAR(STRING_ELT(lik_fun, 0))));
  return R_NilValue;
}
SEXP hbR_set_target(SEXP p, SEXP fun, SEXP data) {   /* get(fun)(x, ...) -> name-keyed device plugin */
  int64_t dims[1] = { isNull(data) ? 0 : (int64_t)XLENGTH(data) };
  chk(H(p), hb_set_target(H(p), CHAR(STRING_ELT(fun, 0)), isNull(data) ? NULL : REAL(data), dims, isNull(data) ? 0 : 1));
  return R_NilValue;
}
SEXP hbR_set_initial(SEXP p, SEXP x0) { chk(H(p), hb_set_initial(H(p), REAL(x0))); return R_NilValue; }  /* R matrices are column-major */
SEXP hbR_run(SEXP p, SEXP n) {                      /* one slice; the R driver ticks txtProgressBar between slices */
  int64_t done = 0;
  chk(H(p), hb_run(H(p), (int64_t)asReal(n), &done));
  R_CheckUserInterrupt();
  return ScalarReal((double)done);
}
SEXP hbR_fetch(SEXP p) {                            /* list(chain, AR, CR [, fx, R_stat], outlier pairs) */
  hb_handle* h = H(p);
  int64_t out_t, ar_r, ar_c, cr_r, cr_c, fx_m, rs;
  chk(h, hb_out_dims(h, &out_t, &ar_r, &ar_c, &cr_r, &cr_c, &fx_m, &rs));
  SEXP res = PROTECT(allocVector(VECSXP, 3));
  /* dims of chain come from the driver (d+3, nc): allocate with alloc3DArray(REALSXP, out_t, d+3, nc) there and pass
     REAL(chain) to hb_fetch_chain; shown inline for AR / CR */
  SEXP ar = PROTECT(allocMatrix(REALSXP, (int)ar_r, (int)ar_c)); chk(h, hb_fetch_ar(h, REAL(ar)));
  SEXP cr = PROTECT(allocMatrix(REALSXP, (int)cr_r, (int)cr_c)); chk(h, hb_fetch_cr(h, REAL(cr)));
  SET_VECTOR_ELT(res, 1, ar); SET_VECTOR_ELT(res, 2, cr);
  UNPROTECT(3);
  return res;
}
SEXP hbR_fetch_fx(SEXP p, SEXP fx) { chk(H(p), hb_fetch_fx(H(p), REAL(fx))); return fx; }            /* fx[out_t, nc, m] :203 */
SEXP hbR_fetch_rstat(SEXP p, SEXP rs) { chk(H(p), hb_fetch_rstat(H(p), REAL(rs))); return rs; }     /* R_stat[out_t-50, d] */
SEXP hbR_fetch_outliers(SEXP p) {                   /* two integer vectors: generation i and 1-based chain (outl[[i]], :418) */
  hb_handle* h = H(p);
  int64_t n = 0;
  chk(h, hb_fetch_outliers(h, &n, NULL, NULL, 0));
  int64_t* g = (int64_t*)R_alloc((size_t)(n > 0 ? n : 1), sizeof(int64_t));
  int64_t* c = (int64_t*)R_alloc((size_t)(n > 0 ? n : 1), sizeof(int64_t));
  chk(h, hb_fetch_outliers(h, &n, g, c, n));
  SEXP res = PROTECT(allocVector(VECSXP, 2));
  SEXP gi = PROTECT(allocVector(INTSXP, (R_xlen_t)n)), ci = PROTECT(allocVector(INTSXP, (R_xlen_t)n));
  for (int64_t q = 0; q < n; ++q) { INTEGER(gi)[q] = (int)g[q]; INTEGER(ci)[q] = (int)c[q] + 1; }
  SET_VECTOR_ELT(res, 0, gi); SET_VECTOR_ELT(res, 1, ci);
  UNPROTECT(3);
  return res;
}
SEXP hbR_fetch_chain(SEXP p, SEXP chain) { chk(H(p), hb_fetch_chain(H(p), REAL(chain))); return chain; }
SEXP hbR_rstat_array(SEXP x) {                      /* R_stat(x): x[t, d, n] */
  SEXP dim = getAttrib(x, R_DimSymbol);
  int d = INTEGER(dim)[1];
  SEXP out = PROTECT(allocVector(REALSXP, d));
  char err[256];
  if (hb_rstat_array(REAL(x), INTEGER(dim)[0], d, INTEGER(dim)[2], REAL(out), err, sizeof err) != HB_OK) error("%s", err);
  UNPROTECT(1);
  return out;
}
