/*
 * hb_rglue.c -- .Call glue between the HydroBayes R drivers (the files under r-package/R) and libhydrobayes_b200.so.
 *
 * The reference has no native boundary at all; this is the binding a HydroBayes maintainer adds:
 * src/hb_rglue.c + src/Makevars, useDynLib(HydroBayes, .registration = TRUE) in NAMESPACE.  Every entry point wraps
 * exactly one function of include/hydrobayes_b200.h; R owns every result array (allocated by R, garbage collected),
 * the library owns what is behind the external pointer (finalizer -> hb_destroy).
 *
 * R itself is absent from the build image (no Rinternals.h), so this file cannot be compiled against real R there;
 * tests/test_r_glue.py compiles it against a minimal stand-in for R's C API (tests/mock_r/) and drives the entry points
 * in the order r-package/R/hb_drive.R calls them -- that checks the glue's own logic, not R.
 */
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>
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

/* cfg: named list built by the R driver from the formals of dream()/dream_parallel(); NULL entries keep the default */
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
  c.rng_mode = HB_RNG_PHILOX;
  c.seed = (uint64_t)num(cfg_list, "seed", 1312); c.check_convergence = (int)num(cfg_list, "checkConvergence", 0);
  c.store_fx = (int)num(cfg_list, "store_fx", 1);
  hb_handle* h = NULL;
  if (hb_create(&c, asInteger(device), &h) != HB_OK) error("%s", hb_last_error(NULL));
  SEXP ptr = PROTECT(R_MakeExternalPtr(h, R_NilValue, R_NilValue));
  R_RegisterCFinalizerEx(ptr, hb_finalizer, TRUE);
  UNPROTECT(1);
  return ptr;
}
SEXP hbR_destroy(SEXP p) { hb_finalizer(p); return R_NilValue; }   /* on.exit() of the driver: do not wait for the GC */

SEXP hbR_set_bounds(SEXP p, SEXP mn, SEXP mx) {      /* par.info$min / $max: length 1 (recycled, R/internals.R:211-212) or d */
  if (length(mn) != length(mx)) error("par.info$min and par.info$max must have the same length");
  chk(H(p), hb_set_bounds(H(p), REAL(mn), REAL(mx), length(mn)));
  return R_NilValue;
}
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
SEXP hbR_set_target(SEXP p, SEXP fun, SEXP data) {   /* get(fun)(x, ...) -> name-keyed device plugin; data = the `...` */
  int64_t dims[1] = { isNull(data) ? 0 : (int64_t)XLENGTH(data) };
  chk(H(p), hb_set_target(H(p), CHAR(STRING_ELT(fun, 0)), isNull(data) ? NULL : REAL(data), dims, isNull(data) ? 0 : 1));
  return R_NilValue;
}
SEXP hbR_target_registered(SEXP fun) { return ScalarLogical(hb_target_registered(CHAR(STRING_ELT(fun, 0)))); }
SEXP hbR_set_initial(SEXP p, SEXP x0) {             /* R matrices are column-major; the driver has checked dim(x0) == c(n0, d) */
  if (TYPEOF(x0) != REALSXP) error("hbR_set_initial: x0 must be a double matrix");
  chk(H(p), hb_set_initial(H(p), REAL(x0)));
  return R_NilValue;
}
SEXP hbR_run(SEXP p, SEXP n) {                      /* one slice; the R driver ticks txtProgressBar between slices */
  int64_t done = 0;
  chk(H(p), hb_run(H(p), (int64_t)asReal(n), &done));
  R_CheckUserInterrupt();
  return ScalarReal((double)done);
}
SEXP hbR_out_dims(SEXP p) {                         /* c(out_t, ar_rows, ar_cols, cr_rows, cr_cols, fx_m, rstat_rows) */
  int64_t v[7];
  chk(H(p), hb_out_dims(H(p), &v[0], &v[1], &v[2], &v[3], &v[4], &v[5], &v[6]));
  SEXP out = PROTECT(allocVector(REALSXP, 7));
  for (int i = 0; i < 7; ++i) REAL(out)[i] = (double)v[i];
  UNPROTECT(1);
  return out;
}
SEXP hbR_fetch_chain(SEXP p, SEXP chain) { chk(H(p), hb_fetch_chain(H(p), REAL(chain))); return chain; }  /* [out_t, d+3, nc] :201 */
SEXP hbR_fetch_ar(SEXP p, SEXP ar) { chk(H(p), hb_fetch_ar(H(p), REAL(ar))); return ar; }              /* :205 / R/dream.R:161 */
SEXP hbR_fetch_cr(SEXP p, SEXP cr) { chk(H(p), hb_fetch_cr(H(p), REAL(cr))); return cr; }              /* :204 / R/dream.R:161 */
SEXP hbR_fetch_fx(SEXP p, SEXP fx) { chk(H(p), hb_fetch_fx(H(p), REAL(fx))); return fx; }              /* fx[out_t, nc, m] :230 */
SEXP hbR_fetch_rstat(SEXP p, SEXP rs) { chk(H(p), hb_fetch_rstat(H(p), REAL(rs))); return rs; }        /* R_stat[out_t-50, d] */
SEXP hbR_fetch_outliers(SEXP p) {                   /* list(generation i, 1-based chain): outl[[i]] <- outliers (:418) */
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
SEXP hbR_rstat_array(SEXP x) {                      /* R_stat(x): x[t, d, n]  (R/gelman_rubin.R:21) */
  SEXP dim = getAttrib(x, R_DimSymbol);
  if (isNull(dim) || length(dim) != 3) error("x must be a 3-d array [samples, parameters, chains]");
  int d = INTEGER(dim)[1];
  SEXP out = PROTECT(allocVector(REALSXP, d));
  char err[256];
  if (hb_rstat_array(REAL(x), INTEGER(dim)[0], d, INTEGER(dim)[2], REAL(out), err, sizeof err) != HB_OK) error("%s", err);
  UNPROTECT(1);
  return out;
}

SEXP hbR_hydromodel_series(SEXP forcing, SEXP param) {   /* HydroModel(forcing, param_i) for every param_i: matrix [T, n] (R/test_HydroModel.R:17-36) */
  if (TYPEOF(forcing) != REALSXP || TYPEOF(param) != REALSXP) error("forcing and param must be double vectors");
  const R_xlen_t T = XLENGTH(forcing), n = XLENGTH(param);
  SEXP out = PROTECT(allocMatrix(REALSXP, (int)T, (int)n));      /* column i = series of param_i: the library's [n][T] row-major */
  char err[256];
  if (hb_hydromodel_series(REAL(forcing), (int64_t)T, REAL(param), (int64_t)n, REAL(out), err, sizeof err) != HB_OK) error("%s", err);
  UNPROTECT(1);
  return out;
}

/* ---- log-density sessions: pdf(x) of rwm() / am() / am_sd() / de_mc() and get(fun) + calc_ll of game() on the device ---- */
static void hb_density_finalizer(SEXP ptr) {
  hb_density* s = (hb_density*)R_ExternalPtrAddr(ptr);
  if (s) { hb_density_close(s); R_ClearExternalPtr(ptr); }
}
/* .hb_density_open(fun, lik, data = the `...` of get(fun), obs, glue_shape, d, max_n) -> external pointer */
SEXP hbR_density_open(SEXP fun, SEXP lik, SEXP data, SEXP obs, SEXP glue_shape, SEXP d, SEXP max_n) {
  int64_t dims[1] = { isNull(data) ? 0 : (int64_t)XLENGTH(data) };
  hb_density* s = NULL;
  char err[256];
  if (hb_density_open(CHAR(STRING_ELT(fun, 0)), asInteger(lik), isNull(data) ? NULL : REAL(data), dims, isNull(data) ? 0 : 1,
                      isNull(obs) ? NULL : REAL(obs), isNull(obs) ? 0 : (int64_t)XLENGTH(obs),
                      isNull(glue_shape) ? 0.0 : asReal(glue_shape), asInteger(d), (int64_t)asReal(max_n), &s, err,
                      sizeof err) != HB_OK) error("%s", err);
  SEXP ptr = PROTECT(R_MakeExternalPtr(s, R_NilValue, R_NilValue));
  R_RegisterCFinalizerEx(ptr, hb_density_finalizer, TRUE);
  UNPROTECT(1);
  return ptr;
}
/* x: double matrix [n, d] (column-major, as R stores it) -> calc_ll(get(fun)(x[i, ], ...), lik, ...) for every row */
SEXP hbR_density_eval(SEXP p, SEXP x) {
  hb_density* s = (hb_density*)R_ExternalPtrAddr(p);
  if (!s) error("log-density session was already closed");
  if (TYPEOF(x) != REALSXP) error("hbR_density_eval: x must be a double matrix");
  SEXP dim = getAttrib(x, R_DimSymbol);
  if (isNull(dim) || length(dim) != 2) error("x must be a matrix [points, parameters]");
  const int n = INTEGER(dim)[0];
  if (INTEGER(dim)[1] != hb_density_dim(s)) error("x has %d columns, the session was opened for d = %d", INTEGER(dim)[1], (int)hb_density_dim(s));
  SEXP out = PROTECT(allocVector(REALSXP, n));
  char err[256];
  if (hb_density_eval(s, REAL(x), (int64_t)n, REAL(out), NULL, err, sizeof err) != HB_OK) error("%s", err);
  UNPROTECT(1);
  return out;
}
SEXP hbR_density_close(SEXP p) { hb_density_finalizer(p); return R_NilValue; }

#define HB_CALLDEF(name, n) { #name, (DL_FUNC)&name, n }
static const R_CallMethodDef hb_call_entries[] = {
  HB_CALLDEF(hbR_create, 2), HB_CALLDEF(hbR_destroy, 1), HB_CALLDEF(hbR_set_bounds, 3), HB_CALLDEF(hbR_set_obs, 2),
  HB_CALLDEF(hbR_set_prior_normal, 3), HB_CALLDEF(hbR_set_lik_args, 4), HB_CALLDEF(hbR_set_target, 3),
  HB_CALLDEF(hbR_target_registered, 1), HB_CALLDEF(hbR_set_initial, 2), HB_CALLDEF(hbR_run, 2), HB_CALLDEF(hbR_out_dims, 1),
  HB_CALLDEF(hbR_fetch_chain, 2), HB_CALLDEF(hbR_fetch_ar, 2), HB_CALLDEF(hbR_fetch_cr, 2), HB_CALLDEF(hbR_fetch_fx, 2),
  HB_CALLDEF(hbR_fetch_rstat, 2), HB_CALLDEF(hbR_fetch_outliers, 1), HB_CALLDEF(hbR_rstat_array, 1),
  HB_CALLDEF(hbR_hydromodel_series, 2), HB_CALLDEF(hbR_density_open, 7), HB_CALLDEF(hbR_density_eval, 2),
  HB_CALLDEF(hbR_density_close, 1),
  { NULL, NULL, 0 }
};
void R_init_HydroBayes(DllInfo* dll) {
  R_registerRoutines(dll, NULL, hb_call_entries, NULL, NULL);
  R_useDynamicSymbols(dll, FALSE);
}
