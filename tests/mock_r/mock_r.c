/*
 * mock_r.c -- the objects behind tests/mock_r/include/Rinternals.h plus a .Call trampoline that turns error() into a
 * return value.  TEST INFRASTRUCTURE ONLY; see the header for scope.
 */
#include <setjmp.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include "Rinternals.h"
#include "R_ext/Rdynload.h"

struct mock_sexprec {
  int type; R_xlen_t n; void* data; SEXP names; SEXP dim; R_CFinalizer_t fin; struct mock_sexprec* next;
};
static struct mock_sexprec nil_rec = { NILSXP, 0, NULL, NULL, NULL, NULL, NULL };
static struct mock_sexprec names_sym = { CHARSXP, 0, (void*)"names", NULL, NULL, NULL, NULL };
static struct mock_sexprec dim_sym = { CHARSXP, 0, (void*)"dim", NULL, NULL, NULL, NULL };
SEXP R_NilValue = &nil_rec, R_NamesSymbol = &names_sym, R_DimSymbol = &dim_sym;

static SEXP all_objects = NULL;
static void* transient[256]; static int n_transient = 0;
static int protect_depth = 0;
static jmp_buf* err_jmp = NULL;
static char err_msg[1024];
static int interrupts_checked = 0;

static SEXP new_obj(int type, R_xlen_t n, size_t elt) {
  SEXP s = (SEXP)calloc(1, sizeof *s);
  s->type = type; s->n = n; s->names = R_NilValue; s->dim = R_NilValue;
  s->data = n > 0 ? calloc((size_t)n, elt) : NULL;
  s->next = all_objects; all_objects = s;
  return s;
}
SEXP allocVector(int type, R_xlen_t n) {
  switch (type) {
    case REALSXP: return new_obj(type, n, sizeof(double));
    case INTSXP: case LGLSXP: return new_obj(type, n, sizeof(int));
    case STRSXP: case VECSXP: { SEXP s = new_obj(type, n, sizeof(SEXP)); for (R_xlen_t i = 0; i < n; ++i) ((SEXP*)s->data)[i] = R_NilValue; return s; }
    default: Rf_error("mock allocVector: type %d not supported", type);
  }
}
static SEXP with_dim(SEXP s, int nd, const int* dims) {
  SEXP d = allocVector(INTSXP, nd);
  for (int i = 0; i < nd; ++i) INTEGER(d)[i] = dims[i];
  s->dim = d;
  return s;
}
SEXP allocMatrix(int type, int nrow, int ncol) { int d[2] = { nrow, ncol }; return with_dim(allocVector(type, (R_xlen_t)nrow * ncol), 2, d); }
SEXP alloc3DArray(int type, int a, int b, int c) { int d[3] = { a, b, c }; return with_dim(allocVector(type, (R_xlen_t)a * b * c), 3, d); }
static void want(SEXP x, int type, const char* what) { if (x->type != type) Rf_error("%s() applied to an object of type %d", what, x->type); }
double* REAL(SEXP x) { want(x, REALSXP, "REAL"); return (double*)x->data; }
int* INTEGER(SEXP x) { if (x->type != INTSXP && x->type != LGLSXP) Rf_error("INTEGER() applied to an object of type %d", x->type); return (int*)x->data; }
int* LOGICAL(SEXP x) { want(x, LGLSXP, "LOGICAL"); return (int*)x->data; }
int length(SEXP x) { return (int)x->n; }
R_xlen_t XLENGTH(SEXP x) { return x->n; }
int isNull(SEXP x) { return x->type == NILSXP; }
int TYPEOF(SEXP x) { return x->type; }
static void bounds(SEXP x, R_xlen_t i) { if (i < 0 || i >= x->n) Rf_error("subscript out of bounds"); }
SEXP STRING_ELT(SEXP x, R_xlen_t i) { want(x, STRSXP, "STRING_ELT"); bounds(x, i); return ((SEXP*)x->data)[i]; }
void SET_STRING_ELT(SEXP x, R_xlen_t i, SEXP v) { want(x, STRSXP, "SET_STRING_ELT"); bounds(x, i); ((SEXP*)x->data)[i] = v; }
SEXP VECTOR_ELT(SEXP x, R_xlen_t i) { want(x, VECSXP, "VECTOR_ELT"); bounds(x, i); return ((SEXP*)x->data)[i]; }
SEXP SET_VECTOR_ELT(SEXP x, R_xlen_t i, SEXP v) { want(x, VECSXP, "SET_VECTOR_ELT"); bounds(x, i); ((SEXP*)x->data)[i] = v; return v; }
const char* CHAR(SEXP x) { want(x, CHARSXP, "CHAR"); return (const char*)x->data; }
SEXP mkChar(const char* s) { SEXP c = new_obj(CHARSXP, 0, 1); c->data = strdup(s); c->n = (R_xlen_t)strlen(s); return c; }
SEXP mkString(const char* s) { SEXP v = allocVector(STRSXP, 1); SET_STRING_ELT(v, 0, mkChar(s)); return v; }
double asReal(SEXP x) {
  if (x->n < 1) return NAN;
  if (x->type == REALSXP) return REAL(x)[0];
  if (x->type == INTSXP || x->type == LGLSXP) return (double)INTEGER(x)[0];
  return NAN;
}
int asInteger(SEXP x) { double v = asReal(x); return isnan(v) ? (int)0x80000000 : (int)v; }
SEXP ScalarReal(double v) { SEXP s = allocVector(REALSXP, 1); REAL(s)[0] = v; return s; }
SEXP ScalarInteger(int v) { SEXP s = allocVector(INTSXP, 1); INTEGER(s)[0] = v; return s; }
SEXP ScalarLogical(int v) { SEXP s = allocVector(LGLSXP, 1); LOGICAL(s)[0] = v != 0; return s; }
SEXP getAttrib(SEXP x, SEXP name) { return name == R_NamesSymbol ? x->names : name == R_DimSymbol ? x->dim : R_NilValue; }
SEXP setAttrib(SEXP x, SEXP name, SEXP val) { if (name == R_NamesSymbol) x->names = val; else if (name == R_DimSymbol) x->dim = val; return val; }
SEXP PROTECT(SEXP x) { ++protect_depth; return x; }
void UNPROTECT(int n) { protect_depth -= n; }
SEXP R_MakeExternalPtr(void* p, SEXP tag, SEXP prot) { (void)tag; (void)prot; SEXP s = new_obj(EXTPTRSXP, 0, 1); s->data = p; return s; }
void* R_ExternalPtrAddr(SEXP s) { want(s, EXTPTRSXP, "R_ExternalPtrAddr"); return s->data; }
void R_ClearExternalPtr(SEXP s) { want(s, EXTPTRSXP, "R_ClearExternalPtr"); s->data = NULL; }
void R_RegisterCFinalizerEx(SEXP s, R_CFinalizer_t fun, Rboolean onexit) { (void)onexit; s->fin = fun; }
char* R_alloc(size_t n, int size) {                 /* freed when the .Call returns, as in R */
  if (n_transient >= 256) Rf_error("mock R_alloc: too many transient blocks");
  return (char*)(transient[n_transient++] = calloc(n ? n : 1, (size_t)size));
}
void R_CheckUserInterrupt(void) { ++interrupts_checked; }
void Rf_error(const char* fmt, ...) {
  va_list ap; va_start(ap, fmt); vsnprintf(err_msg, sizeof err_msg, fmt, ap); va_end(ap);
  if (err_jmp) longjmp(*err_jmp, 1);
  fprintf(stderr, "mock R error outside .Call: %s\n", err_msg);
  abort();
}

/* ---- registration (R_init_<pkg>) ---- */
static const R_CallMethodDef* registered = NULL;
static int dynamic_symbols = 1;
int R_registerRoutines(DllInfo* info, const void* c, const R_CallMethodDef* call, const void* f, const void* e) {
  (void)info; (void)c; (void)f; (void)e; registered = call; return 1;
}
Rboolean R_useDynamicSymbols(DllInfo* info, Rboolean value) { (void)info; int old = dynamic_symbols; dynamic_symbols = value; return old; }

/* ---- what the Python side of the test calls ---- */
extern void R_init_HydroBayes(DllInfo*);
int mockR_load(void) { R_init_HydroBayes(NULL); int n = 0; while (registered && registered[n].name) ++n; return n; }
const char* mockR_entry_name(int i) { return registered[i].name; }
int mockR_entry_nargs(int i) { return registered[i].numArgs; }
int mockR_dynamic_symbols(void) { return dynamic_symbols; }
/* .Call(name, ...): looks the routine up in the registration table, checks the arity as R does, traps error() */
SEXP mockR_dotCall(const char* name, int nargs, SEXP* a, char* err, int errlen) {
  const R_CallMethodDef* e = registered;
  while (e && e->name && strcmp(e->name, name) != 0) ++e;
  if (err && errlen) err[0] = 0;
  if (!e || !e->name) { snprintf(err, (size_t)errlen, "C symbol name \"%s\" not in DLL for package \"HydroBayes\"", name); return NULL; }
  if (e->numArgs != nargs) { snprintf(err, (size_t)errlen, "Incorrect number of arguments (%d), expecting %d for '%s'", nargs, e->numArgs, name); return NULL; }
  jmp_buf jb; jmp_buf* outer = err_jmp; SEXP res = NULL;
  const int depth0 = protect_depth;
  err_jmp = &jb;
  if (setjmp(jb) == 0) {
    typedef SEXP (*f1)(SEXP); typedef SEXP (*f2)(SEXP, SEXP); typedef SEXP (*f3)(SEXP, SEXP, SEXP); typedef SEXP (*f4)(SEXP, SEXP, SEXP, SEXP);
    switch (nargs) {
      case 1: res = ((f1)e->fun)(a[0]); break;
      case 2: res = ((f2)e->fun)(a[0], a[1]); break;
      case 3: res = ((f3)e->fun)(a[0], a[1], a[2]); break;
      case 4: res = ((f4)e->fun)(a[0], a[1], a[2], a[3]); break;
      case 7: { typedef SEXP (*f7)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
                res = ((f7)e->fun)(a[0], a[1], a[2], a[3], a[4], a[5], a[6]); break; }
      default: snprintf(err_msg, sizeof err_msg, "mock .Call: %d arguments not supported", nargs); longjmp(jb, 1);
    }
    if (protect_depth != depth0) { snprintf(err, (size_t)errlen, "stack imbalance in .Call(%s): %d", name, protect_depth - depth0); res = NULL; }
  } else {
    snprintf(err, (size_t)errlen, "%s", err_msg);
    protect_depth = depth0;                         /* R unwinds the protect stack on error */
    res = NULL;
  }
  err_jmp = outer;
  while (n_transient > 0) free(transient[--n_transient]);
  return res;
}
int mockR_interrupt_checks(void) { return interrupts_checked; }
/* constructors for the driver side (what R code would build) */
SEXP mockR_real(const double* v, R_xlen_t n) { SEXP s = allocVector(REALSXP, n); if (n) memcpy(s->data, v, (size_t)n * sizeof(double)); return s; }
SEXP mockR_real_na(R_xlen_t n) { SEXP s = allocVector(REALSXP, n); for (R_xlen_t i = 0; i < n; ++i) ((double*)s->data)[i] = NAN; return s; }
SEXP mockR_int(const int* v, R_xlen_t n) { SEXP s = allocVector(INTSXP, n); if (n) memcpy(s->data, v, (size_t)n * sizeof(int)); return s; }
SEXP mockR_set_dim(SEXP s, const int* dims, int nd) { return with_dim(s, nd, dims); }
SEXP mockR_list(int n, const char** names, SEXP* vals) {
  SEXP l = allocVector(VECSXP, n), nm = allocVector(STRSXP, n);
  for (int i = 0; i < n; ++i) { SET_VECTOR_ELT(l, i, vals[i]); SET_STRING_ELT(nm, i, mkChar(names[i])); }
  l->names = nm;
  return l;
}
SEXP mockR_nil(void) { return R_NilValue; }
int mockR_type(SEXP s) { return s->type; }
void* mockR_data(SEXP s) { return s->data; }
/* end of session: run the finalizers of live external pointers (R does at exit when onexit = TRUE), free everything */
int mockR_free_all(void) {
  int finalized = 0;
  for (SEXP s = all_objects; s; s = s->next)
    if (s->type == EXTPTRSXP && s->fin && s->data) { s->fin(s); ++finalized; }
  while (all_objects) {
    SEXP s = all_objects; all_objects = s->next;
    if (s->type != EXTPTRSXP) free(s->data);
    free(s);
  }
  return finalized;
}
