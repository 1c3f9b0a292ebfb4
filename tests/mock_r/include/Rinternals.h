/*
 * Minimal stand-in for R's C API -- TEST INFRASTRUCTURE ONLY (R is not installed in the build image).
 * Just enough of Rinternals.h for r-package/src/hb_rglue.c to compile and run outside R: vectors (REALSXP, INTSXP,
 * LGLSXP, STRSXP, VECSXP), NULL, external pointers with finalizers, the names/dim attributes, error() as a longjmp
 * to the mock's .Call trampoline.  No garbage collector: PROTECT/UNPROTECT count, objects are freed by mockR_free_all().
 * Semantics follow "Writing R Extensions", section 5 (System and foreign language interfaces); nothing here is R code.
 */
#ifndef MOCK_RINTERNALS_H
#define MOCK_RINTERNALS_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct mock_sexprec* SEXP;
typedef ptrdiff_t R_xlen_t;
typedef int Rboolean;
#ifndef TRUE
#define TRUE 1
#define FALSE 0
#endif
enum { NILSXP = 0, CHARSXP = 9, LGLSXP = 10, INTSXP = 13, REALSXP = 14, STRSXP = 16, VECSXP = 19, EXTPTRSXP = 22 };
extern SEXP R_NilValue, R_NamesSymbol, R_DimSymbol;

SEXP allocVector(int type, R_xlen_t n);
SEXP allocMatrix(int type, int nrow, int ncol);
SEXP alloc3DArray(int type, int a, int b, int c);
double* REAL(SEXP x);
int* INTEGER(SEXP x);
int* LOGICAL(SEXP x);
int length(SEXP x);
R_xlen_t XLENGTH(SEXP x);
int isNull(SEXP x);
int TYPEOF(SEXP x);
SEXP STRING_ELT(SEXP x, R_xlen_t i);
void SET_STRING_ELT(SEXP x, R_xlen_t i, SEXP v);
SEXP VECTOR_ELT(SEXP x, R_xlen_t i);
SEXP SET_VECTOR_ELT(SEXP x, R_xlen_t i, SEXP v);
const char* CHAR(SEXP x);
SEXP mkChar(const char* s);
SEXP mkString(const char* s);
double asReal(SEXP x);
int asInteger(SEXP x);
SEXP ScalarReal(double v);
SEXP ScalarInteger(int v);
SEXP ScalarLogical(int v);
SEXP getAttrib(SEXP x, SEXP name);
SEXP setAttrib(SEXP x, SEXP name, SEXP val);
SEXP PROTECT(SEXP x);
void UNPROTECT(int n);
SEXP R_MakeExternalPtr(void* p, SEXP tag, SEXP prot);
void* R_ExternalPtrAddr(SEXP s);
void R_ClearExternalPtr(SEXP s);
typedef void (*R_CFinalizer_t)(SEXP);
void R_RegisterCFinalizerEx(SEXP s, R_CFinalizer_t fun, Rboolean onexit);
char* R_alloc(size_t n, int size);
void R_CheckUserInterrupt(void);
void Rf_error(const char* fmt, ...) __attribute__((noreturn, format(printf, 1, 2)));
#define error Rf_error
#ifdef __cplusplus
}
#endif
#endif
