/* tests/r_stub/Rinternals.h -- TEST INFRASTRUCTURE: a minimal stand-in for the part of R's C API
 * that r_package/src/icikt_shim.c uses, so that the .Call shim can be compiled and EXECUTED in an
 * image without R.  Semantics follow "Writing R Extensions" section 5 for the calls the shim
 * makes: SEXP vectors with a type and a length, a matrix is a vector with a dim attribute,
 * PROTECT/UNPROTECT keep a stack whose depth the tests inspect, error() does not return
 * (longjmp to the .Call frame, like R's), NA_real_ is the NaN with low word 1954, and routines
 * are registered through R_registerRoutines exactly as the reference does
 * (src/RcppExports.cpp:113-128).  Nothing here is shipped; the product is r_package/ + the
 * shared library. */
#ifndef RSTUB_RINTERNALS_H
#define RSTUB_RINTERNALS_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef ptrdiff_t R_xlen_t;
typedef enum { FALSE = 0, TRUE } Rboolean;

#define NILSXP 0
#define CHARSXP 9
#define LGLSXP 10
#define INTSXP 13
#define REALSXP 14
#define STRSXP 16
#define VECSXP 19

typedef struct rstub_sexp {
  int type;
  R_xlen_t length;
  int nrow, ncol; /* dim attribute, ncol < 0: not a matrix */
  void* data;     /* double[] / int[] / struct rstub_sexp*[] / char[] */
  struct rstub_sexp* names; /* STRSXP or NULL */
} * SEXP;

extern SEXP R_NilValue;
extern double R_NaReal;
extern int R_NaInt;
#define NA_REAL R_NaReal
#define NA_INTEGER R_NaInt
#define NA_LOGICAL R_NaInt
int R_IsNA(double x);
#define ISNAN(x) ((x) != (x))

SEXP Rf_allocVector(unsigned int type, R_xlen_t n);
SEXP Rf_allocMatrix(unsigned int type, int nrow, int ncol);
SEXP Rf_protect(SEXP s);
void Rf_unprotect(int n);
SEXP Rf_mkNamed(unsigned int type, const char** names);
SEXP Rf_mkString(const char* s);
SEXP Rf_mkChar(const char* s);
SEXP Rf_ScalarLogical(int v);
SEXP Rf_ScalarInteger(int v);
SEXP Rf_ScalarReal(double v);
int Rf_asLogical(SEXP s);
int Rf_asInteger(SEXP s);
int Rf_isReal(SEXP s);
int Rf_isInteger(SEXP s);
int Rf_isMatrix(SEXP s);
int Rf_isNull(SEXP s);
int Rf_nrows(SEXP s);
int Rf_ncols(SEXP s);
R_xlen_t Rf_xlength(SEXP s);
double* REAL(SEXP s);
int* INTEGER(SEXP s);
int* LOGICAL(SEXP s);
SEXP STRING_ELT(SEXP s, R_xlen_t i);
const char* CHAR(SEXP s);
SEXP VECTOR_ELT(SEXP s, R_xlen_t i);
SEXP SET_VECTOR_ELT(SEXP s, R_xlen_t i, SEXP v);
char* R_alloc(size_t n, int size);
void Rf_error(const char* fmt, ...) __attribute__((noreturn, format(printf, 1, 2)));
void Rf_warning(const char* fmt, ...) __attribute__((format(printf, 1, 2)));

#define allocVector Rf_allocVector
#define allocMatrix Rf_allocMatrix
#define PROTECT(s) Rf_protect(s)
#define UNPROTECT(n) Rf_unprotect(n)
#define mkNamed Rf_mkNamed
#define mkString Rf_mkString
#define mkChar Rf_mkChar
#define ScalarLogical Rf_ScalarLogical
#define ScalarInteger Rf_ScalarInteger
#define ScalarReal Rf_ScalarReal
#define asLogical Rf_asLogical
#define asInteger Rf_asInteger
#define isReal Rf_isReal
#define isInteger Rf_isInteger
#define isMatrix Rf_isMatrix
#define isNull Rf_isNull
#define nrows Rf_nrows
#define ncols Rf_ncols
#define XLENGTH(s) Rf_xlength(s)
#define LENGTH(s) ((int)Rf_xlength(s))
#define error Rf_error
#define warning Rf_warning

#ifdef __cplusplus
}
#endif
#endif
