/* tests/r_stub/r_stub.c -- TEST INFRASTRUCTURE: the stand-in R runtime behind Rinternals.h plus a
 * small driver API (rstub_*) through which tests/test_r_shim*.py build arguments, make .Call-style
 * calls into the registered routines of r_package/src/icikt_shim.c and read the results back. */
#include <setjmp.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "R.h"
#include "Rinternals.h"
#include "R_ext/Rdynload.h"

/* ---- allocation: everything lives until rstub_reset() ------------------------------------ */
static void** g_blocks = NULL;
static size_t g_nblocks = 0, g_cap = 0;
static void* keep(void* p) {
  if (g_nblocks == g_cap) {
    g_cap = g_cap ? 2 * g_cap : 256;
    g_blocks = (void**)realloc(g_blocks, g_cap * sizeof(void*));
  }
  g_blocks[g_nblocks++] = p;
  return p;
}
static struct rstub_sexp g_nil = {NILSXP, 0, 0, -1, NULL, NULL};
SEXP R_NilValue = &g_nil;
double R_NaReal;
int R_NaInt = (int)0x80000000;

static int g_protect_depth = 0, g_protect_max = 0, g_underflow = 0;
static jmp_buf g_jmp;
static int g_jmp_armed = 0;
static char g_error[1024];
static char g_warning[1024];
static DllInfo g_dll;

__attribute__((constructor)) static void rstub_init(void) {
  /* R's NA_real_: a quiet NaN whose low word is 1954 (arithmetic.c, R_NaReal) */
  const uint64_t bits = 0x7FF00000000007A2ull;
  memcpy(&R_NaReal, &bits, sizeof bits);
}

int R_IsNA(double x) {
  uint64_t b;
  memcpy(&b, &x, sizeof b);
  return x != x && (uint32_t)(b & 0xffffffffu) == 1954u;
}

static size_t elt_size(unsigned int type) {
  switch (type) {
    case REALSXP: return sizeof(double);
    case INTSXP:
    case LGLSXP: return sizeof(int);
    case VECSXP:
    case STRSXP: return sizeof(SEXP);
    case CHARSXP: return 1;
    default: return 0;
  }
}

SEXP Rf_allocVector(unsigned int type, R_xlen_t n) {
  SEXP s = (SEXP)keep(calloc(1, sizeof(struct rstub_sexp)));
  s->type = (int)type;
  s->length = n;
  s->ncol = -1;
  const size_t bytes = elt_size(type) * (size_t)(n > 0 ? n : 0) + (type == CHARSXP ? 1 : 0);
  /* like R, fresh numeric vectors are NOT zeroed: poison them so that unwritten results show */
  s->data = keep(malloc(bytes ? bytes : 1));
  memset(s->data, (type == VECSXP || type == STRSXP || type == CHARSXP) ? 0 : 0xA5, bytes ? bytes : 1);
  if (type == VECSXP || type == STRSXP)
    for (R_xlen_t i = 0; i < n; ++i) ((SEXP*)s->data)[i] = R_NilValue;
  return s;
}
SEXP Rf_allocMatrix(unsigned int type, int nrow, int ncol) {
  SEXP s = Rf_allocVector(type, (R_xlen_t)nrow * ncol);
  s->nrow = nrow;
  s->ncol = ncol;
  return s;
}
SEXP Rf_protect(SEXP s) {
  if (++g_protect_depth > g_protect_max) g_protect_max = g_protect_depth;
  return s;
}
void Rf_unprotect(int n) {
  g_protect_depth -= n;
  if (g_protect_depth < 0) {
    g_underflow = 1;
    g_protect_depth = 0;
  }
}
SEXP Rf_mkChar(const char* str) {
  const size_t n = strlen(str);
  SEXP s = Rf_allocVector(CHARSXP, (R_xlen_t)n);
  memcpy(s->data, str, n + 1);
  return s;
}
SEXP Rf_mkString(const char* str) {
  SEXP s = Rf_allocVector(STRSXP, 1);
  ((SEXP*)s->data)[0] = Rf_mkChar(str);
  return s;
}
SEXP Rf_mkNamed(unsigned int type, const char** names) {
  R_xlen_t n = 0;
  while (names[n][0] != '\0') ++n;
  SEXP s = Rf_allocVector(type, n);
  s->names = Rf_allocVector(STRSXP, n);
  for (R_xlen_t i = 0; i < n; ++i) ((SEXP*)s->names->data)[i] = Rf_mkChar(names[i]);
  return s;
}
SEXP Rf_ScalarLogical(int v) {
  SEXP s = Rf_allocVector(LGLSXP, 1);
  ((int*)s->data)[0] = v;
  return s;
}
SEXP Rf_ScalarInteger(int v) {
  SEXP s = Rf_allocVector(INTSXP, 1);
  ((int*)s->data)[0] = v;
  return s;
}
SEXP Rf_ScalarReal(double v) {
  SEXP s = Rf_allocVector(REALSXP, 1);
  ((double*)s->data)[0] = v;
  return s;
}
int Rf_asLogical(SEXP s) {
  if (s->length < 1) return NA_LOGICAL;
  if (s->type == LGLSXP || s->type == INTSXP) {
    const int v = ((int*)s->data)[0];
    return v == NA_INTEGER ? NA_LOGICAL : v != 0;
  }
  if (s->type == REALSXP) {
    const double v = ((double*)s->data)[0];
    return v != v ? NA_LOGICAL : v != 0.0;
  }
  return NA_LOGICAL;
}
int Rf_asInteger(SEXP s) {
  if (s->length < 1) return NA_INTEGER;
  if (s->type == INTSXP || s->type == LGLSXP) return ((int*)s->data)[0];
  if (s->type == REALSXP) {
    const double v = ((double*)s->data)[0];
    return v != v ? NA_INTEGER : (int)v;
  }
  return NA_INTEGER;
}
int Rf_isReal(SEXP s) { return s->type == REALSXP; }
int Rf_isInteger(SEXP s) { return s->type == INTSXP; }
int Rf_isMatrix(SEXP s) { return s->ncol >= 0; }
int Rf_isNull(SEXP s) { return s->type == NILSXP; }
int Rf_nrows(SEXP s) { return s->ncol >= 0 ? s->nrow : (int)s->length; }
int Rf_ncols(SEXP s) { return s->ncol >= 0 ? s->ncol : 1; }
R_xlen_t Rf_xlength(SEXP s) { return s->length; }
static void type_check(SEXP s, int type, const char* what) {
  if (s->type != type) Rf_error("%s() applied to an object of type %d", what, s->type);
}
double* REAL(SEXP s) { type_check(s, REALSXP, "REAL"); return (double*)s->data; }
int* INTEGER(SEXP s) {
  if (s->type != INTSXP && s->type != LGLSXP) Rf_error("INTEGER() applied to an object of type %d", s->type);
  return (int*)s->data;
}
int* LOGICAL(SEXP s) { type_check(s, LGLSXP, "LOGICAL"); return (int*)s->data; }
SEXP STRING_ELT(SEXP s, R_xlen_t i) {
  type_check(s, STRSXP, "STRING_ELT");
  if (i < 0 || i >= s->length) Rf_error("STRING_ELT index out of range");
  return ((SEXP*)s->data)[i];
}
const char* CHAR(SEXP s) { type_check(s, CHARSXP, "CHAR"); return (const char*)s->data; }
SEXP VECTOR_ELT(SEXP s, R_xlen_t i) {
  type_check(s, VECSXP, "VECTOR_ELT");
  if (i < 0 || i >= s->length) Rf_error("VECTOR_ELT index out of range");
  return ((SEXP*)s->data)[i];
}
SEXP SET_VECTOR_ELT(SEXP s, R_xlen_t i, SEXP v) {
  type_check(s, VECSXP, "SET_VECTOR_ELT");
  if (i < 0 || i >= s->length) Rf_error("SET_VECTOR_ELT index out of range");
  ((SEXP*)s->data)[i] = v;
  return v;
}
char* R_alloc(size_t n, int size) { return (char*)keep(malloc(n * (size_t)size + 1)); }

void Rf_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof g_error, fmt, ap);
  va_end(ap);
  if (!g_jmp_armed) {
    fprintf(stderr, "rstub: error() outside a call: %s\n", g_error);
    abort();
  }
  longjmp(g_jmp, 1);
}
void Rf_warning(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_warning, sizeof g_warning, fmt, ap);
  va_end(ap);
}

int R_registerRoutines(DllInfo* info, const void* c_methods, const R_CallMethodDef* call_methods,
                       const void* fortran_methods, const void* external_methods) {
  (void)c_methods; (void)fortran_methods; (void)external_methods;
  info->call_methods = call_methods;
  return 1;
}
int R_useDynamicSymbols(DllInfo* info, int value) {
  info->use_dynamic_symbols = value;
  return 1;
}

/* ---- driver API for the tests ------------------------------------------------------------- */
void R_init_ICIKendallTauB200(DllInfo* dll); /* the shim's initialiser */

int rstub_load_package(void) { /* what library.dynam() does: run R_init_<pkg> */
  g_dll.call_methods = NULL;
  g_dll.use_dynamic_symbols = 1;
  R_init_ICIKendallTauB200(&g_dll);
  return g_dll.call_methods != NULL;
}
int rstub_use_dynamic_symbols(void) { return g_dll.use_dynamic_symbols; }
int rstub_n_routines(void) {
  int n = 0;
  while (g_dll.call_methods && g_dll.call_methods[n].name) ++n;
  return n;
}
const char* rstub_routine_name(int k) { return g_dll.call_methods[k].name; }
int rstub_routine_nargs(int k) { return g_dll.call_methods[k].numArgs; }

typedef SEXP (*F0)(void);
typedef SEXP (*F6)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
typedef SEXP (*F8)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
typedef SEXP (*F9)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
typedef SEXP (*F12)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);

/* .Call(name, args...): looks the routine up in the REGISTERED table (not by dlsym), refuses a
 * wrong argument count like R does, and runs it under the error handler.  Returns NULL on error
 * (message: rstub_last_error()).  The PROTECT stack must be back at its depth afterwards. */
SEXP rstub_dot_call(const char* name, int nargs, SEXP* a) {
  g_error[0] = '\0';
  const R_CallMethodDef* volatile m = g_dll.call_methods;
  for (; m && m->name; ++m)
    if (strcmp(m->name, name) == 0) break;
  if (!m || !m->name) {
    snprintf(g_error, sizeof g_error, "C symbol name \"%s\" not in load table", name);
    return NULL;
  }
  if (m->numArgs != nargs) {
    snprintf(g_error, sizeof g_error, "Incorrect number of arguments (%d), expecting %d for '%s'", nargs,
             m->numArgs, name);
    return NULL;
  }
  const int depth0 = g_protect_depth;
  SEXP res = NULL;
  g_jmp_armed = 1;
  if (setjmp(g_jmp) == 0) {
    switch (nargs) {
      case 0: res = ((F0)(void (*)(void))m->fun)(); break;
      case 6: res = ((F6)(void (*)(void))m->fun)(a[0], a[1], a[2], a[3], a[4], a[5]); break;
      case 8: res = ((F8)(void (*)(void))m->fun)(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7]); break;
      case 9: res = ((F9)(void (*)(void))m->fun)(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8]); break;
      case 12:
        res = ((F12)(void (*)(void))m->fun)(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8], a[9], a[10], a[11]);
        break;
      default:
        snprintf(g_error, sizeof g_error, "rstub: no trampoline for %d arguments", nargs);
        res = NULL;
    }
  } else {
    res = NULL; /* error(): R unwinds the protect stack to the .Call frame itself; the shim is
                   expected to have balanced it BEFORE calling error(), which is what we record */
  }
  g_jmp_armed = 0;
  (void)depth0;
  return res;
}
const char* rstub_last_error(void) { return g_error; }
const char* rstub_last_warning(void) { return g_warning; }
int rstub_protect_depth(void) { return g_protect_depth; }
int rstub_protect_max(void) { return g_protect_max; }
int rstub_protect_underflow(void) { return g_underflow; }
void rstub_reset(void) {
  for (size_t i = 0; i < g_nblocks; ++i) free(g_blocks[i]);
  g_nblocks = 0;
  g_protect_depth = g_protect_max = g_underflow = 0;
  g_error[0] = g_warning[0] = '\0';
}

SEXP rstub_nil(void) { return R_NilValue; }
SEXP rstub_real(const double* v, R_xlen_t n) {
  SEXP s = Rf_allocVector(REALSXP, n);
  if (n) memcpy(s->data, v, sizeof(double) * (size_t)n);
  return s;
}
SEXP rstub_real_matrix(const double* v, int nrow, int ncol) {
  SEXP s = Rf_allocMatrix(REALSXP, nrow, ncol);
  memcpy(s->data, v, sizeof(double) * (size_t)nrow * (size_t)ncol);
  return s;
}
SEXP rstub_int(const int* v, R_xlen_t n) {
  SEXP s = Rf_allocVector(INTSXP, n);
  if (n) memcpy(s->data, v, sizeof(int) * (size_t)n);
  return s;
}
SEXP rstub_int_matrix(const int* v, int nrow, int ncol) {
  SEXP s = Rf_allocMatrix(INTSXP, nrow, ncol);
  memcpy(s->data, v, sizeof(int) * (size_t)nrow * (size_t)ncol);
  return s;
}
SEXP rstub_logical(int v) { return Rf_ScalarLogical(v); }
SEXP rstub_string(const char* s) { return Rf_mkString(s); }
int rstub_type(SEXP s) { return s->type; }
R_xlen_t rstub_length(SEXP s) { return s->length; }
int rstub_nrow(SEXP s) { return s->nrow; }
int rstub_ncol(SEXP s) { return s->ncol; }
void* rstub_data(SEXP s) { return s->data; }
SEXP rstub_list_get(SEXP list, const char* name) {
  if (list->type != VECSXP || !list->names) return NULL;
  for (R_xlen_t i = 0; i < list->length; ++i)
    if (strcmp((const char*)((SEXP*)list->names->data)[i]->data, name) == 0) return ((SEXP*)list->data)[i];
  return NULL;
}
const char* rstub_list_name(SEXP list, R_xlen_t i) {
  return (list->names && i < list->length) ? (const char*)((SEXP*)list->names->data)[i]->data : "";
}
double rstub_na_real(void) { return R_NaReal; }
