/* icikt_shim.c -- the .Call boundary between R and libicikt_b200.so.
 *
 * Replaces, for the ICI-Kendall-tau path of the reference package,
 *   SEXP _ICIKendallTau_ici_kt(SEXP x6)        src/RcppExports.cpp:83-96   (one pair per call)
 *   the CallEntries[] registration             src/RcppExports.cpp:113-128
 * with two BATCHED entry points (all pairs / explicit pair list) so that the R pair loop
 * (R/kendalltau.R:158, ici_split :280-308, kt_split :310-354) becomes one call.
 *
 * Plain C against R's C API: no Rcpp, no C++ exceptions, nothing is thrown across the C
 * boundary.  The library reports failures by return code + icikt_last_error(); the shim turns
 * them into R errors only after every PROTECT is balanced.  Inputs are never modified (R is
 * copy-on-modify; the reference clones, src/kendallc.cpp:187-188); outputs are allocated
 * here as REALSXP/INTSXP and filled by the library.  Degenerate pairs come back as NaN with a
 * non-zero status; they are rewritten to NA_real_ here because testthat's waldo distinguishes
 * NA from NaN (tests/testthat/test-kendall-tau.R:45-51 expects NA).  The genuine NaN p-value of
 * the n == 2 case (status 0) is left alone.
 */
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>
#include <string.h>

#include "icikt_b200.h"

static void fill_opts(icikt_opts* o, SEXP perspective, SEXP alternative, SEXP continuity,
                      SEXP include_diag, SEXP na_inf, SEXP device) {
  icikt_default_opts(o);
  const char* p = CHAR(STRING_ELT(perspective, 0));
  /* any string other than "local" behaves as global, src/kendallc.cpp:180 */
  o->perspective = strcmp(p, "local") == 0 ? ICIKT_PERSPECTIVE_LOCAL
                 : strcmp(p, "complete") == 0 ? ICIKT_PERSPECTIVE_COMPLETE /* kt_fast, pairwise.complete.obs */
                                              : ICIKT_PERSPECTIVE_GLOBAL;
  const char* a = CHAR(STRING_ELT(alternative, 0));
  if (strcmp(a, "two.sided") == 0) o->alternative = ICIKT_ALT_TWO_SIDED;
  else if (strcmp(a, "less") == 0) o->alternative = ICIKT_ALT_LESS;
  else if (strcmp(a, "greater") == 0) o->alternative = ICIKT_ALT_GREATER;
  else o->alternative = ICIKT_ALT_OTHER; /* p-value stays 0, src/kendallc.cpp:323-332 */
  o->continuity = asLogical(continuity) == TRUE;
  o->include_diag = asLogical(include_diag) == TRUE;
  o->na_inf = asLogical(na_inf) == TRUE;
  o->device = asInteger(device);
}

/* named list(raw, pvalue, taumax, completeness, status, max_taumax) of length-P vectors */
static SEXP make_result(R_xlen_t P, double** raw, double** pv, double** tm, double** comp, int** status,
                        SEXP* max_out) {
  const char* names[] = {"raw", "pvalue", "taumax", "completeness", "status", "max_taumax", ""};
  SEXP res = PROTECT(mkNamed(VECSXP, names));
  SEXP v;
  v = allocVector(REALSXP, P); SET_VECTOR_ELT(res, 0, v); *raw = REAL(v);
  v = allocVector(REALSXP, P); SET_VECTOR_ELT(res, 1, v); *pv = REAL(v);
  v = allocVector(REALSXP, P); SET_VECTOR_ELT(res, 2, v); *tm = REAL(v);
  v = allocVector(REALSXP, P); SET_VECTOR_ELT(res, 3, v); *comp = REAL(v);
  v = allocVector(INTSXP, P);  SET_VECTOR_ELT(res, 4, v); *status = INTEGER(v);
  v = allocVector(REALSXP, 1); SET_VECTOR_ELT(res, 5, v); *max_out = v;
  return res; /* still protected: caller UNPROTECTs */
}

static void na_for_degenerate(R_xlen_t P, double* raw, double* pv, double* tm, double* comp, const int* status) {
  for (R_xlen_t k = 0; k < P; ++k)
    if (status[k] != ICIKT_STATUS_OK) raw[k] = pv[k] = tm[k] = comp[k] = NA_REAL;
}

/* .Call("C_icikt_all_pairs", data, global_na, perspective, alternative, continuity,
 *       include_diag, na_inf, device)
 * data: double matrix, features x samples (column-major, exactly what the library wants). */
SEXP C_icikt_all_pairs(SEXP data, SEXP global_na, SEXP perspective, SEXP alternative, SEXP continuity,
                       SEXP include_diag, SEXP na_inf, SEXP device) {
  if (!isReal(data) || !isMatrix(data)) error("`data` must be a double matrix");
  if (!isReal(global_na)) error("`global_na` must be a double vector");
  const int64_t n = nrows(data), C = ncols(data);
  icikt_opts o;
  fill_opts(&o, perspective, alternative, continuity, include_diag, na_inf, device);
  const R_xlen_t P = (R_xlen_t)(C * (C - 1) / 2 + (o.include_diag ? C : 0));
  double *raw, *pv, *tm, *comp, mx = NA_REAL;
  int* status;
  SEXP mxs;
  SEXP res = make_result(P, &raw, &pv, &tm, &comp, &status, &mxs);
  /* `device`: one ordinal, or several -> the pair order is sliced over those GPUs inside the call */
  const int rc = (isInteger(device) && XLENGTH(device) > 1)
                     ? icikt_all_pairs_multi(REAL(data), n, C, n, REAL(global_na), (int32_t)XLENGTH(global_na), &o,
                                             INTEGER(device), (int32_t)XLENGTH(device), raw, pv, tm, comp, status,
                                             NULL, &mx, NULL)
                     : icikt_all_pairs(REAL(data), n, C, n, REAL(global_na), (int32_t)XLENGTH(global_na), &o, raw,
                                       pv, tm, comp, status, NULL, &mx, NULL);
  if (rc != ICIKT_OK) {
    UNPROTECT(1);
    error("libicikt_b200 (%d): %s", rc, icikt_last_error());
  }
  na_for_degenerate(P, raw, pv, tm, comp, status);
  REAL(mxs)[0] = ISNAN(mx) ? NA_REAL : mx;
  UNPROTECT(1);
  return res;
}

/* .Call("C_icikt_pair_list", data, global_na, i, j, perspective, alternative, continuity,
 *       na_inf, device)  -- i, j: 1-based integer column indices of the pairs */
SEXP C_icikt_pair_list(SEXP data, SEXP global_na, SEXP pi, SEXP pj, SEXP perspective, SEXP alternative,
                       SEXP continuity, SEXP na_inf, SEXP device) {
  if (!isReal(data) || !isMatrix(data)) error("`data` must be a double matrix");
  if (!isInteger(pi) || !isInteger(pj) || XLENGTH(pi) != XLENGTH(pj)) error("`i` and `j` must be integer vectors of one length");
  const int64_t n = nrows(data), C = ncols(data);
  const R_xlen_t P = XLENGTH(pi);
  icikt_opts o;
  fill_opts(&o, perspective, alternative, continuity, ScalarLogical(FALSE), na_inf, device);
  int32_t* zi = (int32_t*)R_alloc((size_t)P, sizeof(int32_t)); /* freed by R at the end of .Call */
  int32_t* zj = (int32_t*)R_alloc((size_t)P, sizeof(int32_t));
  for (R_xlen_t k = 0; k < P; ++k) {
    zi[k] = INTEGER(pi)[k] - 1;
    zj[k] = INTEGER(pj)[k] - 1;
  }
  double *raw, *pv, *tm, *comp, mx = NA_REAL;
  int* status;
  SEXP mxs;
  SEXP res = make_result(P, &raw, &pv, &tm, &comp, &status, &mxs);
  const int rc = icikt_pair_list(REAL(data), n, C, n, REAL(global_na), (int32_t)XLENGTH(global_na), zi, zj,
                                 (int64_t)P, &o, raw, pv, tm, comp, status, NULL, &mx, NULL);
  if (rc != ICIKT_OK) {
    UNPROTECT(1);
    error("libicikt_b200 (%d): %s", rc, icikt_last_error());
  }
  na_for_degenerate(P, raw, pv, tm, comp, status);
  REAL(mxs)[0] = ISNAN(mx) ? NA_REAL : mx;
  UNPROTECT(1);
  return res;
}

/* .Call("C_icikt_matrices", data, global_na, i, j, perspective, alternative, continuity, na_inf,
 *       device, scale_max, diag_good, n_good)
 * scale_and_reshape (R/kendalltau.R:357-421) done on the device: returns the five C x C matrices
 * cor, raw, pvalue, taumax, completeness (degenerate pairs already NA_real_), status_counts
 * (pairs per status class 0..9) and max_taumax.  i = j = NULL: all pairs (+ the diagonal pairs
 * iff !diag_good); else 1-based column indices of the pairs (include_only).  n_good: integer[C],
 * colSums(!exclude_loc).                                                                    */
SEXP C_icikt_matrices(SEXP data, SEXP global_na, SEXP pi, SEXP pj, SEXP perspective, SEXP alternative,
                      SEXP continuity, SEXP na_inf, SEXP device, SEXP scale_max, SEXP diag_good, SEXP n_good) {
  if (!isReal(data) || !isMatrix(data)) error("`data` must be a double matrix");
  if (!isReal(global_na)) error("`global_na` must be a double vector");
  const int64_t n = nrows(data), C = ncols(data);
  if (!isInteger(n_good) || XLENGTH(n_good) != C) error("`n_good` must be an integer vector with one entry per column");
  const int have_list = !isNull(pi);
  if (have_list && (!isInteger(pi) || !isInteger(pj) || XLENGTH(pi) != XLENGTH(pj)))
    error("`i` and `j` must be integer vectors of one length");
  icikt_opts o;
  fill_opts(&o, perspective, alternative, continuity, ScalarLogical(FALSE), na_inf, device);
  const R_xlen_t P = have_list ? XLENGTH(pi) : 0;
  int32_t *zi = NULL, *zj = NULL;
  if (have_list) {
    zi = (int32_t*)R_alloc((size_t)P, sizeof(int32_t));
    zj = (int32_t*)R_alloc((size_t)P, sizeof(int32_t));
    for (R_xlen_t k = 0; k < P; ++k) {
      zi[k] = INTEGER(pi)[k] - 1;
      zj[k] = INTEGER(pj)[k] - 1;
    }
  }
  const char* names[] = {"cor", "raw", "pvalue", "taumax", "completeness", "status_counts", "max_taumax", ""};
  SEXP res = PROTECT(mkNamed(VECSXP, names));
  double* m[5];
  for (int k = 0; k < 5; ++k) {
    SEXP v = allocMatrix(REALSXP, (int)C, (int)C);
    SET_VECTOR_ELT(res, k, v);
    m[k] = REAL(v);
  }
  SEXP counts = allocVector(REALSXP, ICIKT_NSTATUS);
  SET_VECTOR_ELT(res, 5, counts);
  SEXP mxs = allocVector(REALSXP, 1);
  SET_VECTOR_ELT(res, 6, mxs);
  int64_t hist[ICIKT_NSTATUS];
  double mx = NA_REAL;
  /* `device`: one ordinal, or (all pairs only) several -> every GPU computes a slice of the pair order and
   * fills and returns its own block of columns of the five matrices */
  const int multi = !have_list && isInteger(device) && XLENGTH(device) > 1;
  const int rc = multi
                     ? icikt_matrices_multi(REAL(data), n, C, n, REAL(global_na), (int32_t)XLENGTH(global_na), &o,
                                            INTEGER(device), (int32_t)XLENGTH(device), asLogical(scale_max) == TRUE,
                                            asLogical(diag_good) == TRUE, INTEGER(n_good), m[0], m[1], m[2], m[3],
                                            m[4], hist, &mx, NULL)
                     : icikt_matrices(REAL(data), n, C, n, REAL(global_na), (int32_t)XLENGTH(global_na), zi, zj,
                                      (int64_t)P, &o, asLogical(scale_max) == TRUE, asLogical(diag_good) == TRUE,
                                      INTEGER(n_good), m[0], m[1], m[2], m[3], m[4], hist, &mx, NULL);
  if (rc != ICIKT_OK) {
    UNPROTECT(1);
    error("libicikt_b200 (%d): %s", rc, icikt_last_error());
  }
  for (int k = 0; k < ICIKT_NSTATUS; ++k) REAL(counts)[k] = (double)hist[k];
  REAL(mxs)[0] = ISNAN(mx) ? NA_REAL : mx;
  UNPROTECT(1);
  return res;
}

/* .Call("C_icikt_pairwise_completeness", data, global_na, i, j, device, want_matrix)
 * pairwise_completeness (R/kendalltau.R:563-629).  global_na as the user gave it (NA and Inf
 * entries select classes, R/utils.R:6-15).  i = j = NULL: all pairs in combn order followed by
 * the diagonal pairs; want_matrix then adds the symmetric C x C completeness matrix.        */
SEXP C_icikt_pairwise_completeness(SEXP data, SEXP global_na, SEXP pi, SEXP pj, SEXP device, SEXP want_matrix) {
  if (!isReal(data) || !isMatrix(data)) error("`data` must be a double matrix");
  if (!isReal(global_na)) error("`global_na` must be a double vector");
  const int64_t n = nrows(data), C = ncols(data);
  const int have_list = !isNull(pi), want_m = asLogical(want_matrix) == TRUE && !have_list;
  if (have_list && (!isInteger(pi) || !isInteger(pj) || XLENGTH(pi) != XLENGTH(pj)))
    error("`i` and `j` must be integer vectors of one length");
  /* NA_real_ in global_na is a NaN, which is what the library looks for */
  const R_xlen_t P = have_list ? XLENGTH(pi) : (R_xlen_t)(C * (C - 1) / 2 + C);
  int32_t *zi = NULL, *zj = NULL;
  if (have_list) {
    zi = (int32_t*)R_alloc((size_t)P, sizeof(int32_t));
    zj = (int32_t*)R_alloc((size_t)P, sizeof(int32_t));
    for (R_xlen_t k = 0; k < P; ++k) {
      zi[k] = INTEGER(pi)[k] - 1;
      zj[k] = INTEGER(pj)[k] - 1;
    }
  }
  const char* names[] = {"missingness", "completeness", "matrix", ""};
  SEXP res = PROTECT(mkNamed(VECSXP, names));
  SEXP miss = allocVector(INTSXP, P), comp = allocVector(REALSXP, P);
  SET_VECTOR_ELT(res, 0, miss);
  SET_VECTOR_ELT(res, 1, comp);
  double* mat = NULL;
  if (want_m) {
    SEXP v = allocMatrix(REALSXP, (int)C, (int)C);
    SET_VECTOR_ELT(res, 2, v);
    mat = REAL(v);
  }
  const int rc = icikt_pairwise_completeness(REAL(data), n, C, n, REAL(global_na), (int32_t)XLENGTH(global_na),
                                             asInteger(device), zi, zj, (int64_t)P, INTEGER(miss), REAL(comp), mat);
  if (rc != ICIKT_OK) {
    UNPROTECT(1);
    error("libicikt_b200 (%d): %s", rc, icikt_last_error());
  }
  UNPROTECT(1);
  return res;
}

SEXP C_icikt_device_count(void) { return ScalarInteger(icikt_device_count()); }

SEXP C_icikt_release(void) {
  icikt_release_workspace();
  return R_NilValue;
}

/* same registration discipline as src/RcppExports.cpp:113-128 of the reference */
static const R_CallMethodDef CallEntries[] = {
    {"C_icikt_all_pairs", (DL_FUNC)&C_icikt_all_pairs, 8},
    {"C_icikt_pair_list", (DL_FUNC)&C_icikt_pair_list, 9},
    {"C_icikt_matrices", (DL_FUNC)&C_icikt_matrices, 12},
    {"C_icikt_pairwise_completeness", (DL_FUNC)&C_icikt_pairwise_completeness, 6},
    {"C_icikt_device_count", (DL_FUNC)&C_icikt_device_count, 0},
    {"C_icikt_release", (DL_FUNC)&C_icikt_release, 0},
    {NULL, NULL, 0}};

void R_init_ICIKendallTauB200(DllInfo* dll) {
  R_registerRoutines(dll, NULL, CallEntries, NULL, NULL);
  R_useDynamicSymbols(dll, FALSE);
}
