/* tests/r_stub/R_ext/Rdynload.h -- TEST INFRASTRUCTURE, see ../Rinternals.h. */
#ifndef RSTUB_RDYNLOAD_H
#define RSTUB_RDYNLOAD_H
#ifdef __cplusplus
extern "C" {
#endif
typedef void* (*DL_FUNC)(void);
typedef struct {
  const char* name;
  DL_FUNC fun;
  int numArgs;
} R_CallMethodDef;
typedef struct rstub_dllinfo {
  const R_CallMethodDef* call_methods;
  int use_dynamic_symbols;
} DllInfo;
int R_registerRoutines(DllInfo* info, const void* c_methods, const R_CallMethodDef* call_methods,
                       const void* fortran_methods, const void* external_methods);
int R_useDynamicSymbols(DllInfo* info, int value);
#ifdef __cplusplus
}
#endif
#endif
