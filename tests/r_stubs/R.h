/* Minimal stand-in for R's <R.h>, enough to type-check covidcurve_b200/r/ccb200_shim.c where R is not installed
 * (tests/test_abi.py::test_r_shim_compiles).  Declarations only, written from the documented R C API ("Writing R
 * Extensions", section 5-6); nothing here is linked. */
#ifndef CCB_STUB_R_H
#define CCB_STUB_R_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
void Rf_error(const char* fmt, ...);
void Rf_warning(const char* fmt, ...);
char* R_alloc(size_t n, int size); /* transient storage, reclaimed at the end of the .Call */
#ifdef __cplusplus
}
#endif
#endif
