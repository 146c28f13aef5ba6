/* Minimal stand-in for R's C API, just enough to COMPILE AND RUN r_package/src/gr2m_shim.c without R
 * (test infrastructure: tests/test_r_shim.py).  SEXPs are malloc'ed records; nothing is garbage collected.
 * Semantics kept: column-major matrices, NA_INTEGER = INT_MIN, R_NilValue, named lists, external pointers
 * with C finalizers, registered .Call routines with argument counts. */
#ifndef MOCK_RINTERNALS_H
#define MOCK_RINTERNALS_H
#include <limits.h>
#include <math.h>
#include <stddef.h>
#include <stdlib.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum { NILSXP = 0, INTSXP = 13, REALSXP = 14, VECSXP = 19, EXTPTRSXP = 22 } SEXPTYPE;
typedef struct mock_sexp {
    SEXPTYPE type;
    int nrow, ncol;            /* ncol < 0: plain vector */
    long length;
    void *data;                /* double[], int[], SEXP[] or the external pointer's address */
    const char **names;        /* VECSXP made by Rf_mkNamed */
    void (*finalizer)(struct mock_sexp *);
} *SEXP;
typedef int Rboolean;
#ifndef TRUE
#define TRUE 1
#define FALSE 0
#endif

extern SEXP R_NilValue;
#define NA_INTEGER INT_MIN
#define R_FINITE(x) isfinite(x)
#define PROTECT(x) (x)
#define UNPROTECT(n) ((void)(n))

double *REAL(SEXP x);
int *INTEGER(SEXP x);
int Rf_nrows(SEXP x);
int Rf_ncols(SEXP x);
int Rf_isNull(SEXP x);
int Rf_asInteger(SEXP x);
SEXP Rf_allocVector(SEXPTYPE t, long n);
SEXP Rf_allocMatrix(SEXPTYPE t, int nrow, int ncol);
SEXP Rf_mkNamed(SEXPTYPE t, const char **names);
SEXP VECTOR_ELT(SEXP x, long i);
SEXP SET_VECTOR_ELT(SEXP x, long i, SEXP v);
void Rf_error(const char *fmt, ...);
char *R_alloc(size_t n, int size);
SEXP R_MakeExternalPtr(void *p, SEXP tag, SEXP prot);
void *R_ExternalPtrAddr(SEXP s);
void R_ClearExternalPtr(SEXP s);
void R_RegisterCFinalizerEx(SEXP s, void (*fun)(SEXP), Rboolean onexit);

/* helpers of the mock itself (used by the test driver) */
long Rf_xlength(SEXP x);
#define R_Calloc(n, T) ((T *)calloc((size_t)(n), sizeof(T)))
#define R_Free(p) (free((void *)(p)), (p) = NULL)

SEXP mock_scalar_int(int v);
void mock_run_finalizers(void);

#ifdef __cplusplus
}
#endif
#endif
