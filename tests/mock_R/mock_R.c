/* Implementation of the mock R API (see Rinternals.h). */
#define _POSIX_C_SOURCE 200809L
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "R.h"
#include "R_ext/Rdynload.h"

static struct mock_sexp nil_value = {NILSXP, 0, -1, 0, NULL, NULL, NULL};
SEXP R_NilValue = &nil_value;
static SEXP finalizable[64];
static int nfinalizable = 0;

static SEXP mk(SEXPTYPE t, long n)
{
    SEXP s = (SEXP)calloc(1, sizeof(*s));
    size_t el = t == REALSXP ? sizeof(double) : t == INTSXP ? sizeof(int) : sizeof(SEXP);
    s->type = t; s->length = n; s->ncol = -1; s->nrow = (int)n;
    s->data = calloc(n > 0 ? n : 1, el);
    if (t == VECSXP) for (long i = 0; i < n; ++i) ((SEXP *)s->data)[i] = R_NilValue;
    return s;
}
static void need(SEXP x, SEXPTYPE t, const char *what)
{
    if (!x || x->type != t) { fprintf(stderr, "mock R: %s on an object of type %d\n", what, x ? (int)x->type : -1); exit(3); }
}
double *REAL(SEXP x) { need(x, REALSXP, "REAL()"); return (double *)x->data; }
int *INTEGER(SEXP x) { need(x, INTSXP, "INTEGER()"); return (int *)x->data; }
int Rf_nrows(SEXP x) { return x->ncol < 0 ? (int)x->length : x->nrow; }
int Rf_ncols(SEXP x) { return x->ncol < 0 ? 1 : x->ncol; }
long Rf_xlength(SEXP x) { return x->length; }
int Rf_isNull(SEXP x) { return x == R_NilValue || x->type == NILSXP; }
int Rf_asInteger(SEXP x)
{
    if (x->type == INTSXP && x->length >= 1) return ((int *)x->data)[0];
    if (x->type == REALSXP && x->length >= 1) return (int)((double *)x->data)[0];
    return NA_INTEGER;
}
SEXP Rf_allocVector(SEXPTYPE t, long n) { return mk(t, n); }
SEXP Rf_allocMatrix(SEXPTYPE t, int nrow, int ncol)
{
    SEXP s = mk(t, (long)nrow * ncol);
    s->nrow = nrow; s->ncol = ncol;
    return s;
}
SEXP Rf_mkNamed(SEXPTYPE t, const char **names)
{
    long n = 0;
    while (names[n][0]) ++n;
    SEXP s = mk(t, n);
    const char **copy = (const char **)calloc(n + 1, sizeof(char *));     /* R copies the strings; the caller's array may be a local */
    for (long i = 0; i < n; ++i) copy[i] = strdup(names[i]);
    copy[n] = "";
    s->names = copy;
    return s;
}
SEXP VECTOR_ELT(SEXP x, long i) { need(x, VECSXP, "VECTOR_ELT()"); return ((SEXP *)x->data)[i]; }
SEXP SET_VECTOR_ELT(SEXP x, long i, SEXP v) { need(x, VECSXP, "SET_VECTOR_ELT()"); ((SEXP *)x->data)[i] = v; return v; }
void Rf_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    fprintf(stderr, "Error: ");
    vfprintf(stderr, fmt, ap);
    fprintf(stderr, "\n");
    va_end(ap);
    exit(2);      /* R would longjmp to top level */
}
char *R_alloc(size_t n, int size) { return (char *)calloc(n ? n : 1, (size_t)size); }
SEXP R_MakeExternalPtr(void *p, SEXP tag, SEXP prot)
{
    (void)tag; (void)prot;
    SEXP s = (SEXP)calloc(1, sizeof(*s));
    s->type = EXTPTRSXP; s->ncol = -1; s->data = p;
    return s;
}
void *R_ExternalPtrAddr(SEXP s) { need(s, EXTPTRSXP, "R_ExternalPtrAddr()"); return s->data; }
void R_ClearExternalPtr(SEXP s) { need(s, EXTPTRSXP, "R_ClearExternalPtr()"); s->data = NULL; }
void R_RegisterCFinalizerEx(SEXP s, void (*fun)(SEXP), Rboolean onexit)
{
    (void)onexit;
    s->finalizer = fun;
    if (nfinalizable < 64) finalizable[nfinalizable++] = s;
}
SEXP mock_scalar_int(int v) { SEXP s = mk(INTSXP, 1); ((int *)s->data)[0] = v; return s; }
void mock_run_finalizers(void)
{
    for (int i = 0; i < nfinalizable; ++i) if (finalizable[i]->finalizer) finalizable[i]->finalizer(finalizable[i]);
    nfinalizable = 0;
}
int R_registerRoutines(DllInfo *info, const void *c, const R_CallMethodDef *call, const void *f, const void *e)
{
    (void)c; (void)f; (void)e;
    info->calls = call;
    return 1;
}
Rboolean R_useDynamicSymbols(DllInfo *info, Rboolean value) { info->dynamic = value; return value; }
