/*
 * gr2m_shim.c -- .Call glue between R and the C ABI of include/gr2m_b200.h.  Zero logic: unpack SEXPs,
 * call, wrap.  Never run under R (no R in the build image); compiled and EXECUTED against tests/mock_R by
 * tests/test_r_shim.py, which compares every routine's output with the ctypes path bit for bit.
 *
 * Seams replaced in the reference (file:line under hllauca/GR2MSemiDistr):
 *   C_gr2m_run            parLapply + airGR::RunModel_GR2M per subbasin   R/Run_GR2MSemiDistr.R:126-189
 *   C_gr2m_basin / C_gr2m_objective   the body of OFUN                    R/Optim_GR2MSemiDistr.R:125-217
 *   C_wfa_route           the AreaD8 month loop                           R/Routing_GR2MSemiDistr.R:100-123
 *   C_wfa_router / C_gr2m_objective_routed   Run -> Routing -> hydroGOF for a batch of candidates (no reference seam:
 *                         the reference can only calibrate on the outlet sum)
 */
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>
#include <stdint.h>
#include "gr2m_b200.h"

static void fail(int rc) { if (rc != GR2M_OK) Rf_error("libgr2m_b200 (%d): %s", rc, gr2m_last_error()); }
static double *opt(SEXP x) { return Rf_isNull(x) ? NULL : REAL(x); }

SEXP C_gr2m_run(SEXP P, SEXP E, SEXP area, SEXP ndays, SEXP region, SEXP nreg, SEXP params, SEXP S0, SEXP R0,
                SEXP flags)
{
    int ntime = Rf_nrows(P), ncell = Rf_ncols(P);
    const char *nm[] = {"PR", "AE", "SM", "RU", "QS", "S_end", "R_end", ""};
    SEXP out = PROTECT(Rf_mkNamed(VECSXP, nm));
    for (int i = 0; i < 5; ++i) SET_VECTOR_ELT(out, i, Rf_allocMatrix(REALSXP, ntime, ncell));
    for (int i = 5; i < 7; ++i) SET_VECTOR_ELT(out, i, Rf_allocVector(REALSXP, ncell));
    fail(gr2m_run_once(ntime, ncell, REAL(P), REAL(E), REAL(area), INTEGER(ndays), INTEGER(region),
                       Rf_asInteger(nreg), REAL(params), opt(S0), opt(R0), Rf_asInteger(flags),
                       REAL(VECTOR_ELT(out, 0)), REAL(VECTOR_ELT(out, 1)), REAL(VECTOR_ELT(out, 2)),
                       REAL(VECTOR_ELT(out, 3)), REAL(VECTOR_ELT(out, 4)), REAL(VECTOR_ELT(out, 5)),
                       REAL(VECTOR_ELT(out, 6))));
    UNPROTECT(1);
    return out;
}

static void basin_finalizer(SEXP p)
{
    gr2m_basin *b = (gr2m_basin *)R_ExternalPtrAddr(p);
    if (b) { gr2m_basin_destroy(b); R_ClearExternalPtr(p); }
}

SEXP C_gr2m_basin(SEXP P, SEXP E, SEXP area, SEXP ndays, SEXP region, SEXP nreg)
{
    gr2m_basin *b = NULL;
    fail(gr2m_basin_create(Rf_nrows(P), Rf_ncols(P), REAL(P), REAL(E), REAL(area), INTEGER(ndays), INTEGER(region),
                           Rf_asInteger(nreg), &b));
    SEXP p = PROTECT(R_MakeExternalPtr(b, R_NilValue, R_NilValue));
    R_RegisterCFinalizerEx(p, basin_finalizer, TRUE);
    UNPROTECT(1);
    return p;
}

/* params: numeric matrix [4*nreg x nsets] (one candidate per COLUMN == row-major [nsets x 4*nreg] for C) */
SEXP C_gr2m_objective(SEXP ptr, SEXP params, SEXP qobs, SEXP warmup, SEXP which, SEXP flags)
{
    gr2m_basin *b = (gr2m_basin *)R_ExternalPtrAddr(ptr);
    if (!b) Rf_error("basin handle already released");
    int nsets = Rf_ncols(params);
    SEXP of = PROTECT(Rf_allocVector(REALSXP, nsets));
    fail(gr2m_objective_batch(b, REAL(params), nsets, REAL(qobs), Rf_asInteger(warmup), Rf_asInteger(which),
                              Rf_asInteger(flags), REAL(of), NULL, NULL));
    UNPROTECT(1);
    return of;
}

/* flowdir: integer matrix as raster::as.matrix gives it (nrow x ncol, column-major) -> row-major int16 */
SEXP C_wfa_route(SEXP flowdir, SEXP QS, SEXP src_cell, SEXP poly_off, SEXP poly_cells, SEXP flags)
{
    int nrow = Rf_nrows(flowdir), ncol = Rf_ncols(flowdir), ntime = Rf_nrows(QS), nsub = Rf_ncols(QS);
    int16_t *d = (int16_t *)R_alloc((size_t)nrow * ncol, sizeof(int16_t));
    const int *f = INTEGER(flowdir);
    for (int r = 0; r < nrow; ++r)
        for (int c = 0; c < ncol; ++c) {
            int v = f[(size_t)c * nrow + r];
            d[(size_t)r * ncol + c] = (v == NA_INTEGER || v < 1 || v > 8) ? (int16_t)-32768 : (int16_t)v;
        }
    wfa_plan *plan = NULL;
    fail(wfa_plan_create(d, nrow, ncol, 0, &plan));
    SEXP QR = PROTECT(Rf_allocMatrix(REALSXP, ntime, nsub));
    int rc = wfa_route(plan, ntime, nsub, REAL(QS), INTEGER(src_cell), INTEGER(poly_off), INTEGER(poly_cells),
                       Rf_asInteger(flags), REAL(QR));
    wfa_plan_destroy(plan);
    fail(rc);
    UNPROTECT(1);
    return QR;
}

/* ---- routed objective: plan + routing operator kept behind one external pointer --------------------------- */
typedef struct { wfa_plan *plan; wfa_router *router; } routing_handle;

static void routing_finalizer(SEXP p)
{
    routing_handle *h = (routing_handle *)R_ExternalPtrAddr(p);
    if (h) {
        if (h->router) wfa_router_destroy(h->router);
        if (h->plan) wfa_plan_destroy(h->plan);
        R_Free(h);
        R_ClearExternalPtr(p);
    }
}

/* flowdir as in C_wfa_route; src_cell / poly_off / poly_cells 0-based -> external pointer to (plan, operator) */
SEXP C_wfa_router(SEXP flowdir, SEXP src_cell, SEXP poly_off, SEXP poly_cells)
{
    int nrow = Rf_nrows(flowdir), ncol = Rf_ncols(flowdir), nsub = (int)Rf_xlength(src_cell);
    int16_t *d = (int16_t *)R_alloc((size_t)nrow * ncol, sizeof(int16_t));
    const int *f = INTEGER(flowdir);
    for (int r = 0; r < nrow; ++r)
        for (int c = 0; c < ncol; ++c) {
            int v = f[(size_t)c * nrow + r];
            d[(size_t)r * ncol + c] = (v == NA_INTEGER || v < 1 || v > 8) ? (int16_t)-32768 : (int16_t)v;
        }
    routing_handle *h = R_Calloc(1, routing_handle);
    int rc = wfa_plan_create(d, nrow, ncol, 0, &h->plan);
    if (rc == GR2M_OK) rc = wfa_router_create(h->plan, nsub, INTEGER(src_cell), INTEGER(poly_off), INTEGER(poly_cells), 0, &h->router);
    if (rc != GR2M_OK) {
        if (h->plan) wfa_plan_destroy(h->plan);
        R_Free(h);
        fail(rc);
    }
    SEXP p = PROTECT(R_MakeExternalPtr(h, R_NilValue, R_NilValue));
    R_RegisterCFinalizerEx(p, routing_finalizer, TRUE);
    UNPROTECT(1);
    return p;
}

/* OFUN on the routed discharge of polygon `gauge` (0-based); params as in C_gr2m_objective */
SEXP C_gr2m_objective_routed(SEXP ptr, SEXP routing, SEXP params, SEXP qobs, SEXP gauge, SEXP warmup, SEXP which, SEXP flags)
{
    gr2m_basin *b = (gr2m_basin *)R_ExternalPtrAddr(ptr);
    routing_handle *h = (routing_handle *)R_ExternalPtrAddr(routing);
    if (!b || !h) Rf_error("basin or routing handle already released");
    int nsets = Rf_ncols(params);
    SEXP of = PROTECT(Rf_allocVector(REALSXP, nsets));
    fail(gr2m_objective_routed_batch(b, h->router, REAL(params), nsets, REAL(qobs), Rf_asInteger(gauge), Rf_asInteger(warmup),
                                     Rf_asInteger(which), Rf_asInteger(flags), REAL(of), NULL, NULL));
    UNPROTECT(1);
    return of;
}

/* dem: numeric matrix (nrow x ncol, column-major, NA = nodata) -> integer matrix of TauDEM D8 codes (NA = nodata) */
SEXP C_wfa_flowdir(SEXP dem)
{
    int nrow = Rf_nrows(dem), ncol = Rf_ncols(dem);
    size_t n = (size_t)nrow * ncol;
    double *z = (double *)R_alloc(n, sizeof(double));
    uint8_t *valid = (uint8_t *)R_alloc(n, 1);
    int16_t *fd = (int16_t *)R_alloc(n, sizeof(int16_t));
    const double *d = REAL(dem);
    for (int r = 0; r < nrow; ++r)
        for (int c = 0; c < ncol; ++c) {
            double v = d[(size_t)c * nrow + r];
            valid[(size_t)r * ncol + c] = (uint8_t)(R_FINITE(v) ? 1 : 0);
            z[(size_t)r * ncol + c] = R_FINITE(v) ? v : 0.0;
        }
    fail(wfa_flowdir_from_dem(z, valid, nrow, ncol, fd, NULL));
    SEXP out = PROTECT(Rf_allocMatrix(INTSXP, nrow, ncol));
    int *o = INTEGER(out);
    for (int r = 0; r < nrow; ++r)
        for (int c = 0; c < ncol; ++c) {
            int16_t v = fd[(size_t)r * ncol + c];
            o[(size_t)c * nrow + r] = v > 0 ? (int)v : NA_INTEGER;
        }
    UNPROTECT(1);
    return out;
}

static const R_CallMethodDef calls[] = {
    {"C_gr2m_run", (DL_FUNC)&C_gr2m_run, 10},
    {"C_gr2m_basin", (DL_FUNC)&C_gr2m_basin, 6},
    {"C_gr2m_objective", (DL_FUNC)&C_gr2m_objective, 6},
    {"C_wfa_route", (DL_FUNC)&C_wfa_route, 6},
    {"C_wfa_router", (DL_FUNC)&C_wfa_router, 4},
    {"C_gr2m_objective_routed", (DL_FUNC)&C_gr2m_objective_routed, 8},
    {"C_wfa_flowdir", (DL_FUNC)&C_wfa_flowdir, 1},
    {NULL, NULL, 0}};

void R_init_GR2MSemiDistrB200(DllInfo *dll)
{
    R_registerRoutines(dll, NULL, calls, NULL, NULL);
    R_useDynamicSymbols(dll, FALSE);
}
