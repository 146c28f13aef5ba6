/* Drives r_package/src/gr2m_shim.c through its REGISTERED .Call table, as R would after
 * library(GR2MSemiDistrB200): R_init_GR2MSemiDistrB200 -> look the routine up by name -> check the argument
 * count -> call.  Inputs / outputs are raw little-endian arrays in the directory given on the command line
 * (written and read back by tests/test_r_shim.py, which compares them with the ctypes path).
 *   usage: shim_driver <dir>
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "R.h"
#include "R_ext/Rdynload.h"

void R_init_GR2MSemiDistrB200(DllInfo *dll);
static DllInfo dll;
static const char *dir;

static void *slurp(const char *name, size_t *bytes)
{
    char path[1024];
    snprintf(path, sizeof path, "%s/%s", dir, name);
    FILE *f = fopen(path, "rb");
    if (!f) { fprintf(stderr, "cannot open %s\n", path); exit(4); }
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    void *p = malloc(n > 0 ? n : 1);
    if (fread(p, 1, n, f) != (size_t)n) exit(4);
    fclose(f);
    *bytes = (size_t)n;
    return p;
}
static void dump(const char *name, const void *p, size_t bytes)
{
    char path[1024];
    snprintf(path, sizeof path, "%s/%s", dir, name);
    FILE *f = fopen(path, "wb");
    if (!f || fwrite(p, 1, bytes, f) != bytes) { fprintf(stderr, "cannot write %s\n", path); exit(4); }
    fclose(f);
}
static SEXP real_matrix(const char *name, int nrow, int ncol)       /* column-major doubles */
{
    size_t b;
    double *v = (double *)slurp(name, &b);
    if (b != (size_t)nrow * (ncol < 0 ? 1 : ncol) * 8) { fprintf(stderr, "%s: %zu bytes, expected %d x %d doubles\n", name, b, nrow, ncol); exit(4); }
    SEXP s = ncol < 0 ? Rf_allocVector(REALSXP, nrow) : Rf_allocMatrix(REALSXP, nrow, ncol);
    memcpy(REAL(s), v, b);
    free(v);
    return s;
}
static SEXP int_array(const char *name, int nrow, int ncol)
{
    size_t b;
    int *v = (int *)slurp(name, &b);
    long n = (long)nrow * (ncol < 0 ? 1 : ncol);
    if (b != (size_t)n * 4) { fprintf(stderr, "%s: %zu bytes, expected %ld ints\n", name, b, n); exit(4); }
    SEXP s = ncol < 0 ? Rf_allocVector(INTSXP, nrow) : Rf_allocMatrix(INTSXP, nrow, ncol);
    memcpy(INTEGER(s), v, b);
    free(v);
    return s;
}
static DL_FUNC routine(const char *name, int nargs)
{
    for (const R_CallMethodDef *c = dll.calls; c && c->name; ++c)
        if (strcmp(c->name, name) == 0) {
            if (c->numArgs != nargs) { fprintf(stderr, "%s registered with %d arguments, called with %d\n", name, c->numArgs, nargs); exit(5); }
            return c->fun;
        }
    fprintf(stderr, ".Call routine %s is not registered\n", name);
    exit(5);
}
typedef SEXP (*F1)(SEXP);
typedef SEXP (*F4)(SEXP, SEXP, SEXP, SEXP);
typedef SEXP (*F6)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
typedef SEXP (*F8)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
typedef SEXP (*F10)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);

int main(int argc, char **argv)
{
    if (argc < 2) { fprintf(stderr, "usage: shim_driver <dir>\n"); return 1; }
    dir = argv[1];
    size_t b;
    int *dims = (int *)slurp("dims.bin", &b);      /* ntime ncell nreg nsets warmup which nrow ncol rt_ntime rt_nsub npoly flags
                                                        rt2_nrow rt2_ncol rt2_npoly gauge */
    const int ntime = dims[0], ncell = dims[1], nreg = dims[2], nsets = dims[3], warmup = dims[4], which = dims[5];
    const int nrow = dims[6], ncol = dims[7], rt_ntime = dims[8], rt_nsub = dims[9], npoly = dims[10], flags = dims[11];
    R_init_GR2MSemiDistrB200(&dll);
    if (dll.dynamic != FALSE) { fprintf(stderr, "R_useDynamicSymbols(dll, FALSE) was not called\n"); return 5; }

    SEXP P = real_matrix("P.bin", ntime, ncell), E = real_matrix("E.bin", ntime, ncell);
    SEXP area = real_matrix("area.bin", ncell, -1), ndays = int_array("ndays.bin", ntime, -1);
    SEXP region = int_array("region.bin", ncell, -1), par0 = real_matrix("par0.bin", 4 * nreg, -1);

    /* Run_GR2MSemiDistr's .Call (r_package/R/gr2m_b200.R) */
    SEXP o = ((F10)routine("C_gr2m_run", 10))(P, E, area, ndays, region, mock_scalar_int(nreg), par0, R_NilValue, R_NilValue,
                                               mock_scalar_int(flags));
    const char *want[] = {"PR", "AE", "SM", "RU", "QS", "S_end", "R_end"};
    for (int i = 0; i < 7; ++i) {
        if (strcmp(o->names[i], want[i]) != 0) { fprintf(stderr, "list element %d is named %s\n", i, o->names[i]); return 6; }
        SEXP v = VECTOR_ELT(o, i);
        if (i < 5 && (Rf_nrows(v) != ntime || Rf_ncols(v) != ncell)) { fprintf(stderr, "%s is %d x %d\n", want[i], Rf_nrows(v), Rf_ncols(v)); return 6; }
        char fn[64];
        snprintf(fn, sizeof fn, "out_%s.bin", want[i]);
        dump(fn, REAL(v), (size_t)v->length * 8);
    }
    /* second run from the end state (IniState hand-off) */
    SEXP o2 = ((F10)routine("C_gr2m_run", 10))(P, E, area, ndays, region, mock_scalar_int(nreg), par0, VECTOR_ELT(o, 5),
                                                VECTOR_ELT(o, 6), mock_scalar_int(flags));
    dump("out_QS_restart.bin", REAL(VECTOR_ELT(o2, 4)), (size_t)ntime * ncell * 8);

    /* Optim_GR2MSemiDistr: handle + batched OFUN, one candidate per COLUMN */
    SEXP h = ((F6)routine("C_gr2m_basin", 6))(P, E, area, ndays, region, mock_scalar_int(nreg));
    SEXP params = real_matrix("params.bin", 4 * nreg, nsets), qobs = real_matrix("qobs.bin", ntime, -1);
    SEXP of = ((F6)routine("C_gr2m_objective", 6))(h, params, qobs, mock_scalar_int(warmup), mock_scalar_int(which), mock_scalar_int(flags));
    dump("out_of.bin", REAL(of), (size_t)nsets * 8);
    /* Optim_GR2MSemiDistr(Gauge = ...): routing operator on the run's own subbasins, objective on the routed discharge */
    {
        SEXP fd2 = int_array("rt2_flowdir.bin", dims[12], dims[13]), src2 = int_array("rt2_src.bin", ncell, -1);
        SEXP off2 = int_array("rt2_off.bin", ncell + 1, -1), cells2 = int_array("rt2_cells.bin", dims[14] > 0 ? dims[14] : 1, -1);
        SEXP rh = ((F4)routine("C_wfa_router", 4))(fd2, src2, off2, cells2);
        SEXP ofr = ((F8)routine("C_gr2m_objective_routed", 8))(h, rh, params, qobs, mock_scalar_int(dims[15]), mock_scalar_int(warmup),
                                                               mock_scalar_int(which), mock_scalar_int(flags));
        dump("out_of_routed.bin", REAL(ofr), (size_t)nsets * 8);
        mock_run_finalizers();                               /* gc(): the finalizers destroy the basin, the operator and the plan */
        if (R_ExternalPtrAddr(h) != NULL || R_ExternalPtrAddr(rh) != NULL) { fprintf(stderr, "finalizer did not clear a handle\n"); return 6; }
    }

    /* Routing_GR2MSemiDistr: DEM -> directions, then the routing seam; matrices as raster::as.matrix gives them */
    SEXP dem = real_matrix("dem.bin", nrow, ncol);
    SEXP fd = ((F1)routine("C_wfa_flowdir", 1))(dem);
    dump("out_flowdir.bin", INTEGER(fd), (size_t)nrow * ncol * 4);
    SEXP QS = real_matrix("rt_QS.bin", rt_ntime, rt_nsub), src = int_array("rt_src.bin", rt_nsub, -1);
    SEXP off = int_array("rt_off.bin", rt_nsub + 1, -1), cells = int_array("rt_cells.bin", npoly > 0 ? npoly : 1, -1);
    SEXP QR = ((F6)routine("C_wfa_route", 6))(fd, QS, src, off, cells, mock_scalar_int(0));
    if (Rf_nrows(QR) != rt_ntime || Rf_ncols(QR) != rt_nsub) { fprintf(stderr, "QR is %d x %d\n", Rf_nrows(QR), Rf_ncols(QR)); return 6; }
    dump("out_QR.bin", REAL(QR), (size_t)rt_ntime * rt_nsub * 8);
    printf("shim driver ok\n");
    return 0;
}
