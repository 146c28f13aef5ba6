/*
 * oracle/gr2m_oracle.c -- CPU ORACLE for the GR2MSemiDistr hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under gr2msemidistr_b200/ may include, link, load or call
 * this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs do.  The product path is the CUDA library (gr2msemidistr_b200/csrc) and has no CPU fallback.
 *
 * PARITY UNPINNED.  The reference (hllauca/GR2MSemiDistr, pure R) ships no tests, no golden
 * vectors and none of the arithmetic: the hot path lives in un-vendored, un-pinned third-party
 * code that cannot run in this image (no R, gfortran, MPI, TauDEM):
 *     airGR   (CRAN, unpinned; DESCRIPTION:13)  src/frun_GR2M.f90 (MOD_GR2M), R/RunModel_GR2M.R,
 *                                               R/CreateRunOptions.R (IniResLevels default)
 *     R >= 4.0 (DESCRIPTION:11)                 src/nmath/fround.c        (round(x, digits))
 *     hydroGOF (CRAN, unpinned; DESCRIPTION:15) R/KGE.R, R/NSE.R, R/rmse.R
 *     TauDEM 5.3.7 (README.md:12)               src/aread8.cpp, src/commonLib.cpp (d1/d2)
 * Each function below restates the published algorithm of that dependency and cites the reference
 * call site (file:line relative to /root/reference) whose contract it follows.  It is pinned only
 * by analytic known-answer tests (tests/test_oracle_*.py) and by a NumPy twin (oracle/twin.py).
 *
 * Build: make -C oracle      (gcc -O2, no -ffast-math, -ffp-contract=off)
 */
#include <math.h>
#include <float.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

#define ORC_API __attribute__((visibility("default")))

/* flags shared with include/gr2m_b200.h (same numeric values; tests assert that) */
#define ORC_PERC_F32_EXPONENT 1  /* airGR literal (1./3.) is a default-REAL constant under gfortran */
#define ORC_TANH_APPENDIX_A   2  /* (e^{2w}-1)/(e^{2w}+1) instead of airGR's (e - 1/e)/(e + 1/e) */
#define ORC_TANH_LIBM         4  /* libm tanh() */
#define ORC_SINGLE_UNROUNDED  8  /* internal */

/* ------------------------------------------------------------------------------------------ */
/* R >= 4.0.0  round(x, digits)   [UPSTREAM R src/nmath/fround.c]                              */
/* call sites: R/Run_GR2MSemiDistr.R:101-102,195-199,231-236,255; R/Optim_GR2MSemiDistr.R:103- */
/* 104,195,210-212,232-236; R/Routing_GR2MSemiDistr.R:126,131                                  */
/* ------------------------------------------------------------------------------------------ */
static double r_round_pos(double x, int dig)
{
    /* x > 0 finite.  "Round to even at the last digit" on the two representable candidates
       xd = floor(x*10^dig)/10^dig and xu = ceil(x*10^dig)/10^dig; pick the closer one in double
       arithmetic, the even one on a tie.  If x cannot have more than 15 significant digits at
       `dig` decimals it is returned unchanged. */
    double l10 = log10(x);
    int e10 = (int)floor(l10) + 1;               /* digits before the decimal point (may be <= 0) */
    if (dig + e10 > DBL_DIG) return x;
    double p10 = 1.0, x10, i10, xd, xu;
    int ad = dig < 0 ? -dig : dig;
    for (int i = 0; i < ad; ++i) p10 *= 10.0;    /* exact for |dig| <= 22 */
    if (dig >= 0) {
        x10 = p10 * x;  i10 = floor(x10);
        xd = i10 / p10;  xu = ceil(x10) / p10;
    } else {
        x10 = x / p10;  i10 = floor(x10);
        xd = i10 * p10;  xu = ceil(x10) * p10;
    }
    double du = xu - x, dd = x - xd;
    int even = fmod(i10, 2.0) == 0.0;
    return (dd < du || (dd == du && even)) ? xd : xu;
}

ORC_API double orc_r_round(double x, int digits)
{
    if (isnan(x)) return x;
    if (!isfinite(x) || x == 0.0) return x;
    if (digits == 0) return nearbyint(x);        /* private_rint: half-to-even */
    return x < 0.0 ? -r_round_pos(-x, digits) : r_round_pos(x, digits);
}

ORC_API void orc_r_round_vec(const double *x, int n, int digits, double *out)
{
    for (int i = 0; i < n; ++i) out[i] = orc_r_round(x[i], digits);
}

/* ------------------------------------------------------------------------------------------ */
/* airGR GR2M, one catchment, one parameter pair  [UPSTREAM airGR frun_GR2M.f90 / RunModel_GR2M.R]
 * call sites: R/Run_GR2MSemiDistr.R:145-184 ; R/Optim_GR2MSemiDistr.R:160-177
 * contract:   Precip/PotEvap already factor-corrected and rounded (Run:101-102); IndPeriod_Run =
 *             1:ntime so there is no warm-up run (Run:156); outputs used: Precip, AE, Prod, Qsim,
 *             StateEnd (Run:224-228).
 * ------------------------------------------------------------------------------------------ */
static inline double orc_tanh(double w, int flags)
{
    if (flags & ORC_TANH_LIBM) return tanh(w);
    if (flags & ORC_TANH_APPENDIX_A) { double e = exp(2.0 * w); return (e - 1.0) / (e + 1.0); }
    double e = exp(w);                           /* airGR tanHyp(): ValExp form */
    return (e - 1.0 / e) / (e + 1.0 / e);
}

ORC_API void orc_gr2m_cell(int ntime, const double *P, const double *E, double X1, double X2,
                           int has_ini, double S0, double R0, int flags,
                           double *AE, double *Prod, double *Qsim, double *state_end)
{
    /* RunModel_GR2M.R: parameters floored at 1e-2; default IniResLevels = c(0.3, 0.5) applied as
       Prod = 0.3*X1, Rout = 0.5*X2 (sic: the routing level is scaled by X2). */
    if (X1 < 0.01) X1 = 0.01;
    if (X2 < 0.01) X2 = 0.01;
    double S = has_ini ? S0 : 0.3 * X1;
    double R = has_ini ? R0 : 0.5 * X2;
    const double pexp = (flags & ORC_PERC_F32_EXPONENT) ? (double)(1.0f / 3.0f) : 1.0 / 3.0;
    for (int t = 0; t < ntime; ++t) {
        double WS = P[t] / X1;
        if (WS > 13.0) WS = 13.0;
        double TWS = orc_tanh(WS, flags);
        double Sr = S / X1;
        double S1 = (S + X1 * TWS) / (1.0 + Sr * TWS);
        double P1 = P[t] + S - S1;
        WS = E[t] / X1;
        if (WS > 13.0) WS = 13.0;
        TWS = orc_tanh(WS, flags);
        double S2 = S1 * (1.0 - TWS) / (1.0 + (1.0 - S1 / X1) * TWS);
        double ae = S1 - S2;
        Sr = S2 / X1;
        Sr = Sr * Sr * Sr + 1.0;
        S = S2 / pow(Sr, pexp);
        double P2 = S2 - S;
        double P3 = P1 + P2;
        double R1 = R + P3;
        double R2 = X2 * R1;
        double Q = R2 * R2 / (R2 + 60.0);
        R = R2 - Q;
        if (AE) AE[t] = ae;
        if (Prod) Prod[t] = S;
        if (Qsim) Qsim[t] = Q;
    }
    if (state_end) { state_end[0] = S; state_end[1] = R; }
}

/* ------------------------------------------------------------------------------------------ */
/* Run_GR2MSemiDistr simulation core (a1..a5 of SURVEY section 8): parameter table lookup
 * (Run:82-95), forcing correction round(f*x,1) (Run:96-107), airGR chain (Run:133-187),
 * mm -> m3/s  QS = area*Qsim/(86.4*nDays) (Run:229).  All outputs UNROUNDED, column-major
 * [ntime x ncell] like the R matrices.  params = [X1 x nreg | X2 x nreg | Fpp x nreg | Fpe x nreg].
 * ------------------------------------------------------------------------------------------ */
typedef struct {
    int ntime, ncell, nreg, flags;
    const double *P, *E, *area, *params, *S0, *R0;
    const int *ndays, *region;
    double *PR, *AE, *SM, *RU, *QS, *S_end, *R_end;
    int c0, c1;
} run_job;

static void run_cells(const run_job *j)
{
    int nt = j->ntime;
    double *p = malloc(sizeof(double) * nt), *e = malloc(sizeof(double) * nt);
    double *ae = malloc(sizeof(double) * nt), *sm = malloc(sizeof(double) * nt), *q = malloc(sizeof(double) * nt);
    for (int c = j->c0; c < j->c1; ++c) {
        int r = j->region[c];
        double X1 = j->params[r], X2 = j->params[j->nreg + r];
        double fp = j->params[2 * j->nreg + r], fe = j->params[3 * j->nreg + r];
        for (int t = 0; t < nt; ++t) {
            p[t] = orc_r_round(fp * j->P[(size_t)c * nt + t], 1);
            e[t] = orc_r_round(fe * j->E[(size_t)c * nt + t], 1);
        }
        double st[2];
        int has_ini = j->S0 != NULL && j->R0 != NULL;
        orc_gr2m_cell(nt, p, e, X1, X2, has_ini, has_ini ? j->S0[c] : 0.0, has_ini ? j->R0[c] : 0.0,
                      j->flags, ae, sm, q, st);
        for (int t = 0; t < nt; ++t) {
            size_t k = (size_t)c * nt + t;
            if (j->PR) j->PR[k] = p[t];
            if (j->AE) j->AE[k] = ae[t];
            if (j->SM) j->SM[k] = sm[t];
            if (j->RU) j->RU[k] = q[t];
            if (j->QS) j->QS[k] = (j->area[c] * q[t]) / (86.4 * (double)j->ndays[t]);
        }
        if (j->S_end) j->S_end[c] = st[0];
        if (j->R_end) j->R_end[c] = st[1];
    }
    free(p); free(e); free(ae); free(sm); free(q);
}

static void *run_thread(void *a) { run_cells((const run_job *)a); return NULL; }

ORC_API int orc_run(int ntime, int ncell, const double *P, const double *E, const double *area,
                    const int *ndays, const int *region, int nreg, const double *params,
                    const double *S0, const double *R0, int flags, int nthreads,
                    double *PR, double *AE, double *SM, double *RU, double *QS,
                    double *S_end, double *R_end)
{
    if (ntime <= 0 || ncell <= 0 || nreg <= 0) return -1;
    for (int c = 0; c < ncell; ++c) if (region[c] < 0 || region[c] >= nreg) return -2;
    if (nthreads < 1) nthreads = 1;
    if (nthreads > ncell) nthreads = ncell;
    run_job *jobs = calloc(nthreads, sizeof(run_job));
    pthread_t *th = calloc(nthreads, sizeof(pthread_t));
    for (int i = 0; i < nthreads; ++i) {
        run_job j = { ntime, ncell, nreg, flags, P, E, area, params, S0, R0, ndays, region,
                      PR, AE, SM, RU, QS, S_end, R_end,
                      (int)((long long)ncell * i / nthreads), (int)((long long)ncell * (i + 1) / nthreads) };
        jobs[i] = j;
        if (nthreads == 1) run_cells(&jobs[i]);
        else pthread_create(&th[i], NULL, run_thread, &jobs[i]);
    }
    if (nthreads > 1) for (int i = 0; i < nthreads; ++i) pthread_join(th[i], NULL);
    free(jobs); free(th);
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* Outlet aggregation -- the two variants the reference has.
 *   Run   (Run:229-236):   qs = round(QS,1);  qt = round(rowSums(qs),2)
 *          nsub == 1 (Run:195-200): qs = round(area*round(ru,1)/(86.4*nDays),1); qt = qs
 *   Optim (Optim:186-196): sink = round(rowSums(QS_unrounded),2);   nsub == 1: sink = QS (unrounded)
 * R's sum() accumulates in long double, left to right.
 * ------------------------------------------------------------------------------------------ */
ORC_API void orc_outlet_run(int ntime, int ncell, const double *QS, const double *RU,
                            const double *area, const int *ndays, double *qt)
{
    if (ncell == 1) {
        for (int t = 0; t < ntime; ++t)
            qt[t] = orc_r_round((area[0] * orc_r_round(RU[t], 1)) / (86.4 * (double)ndays[t]), 1);
        return;
    }
    for (int t = 0; t < ntime; ++t) {
        long double s = 0.0L;
        for (int c = 0; c < ncell; ++c) s += orc_r_round(QS[(size_t)c * ntime + t], 1);
        qt[t] = orc_r_round((double)s, 2);
    }
}

ORC_API void orc_outlet_optim(int ntime, int ncell, const double *QS, double *sink, int unrounded)
{
    for (int t = 0; t < ntime; ++t) {
        if (ncell == 1) { sink[t] = QS[t]; continue; }
        long double s = 0.0L;
        for (int c = 0; c < ncell; ++c) s += QS[(size_t)c * ntime + t];
        sink[t] = unrounded ? (double)s : orc_r_round((double)s, 2);
    }
}

/* ------------------------------------------------------------------------------------------ */
/* hydroGOF KGE (method "2009", s = c(1,1,1)), NSE, rmse   [UPSTREAM hydroGOF R/KGE.R, NSE.R, rmse.R]
 * call site: R/Optim_GR2MSemiDistr.R:199-216.  Warm-up rows dropped (Optim:199-205), NA pairs
 * dropped (na.omit, Optim:209).  R's mean()/sd()/cor() are two-pass in long double.
 * crit[0..2] = KGE, NSE, RMSE (raw);  of[0..2] = 1-round(KGE,3), 1-round(NSE,3), round(RMSE,3).
 * Returns the number of valid pairs.
 * ------------------------------------------------------------------------------------------ */
static long double r_mean(const double *x, int n)
{
    long double s = 0.0L;
    for (int i = 0; i < n; ++i) s += x[i];
    s /= n;
    long double t = 0.0L;
    for (int i = 0; i < n; ++i) t += (x[i] - s);
    return s + t / n;
}

ORC_API int orc_gof(int n_in, const double *sim_in, const double *obs_in, int warmup,
                    double *crit, double *of)
{
    int n = 0;
    double *sim = malloc(sizeof(double) * (n_in > 0 ? n_in : 1)), *obs = malloc(sizeof(double) * (n_in > 0 ? n_in : 1));
    for (int t = warmup > 0 ? warmup : 0; t < n_in; ++t)
        if (!isnan(sim_in[t]) && !isnan(obs_in[t])) { sim[n] = sim_in[t]; obs[n] = obs_in[t]; ++n; }
    double kge = NAN, nse = NAN, rmse = NAN;
    if (n > 0) {
        long double ms = r_mean(sim, n), mo = r_mean(obs, n);
        double msd = (double)ms, mod = (double)mo;     /* hydroGOF works on R doubles */
        long double sse = 0.0L, den = 0.0L, vs = 0.0L, vo = 0.0L, cv = 0.0L;
        for (int i = 0; i < n; ++i) {
            double d = obs[i] - sim[i];      sse += (long double)(d * d);
            double a = obs[i] - mod;         den += (long double)(a * a);
        }
        /* NSE.R: denominator <- sum((obs - mean(obs))^2); NS <- 1 - sum((obs-sim)^2)/denominator */
        if ((double)den != 0.0) nse = 1.0 - ((double)sse / (double)den);
        /* rmse.R: sqrt(mean((sim-obs)^2)) -- mean() is two-pass */
        {
            double *sq = malloc(sizeof(double) * n);
            for (int i = 0; i < n; ++i) { double d = sim[i] - obs[i]; sq[i] = d * d; }
            rmse = sqrt((double)r_mean(sq, n));
            free(sq);
        }
        /* KGE.R: sd() = sqrt(var()), cor(method="pearson") -- R cov.c two-pass, n-1 denominators */
        if (n > 1) {
            for (int i = 0; i < n; ++i) {
                long double a = sim[i] - ms, b = obs[i] - mo;
                vs += a * a; vo += b * b; cv += a * b;
            }
            double sd_s = sqrt((double)(vs / (n - 1))), sd_o = sqrt((double)(vo / (n - 1)));
            long double rr = cv / (sqrtl(vs) * sqrtl(vo));
            if (rr > 1.0L) rr = 1.0L;
            if (rr < -1.0L) rr = -1.0L;
            double r = (double)rr;
            double alpha = sd_s / sd_o, beta = msd / mod;
            if (mod != 0.0 || sd_o != 0.0) {
                double a = r - 1.0, b = alpha - 1.0, c = beta - 1.0;
                kge = 1.0 - sqrt(a * a + b * b + c * c);
            }
        }
    }
    if (crit) { crit[0] = kge; crit[1] = nse; crit[2] = rmse; }
    if (of) {
        of[0] = 1.0 - orc_r_round(kge, 3);
        of[1] = 1.0 - orc_r_round(nse, 3);
        of[2] = orc_r_round(rmse, 3);
    }
    free(sim); free(obs);
    return n;
}

/* ------------------------------------------------------------------------------------------ */
/* OFUN(parameter_set) of Optim_GR2MSemiDistr (Optim:125-217), batched over parameter sets.
 * params is [nsets x 4*nreg] row-major (one reference-ordered parameter vector per set).
 * Threads split the parameter sets (this is also the timed CPU baseline of bench.py).
 * out_crit/out_of: [nsets x 3];  out_qt (nullable): [ntime x nsets] column-major sink series.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
    int ntime, ncell, nreg, flags, warmup, s0, s1;
    const double *P, *E, *area, *params, *qobs;
    const int *ndays, *region;
    double *crit, *of, *qt;
} obj_job;

static void obj_sets(const obj_job *j)
{
    int nt = j->ntime;
    double *p = malloc(sizeof(double) * nt), *e = malloc(sizeof(double) * nt), *q = malloc(sizeof(double) * nt);
    long double *acc = malloc(sizeof(long double) * nt);
    double *sink = malloc(sizeof(double) * nt);
    for (int s = j->s0; s < j->s1; ++s) {
        const double *par = j->params + (size_t)s * 4 * j->nreg;
        for (int t = 0; t < nt; ++t) acc[t] = 0.0L;
        for (int c = 0; c < j->ncell; ++c) {
            int r = j->region[c];
            double fp = par[2 * j->nreg + r], fe = par[3 * j->nreg + r];
            for (int t = 0; t < nt; ++t) {
                p[t] = orc_r_round(fp * j->P[(size_t)c * nt + t], 1);
                e[t] = orc_r_round(fe * j->E[(size_t)c * nt + t], 1);
            }
            orc_gr2m_cell(nt, p, e, par[r], par[j->nreg + r], 0, 0.0, 0.0, j->flags, NULL, NULL, q, NULL);
            for (int t = 0; t < nt; ++t) {
                double qs = (j->area[c] * q[t]) / (86.4 * (double)j->ndays[t]);
                if (j->ncell == 1) sink[t] = qs; else acc[t] += qs;
            }
        }
        if (j->ncell > 1)
            for (int t = 0; t < nt; ++t)
                sink[t] = (j->flags & ORC_SINGLE_UNROUNDED) ? (double)acc[t] : orc_r_round((double)acc[t], 2);
        if (j->qt) memcpy(j->qt + (size_t)s * nt, sink, sizeof(double) * nt);
        if (j->qobs)
            orc_gof(nt, sink, j->qobs, j->warmup, j->crit ? j->crit + 3 * (size_t)s : NULL,
                    j->of ? j->of + 3 * (size_t)s : NULL);
    }
    free(p); free(e); free(q); free(acc); free(sink);
}

static void *obj_thread(void *a) { obj_sets((const obj_job *)a); return NULL; }

ORC_API int orc_objective_batch(int ntime, int ncell, const double *P, const double *E,
                                const double *area, const int *ndays, const int *region, int nreg,
                                const double *params, int nsets, const double *qobs, int warmup,
                                int flags, int nthreads, double *out_crit, double *out_of, double *out_qt)
{
    if (ntime <= 0 || ncell <= 0 || nreg <= 0 || nsets <= 0) return -1;
    for (int c = 0; c < ncell; ++c) if (region[c] < 0 || region[c] >= nreg) return -2;
    if (nthreads < 1) nthreads = 1;
    if (nthreads > nsets) nthreads = nsets;
    obj_job *jobs = calloc(nthreads, sizeof(obj_job));
    pthread_t *th = calloc(nthreads, sizeof(pthread_t));
    for (int i = 0; i < nthreads; ++i) {
        obj_job j = { ntime, ncell, nreg, flags, warmup,
                      (int)((long long)nsets * i / nthreads), (int)((long long)nsets * (i + 1) / nthreads),
                      P, E, area, params, qobs, ndays, region, out_crit, out_of, out_qt };
        jobs[i] = j;
        if (nthreads == 1) obj_sets(&jobs[i]);
        else pthread_create(&th[i], NULL, obj_thread, &jobs[i]);
    }
    if (nthreads > 1) for (int i = 0; i < nthreads; ++i) pthread_join(th[i], NULL);
    free(jobs); free(th);
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* TauDEM AreaD8 with a weight grid  [UPSTREAM TauDEM 5.3.7 src/aread8.cpp, commonLib.cpp d1/d2]
 * call site: R/Routing_GR2MSemiDistr.R:117  ("AreaD8 -p FlowDir.tif -wg Weights.tif -ad8 ...",
 * no -nc  =>  edge-contamination check ON).
 * Flow codes 1=E 2=NE 3=N 4=NW 5=W 6=SW 7=S 8=SE; neighbour k of c drains into c iff
 * |dir[nbr_k] - k| == 4.  Cells are evaluated when all their contributing neighbours are done
 * (FIFO queue); on evaluation A[c] = W[c] then += A[nbr_k] for k = 1..8 in that order.
 * ------------------------------------------------------------------------------------------ */
static const int DX[9] = { 0, 1, 1, 0, -1, -1, -1, 0, 1 };   /* column offset */
static const int DY[9] = { 0, 0, -1, -1, -1, 0, 1, 1, 1 };   /* row offset    */

static inline int d8_valid(int16_t d) { return d >= 1 && d <= 8; }

static int cmp_i32(const void *a, const void *b)
{
    int32_t x = *(const int32_t *)a, y = *(const int32_t *)b;
    return (x > y) - (x < y);
}

/* Topology plan (month-invariant).  Arrays sized ncell = nrow*ncol unless noted.
 *   receiver[c]  downstream cell id, -1 if c is nodata, or drains off-grid / into nodata
 *   inmask[c]    bit (k-1) set iff neighbour k is valid and drains into c (0 for nodata cells)
 *   indeg[c]     popcount(inmask[c])
 *   level[c]     0 for in-degree-0 valid cells, else 1+max(level[upstream]); -1 nodata or unreached
 *   order[nvalid_reached]  cells sorted by (level, cell id); level_off[nlevels+1] offsets into order
 *   contam[c]    1 iff edge-contaminated under TauDEM's rule (any of the 8 neighbours off-grid or
 *                nodata, or any draining neighbour contaminated); 0 otherwise / nodata
 * Returns nlevels (>= 0) or a negative error.  *n_order receives the number of entries in order.
 */
ORC_API int orc_d8_plan(int nrow, int ncol, const int16_t *dir, int32_t *receiver, uint8_t *inmask,
                        int32_t *indeg, int32_t *level, int32_t *order, int32_t *level_off,
                        uint8_t *contam, int32_t *n_order)
{
    if (nrow <= 0 || ncol <= 0) return -1;
    size_t n = (size_t)nrow * ncol;
    if (n > 0x7fffffff) return -2;
    int32_t *cnt = malloc(sizeof(int32_t) * n);
    uint8_t *edge = malloc(n);
    for (int r = 0; r < nrow; ++r) for (int c = 0; c < ncol; ++c) {
        size_t i = (size_t)r * ncol + c;
        level[i] = -1; receiver[i] = -1; inmask[i] = 0; indeg[i] = 0; contam[i] = 0; edge[i] = 0;
        if (!d8_valid(dir[i])) { cnt[i] = -1; continue; }
        int rr = r + DY[dir[i]], cc = c + DX[dir[i]];
        if (rr >= 0 && rr < nrow && cc >= 0 && cc < ncol && d8_valid(dir[(size_t)rr * ncol + cc]))
            receiver[i] = (int32_t)((size_t)rr * ncol + cc);
        uint8_t m = 0; int e = 0;
        for (int k = 1; k <= 8; ++k) {
            int r2 = r + DY[k], c2 = c + DX[k];
            if (r2 < 0 || r2 >= nrow || c2 < 0 || c2 >= ncol) { e = 1; continue; }
            int16_t d = dir[(size_t)r2 * ncol + c2];
            if (!d8_valid(d)) { e = 1; continue; }
            if (d - k == 4 || d - k == -4) m |= (uint8_t)(1u << (k - 1));
        }
        inmask[i] = m; indeg[i] = __builtin_popcount(m); cnt[i] = indeg[i]; edge[i] = (uint8_t)e;
    }
    /* Kahn peeling by whole frontiers => level = frontier index; scanning cells in id order inside
       a frontier gives the (level, cell id) order directly after a stable bucket pass. */
    int32_t *cur = malloc(sizeof(int32_t) * n), *nxt = malloc(sizeof(int32_t) * n);
    size_t ncur = 0, nord = 0;
    for (size_t i = 0; i < n; ++i) if (cnt[i] == 0) cur[ncur++] = (int32_t)i;
    int lev = 0;
    level_off[0] = 0;
    while (ncur > 0) {
        size_t nn = 0;
        for (size_t q = 0; q < ncur; ++q) {
            int32_t i = cur[q];
            level[i] = lev;
            int r = i / ncol, c = i % ncol;
            uint8_t con = edge[i];
            for (int k = 1; k <= 8; ++k)
                if (inmask[i] & (1u << (k - 1))) con |= contam[(size_t)(r + DY[k]) * ncol + (c + DX[k])];
            contam[i] = con;
            int32_t d = receiver[i];
            if (d >= 0 && --cnt[d] == 0) nxt[nn++] = d;
        }
        /* emit this level in ascending cell id */
        int32_t *dst = order + nord;
        memcpy(dst, cur, sizeof(int32_t) * ncur);
        /* frontier built from an id-ordered scan is sorted for level 0; later ones need a sort */
        if (lev > 0) {
            qsort(dst, ncur, sizeof(int32_t), cmp_i32);
        }
        nord += ncur;
        ++lev;
        level_off[lev] = (int32_t)nord;
        int32_t *tmp = cur; cur = nxt; nxt = tmp; ncur = nn;
    }
    if (n_order) *n_order = (int32_t)nord;
    free(cnt); free(edge); free(cur); free(nxt);
    return lev;
}

/* AreaD8 proper, FIFO (Kahn) order as in aread8.cpp.  W is [ncell] for one month.
 * f32 != 0: TauDEM-faithful -- weights quantised to float32 (raster FLT4S write, Routing:114) and
 * accumulated in float32.  contcheck != 0: edge contamination ON (reference default).
 * Output A[ncell] as double; nodata (TauDEM -1 => NA after raster()) is returned as NaN. */
ORC_API int orc_aread8(int nrow, int ncol, const int16_t *dir, const double *W, int f32,
                       int contcheck, double *A)
{
    size_t n = (size_t)nrow * ncol;
    int16_t *cnt = malloc(sizeof(int16_t) * n);
    uint8_t *nod = malloc(n);                 /* 1 = result is nodata */
    int32_t *que = malloc(sizeof(int32_t) * n);
    float *Af = f32 ? malloc(sizeof(float) * n) : NULL;
    size_t qh = 0, qt = 0;
    for (size_t i = 0; i < n; ++i) { nod[i] = 1; A[i] = NAN; }
    for (int r = 0; r < nrow; ++r) for (int c = 0; c < ncol; ++c) {
        size_t i = (size_t)r * ncol + c;
        cnt[i] = -1;
        if (!d8_valid(dir[i])) continue;
        cnt[i] = 0;
        for (int k = 1; k <= 8; ++k) {
            int r2 = r + DY[k], c2 = c + DX[k];
            if (r2 < 0 || r2 >= nrow || c2 < 0 || c2 >= ncol) continue;
            int16_t d = dir[(size_t)r2 * ncol + c2];
            if (d8_valid(d) && (d - k == 4 || d - k == -4)) cnt[i]++;
        }
        if (cnt[i] == 0) que[qt++] = (int32_t)i;
    }
    while (qh < qt) {
        int32_t i = que[qh++];
        int r = i / ncol, c = i % ncol;
        int con = 0;
        double acc = W[i];
        float accf = (float)W[i];
        for (int k = 1; k <= 8; ++k) {
            int r2 = r + DY[k], c2 = c + DX[k];
            int inside = !(r2 < 0 || r2 >= nrow || c2 < 0 || c2 >= ncol);
            size_t j = inside ? (size_t)r2 * ncol + c2 : 0;
            if (inside && d8_valid(dir[j])) {
                int16_t d = dir[j];
                if (d - k == 4 || d - k == -4) {
                    if (nod[j]) con = 1;
                    else if (f32) accf += Af[j];
                    else acc += A[j];
                }
            } else con = 1;
        }
        if (con && contcheck) { nod[i] = 1; A[i] = NAN; }
        else { nod[i] = 0; if (f32) { Af[i] = accf; A[i] = (double)accf; } else A[i] = acc; }
        int rr = r + DY[dir[i]], cc = c + DX[dir[i]];
        if (rr >= 0 && rr < nrow && cc >= 0 && cc < ncol) {
            size_t d = (size_t)rr * ncol + cc;
            if (cnt[d] > 0 && --cnt[d] == 0) que[qt++] = (int32_t)d;
        }
    }
    free(cnt); free(nod); free(que); free(Af);
    return 0;
}

/* Routing_GR2MSemiDistr month loop (Routing:100-132): per month scatter QS[t,] to the subbasin
 * cells (duplicates: last assignment wins, as R's `[<-`), AreaD8, per-polygon max over the cells
 * with coverage fraction 1, NaN (nodata) ignored (na.rm=TRUE); empty => -Inf.
 * QS, QR: [ntime x nsub] column-major, QR UNROUNDED (Routing:131 rounds to 1 dp in the wrapper).
 * src_cell: 0-based cell ids; poly_off[nsub+1], poly_cells[]: interior cells of each polygon.
 * Threads split the months. */
typedef struct {
    int nrow, ncol, ntime, nsub, f32, contcheck, t0, t1;
    const int16_t *dir; const double *QS; const int32_t *src, *poff, *pcells; double *QR;
} route_job;

static void route_months(const route_job *j)
{
    size_t n = (size_t)j->nrow * j->ncol;
    double *W = malloc(sizeof(double) * n), *A = malloc(sizeof(double) * n);
    for (int t = j->t0; t < j->t1; ++t) {
        memset(W, 0, sizeof(double) * n);
        for (int s = 0; s < j->nsub; ++s) W[j->src[s]] = j->QS[(size_t)s * j->ntime + t];
        orc_aread8(j->nrow, j->ncol, j->dir, W, j->f32, j->contcheck, A);
        for (int s = 0; s < j->nsub; ++s) {
            double m = -INFINITY;
            for (int q = j->poff[s]; q < j->poff[s + 1]; ++q) {
                double v = A[j->pcells[q]];
                if (!isnan(v) && v > m) m = v;
            }
            j->QR[(size_t)s * j->ntime + t] = m;
        }
    }
    free(W); free(A);
}

static void *route_thread(void *a) { route_months((const route_job *)a); return NULL; }

ORC_API int orc_route(int nrow, int ncol, const int16_t *dir, int ntime, int nsub, const double *QS,
                      const int32_t *src_cell, const int32_t *poly_off, const int32_t *poly_cells,
                      int f32, int contcheck, int nthreads, double *QR)
{
    size_t n = (size_t)nrow * ncol;
    for (int s = 0; s < nsub; ++s) if (src_cell[s] < 0 || (size_t)src_cell[s] >= n) return -3;
    if (nthreads < 1) nthreads = 1;
    if (nthreads > ntime) nthreads = ntime;
    route_job *jobs = calloc(nthreads, sizeof(route_job));
    pthread_t *th = calloc(nthreads, sizeof(pthread_t));
    for (int i = 0; i < nthreads; ++i) {
        route_job j = { nrow, ncol, ntime, nsub, f32, contcheck,
                        (int)((long long)ntime * i / nthreads), (int)((long long)ntime * (i + 1) / nthreads),
                        dir, QS, src_cell, poly_off, poly_cells, QR };
        jobs[i] = j;
        if (nthreads == 1) route_months(&jobs[i]);
        else pthread_create(&th[i], NULL, route_thread, &jobs[i]);
    }
    if (nthreads > 1) for (int i = 0; i < nthreads; ++i) pthread_join(th[i], NULL);
    free(jobs); free(th);
    return 0;
}

/* Dense multi-month AreaD8 (config C5): W, A are [nmonth][ncell] (month-major planes). */
ORC_API int orc_aread8_months(int nrow, int ncol, const int16_t *dir, int nmonth, const double *W,
                              int f32, int contcheck, double *A)
{
    size_t n = (size_t)nrow * ncol;
    for (int m = 0; m < nmonth; ++m)
        orc_aread8(nrow, ncol, dir, W + (size_t)m * n, f32, contcheck, A + (size_t)m * n);
    return 0;
}

/* numberOfDays (Run:108-117): days in the month of (year, month). */
ORC_API int orc_days_in_month(int year, int month)
{
    static const int d[12] = { 31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31 };
    if (month == 2 && ((year % 4 == 0 && year % 100 != 0) || year % 400 == 0)) return 29;
    return d[month - 1];
}
