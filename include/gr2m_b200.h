/*
 * gr2m_b200.h -- C ABI of libgr2m_b200.so: the GR2MSemiDistr hot path on NVIDIA B200 (sm_100a).
 *
 * Plain C, plain pointers and sizes, caller-owned host buffers, blocking calls, int status
 * (0 = ok, < 0 = error; text via gr2m_last_error()).  Never aborts, never longjmps, and has NO CPU
 * fallback: without a CUDA device every compute entry point returns GR2M_ERR_NODEVICE.
 *
 * Matrices cross the boundary exactly as R holds them: column-major [ntime x ncell] doubles
 * (element (t, c) at c*ntime + t), so an R numeric matrix / the columns of the Data frame
 * (R/Run_GR2MSemiDistr.R:99: column i+1 = P of subbasin i, column i+1+nsub = E) pass without a
 * copy.  Reference citations are file:line under /root/reference.
 *
 * What each group replaces in the reference:
 *   gr2m_basin_* + gr2m_run      the per-subbasin PSOCK fan-out and airGR chain of
 *                                Run_GR2MSemiDistr (R/Run_GR2MSemiDistr.R:126-189, 223-229),
 *                                i.e. airGR::RunModel_GR2M -> .Fortran("frun_gr2m") per subbasin.
 *   gr2m_objective_batch         the body of OFUN(parameter_set) in Optim_GR2MSemiDistr
 *                                (R/Optim_GR2MSemiDistr.R:125-217), batched over parameter sets.
 *   wfa_plan_* + wfa_route       the month loop of Routing_GR2MSemiDistr
 *                                (R/Routing_GR2MSemiDistr.R:100-123): Weights.tif + "mpiexec -n 8
 *                                AreaD8 -p -wg -ad8" + exact_extract max, once per month.
 *   wfa_accumulate               TauDEM "AreaD8 -p FlowDir.tif -wg Weights.tif -ad8 FlowAcum.tif"
 *                                (R/Routing_GR2MSemiDistr.R:117) on dense weight grids, many months.
 */
#ifndef GR2M_B200_H
#define GR2M_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GR2M_OK             0
#define GR2M_ERR_INVALID   -1   /* bad argument (null pointer, size <= 0, region id out of range ...) */
#define GR2M_ERR_CUDA      -2   /* a CUDA runtime call or kernel failed */
#define GR2M_ERR_NOMEM     -3   /* device or host allocation failed */
#define GR2M_ERR_NODEVICE  -4   /* no usable CUDA device: there is no CPU path */

/* flags for gr2m_run / gr2m_objective_batch */
#define GR2M_PERC_F32_EXPONENT 1   /* percolation exponent = (double)(1.f/3.f), the value gfortran gives
                                      airGR's literal (1./3.); default is the exact cube root */
#define GR2M_SINK_UNROUNDED    8   /* objective: skip round(rowSums(qs),2) of Optim:195 */
#define GR2M_MATH_LITERAL     16   /* evaluate the step with IEEE division / libdevice exp / pow exactly
                                      as written in airGR instead of the fused fast formulation */

/* objective selector, as the `Optimization` argument of Optim_GR2MSemiDistr (Optim:50,215) */
#define GR2M_OF_KGE  0
#define GR2M_OF_NSE  1
#define GR2M_OF_RMSE 2

/* flags for wfa_* */
#define WFA_F32               1   /* TauDEM-faithful: float32 weights (FLT4S write, Routing:114) and
                                     float32 accumulation; default accumulates in FP64 */
#define WFA_NO_CONTAMINATION  2   /* AreaD8 -nc; default keeps the edge-contamination check ON as the
                                     reference does (no -nc at Routing:117) */
#define WFA_LEVEL_SYNC        4   /* sweep the upper levels grid-wide (one launch / one CTA per level) instead of
                                     per outlet basin, which is the default when the forest splits well */
#define WFA_DATAFLOW          8   /* flag-driven dataflow kernel for the thin tail instead */
#define WFA_BY_BASIN          32   /* sweep the upper levels per outlet basin (level by level inside a CTA) instead of per
                                      river reach, the default */
#define WFA_ROUTE_DENSE       64   /* wfa_route: sweep the whole grid month by month (as AreaD8 does) instead of
                                      applying the month-invariant routing operator, the default */
#define WFA_NO_OVERLAP       128   /* launch the wide levels without programmatic stream serialization */
#define WFA_NO_SEGMENTS      256   /* sweep the levels below the cut cell by cell instead of by chain segment (A/B switch) */
#define WFA_NO_FORWARD        16   /* per-basin sweep without shared-memory forwarding of fresh rows (A/B switch) */

typedef struct gr2m_basin gr2m_basin;   /* forcing + static subbasin data resident in HBM */
typedef struct wfa_plan   wfa_plan;     /* D8 topology plan resident in HBM */

/* ---- library ------------------------------------------------------------------------------ */
const char *gr2m_last_error(void);               /* thread-local message of the last failing call */
int  gr2m_version(void);                         /* ABI version, currently 1 */
int  gr2m_device_count(void);                    /* number of CUDA devices, 0 if none */
/* Handles keep up to 8 GB of freed device blocks (up to 4 GB each) for the next handle, because the
 * reference's workflow creates and destroys one per call; this gives them back to the driver. */
int  gr2m_release_cached_memory(void);
int  gr2m_set_device(int device);                /* device used by handles the calling thread creates from now on */
/* GR2M_GUARD=1 in the environment: every handle buffer gets guard bands that are checked when it is released and
 * an 0xFF-filled payload (debug stand-in for a memory checker); these count the buffers checked / found damaged. */
long long gr2m_guard_checked(void);
long long gr2m_guard_violations(void);
long long gr2m_kernel_launches(void);            /* kernels this library has launched since it was loaded (all devices) */

/* Measured FP64 pipe peak of the current device in TFLOP/s (independent DFMA chains, 2 flop each,
 * `iters` x 64 DFMA per thread): the roofline denominator bench.py reports the recursion against. */
int  gr2m_fp64_peak(int iters, double *tflops, float *ms);

/* ---- recursion: Run_GR2MSemiDistr core ------------------------------------------------------
 * gr2m_basin_create uploads the run-period rows of the Data table once:
 *   P, E      [ntime x ncell] column-major, raw (uncorrected) forcing   (Run:78, 99)
 *   area      [ncell] km2  (Subbasins@data$Area, Run:68)
 *   ndays     [ntime] days in each month (numberOfDays, Run:108-120)
 *   region_id [ncell] 0-based index into sort(unique(Region))           (Run:82-88)
 */
int gr2m_basin_create(int ntime, int ncell, const double *P, const double *E, const double *area,
                      const int *ndays, const int *region_id, int nreg, gr2m_basin **out);
/* Same, with P and E already in device memory (same layout); `cuda_stream` is the cudaStream_t they were
 * written on (NULL = complete).  Used when N ranks each upload 1/N of the table and all-gather it. */
int gr2m_basin_create_dev(int ntime, int ncell, const double *P_dev, const double *E_dev, const double *area,
                          const int *ndays, const int *region_id, int nreg, void *cuda_stream, gr2m_basin **out);
int gr2m_basin_destroy(gr2m_basin *b);
/* run all work of this handle on a caller-provided cudaStream_t (NULL = the handle's own stream) */
int gr2m_basin_set_stream(gr2m_basin *b, void *cuda_stream);

/* One simulation with one parameter vector params[4*nreg] = [X1 x nreg | X2 x nreg | Fpp x nreg |
 * Fpe x nreg] (Run:85-88).  S0/R0 [ncell] = IniState Store$Prod / Store$Rout (Run:163-164), both NULL
 * for airGR's defaults 0.3*X1 / 0.5*X2.  Outputs (any may be NULL) are UNROUNDED, [ntime x ncell]
 * column-major: PR = round(Fpp*P,1) (Run:101), AE, SM (= airGR Prod), RU (= Qsim) as gathered at
 * Run:224-228, QS = area*RU/(86.4*nDays) (Run:229); S_end/R_end [ncell] = StateEnd (Run:224).
 * The R wrapper applies round(,1)/round(,2), the warm-up trim and the list assembly (Run:231-277). */
int gr2m_run(gr2m_basin *b, const double *params, const double *S0, const double *R0, int flags,
             double *PR, double *AE, double *SM, double *RU, double *QS, double *S_end, double *R_end);

/* One-shot form of the above (create + run + destroy). */
int gr2m_run_once(int ntime, int ncell, const double *P, const double *E, const double *area,
                  const int *ndays, const int *region_id, int nreg, const double *params,
                  const double *S0, const double *R0, int flags,
                  double *PR, double *AE, double *SM, double *RU, double *QS,
                  double *S_end, double *R_end);

/* ---- objective: OFUN of Optim_GR2MSemiDistr, nsets parameter vectors at once -----------------
 * params   [nsets x 4*nreg] row-major: one reference-ordered vector per candidate set
 * qobs     [ntime] observed discharge, NaN = NA (dropped pairwise, Optim:209); may be NULL if only
 *          out_qt is wanted
 * warmup   number of leading months dropped (Optim:199-205), 0 for WarmUp=NULL
 * which    GR2M_OF_KGE / _NSE / _RMSE
 * out_of   [nsets]   the value OFUN returns: 1-round(KGE,3), 1-round(NSE,3) or round(RMSE,3)
 * out_crit [nsets x 3] row-major raw KGE, NSE, RMSE (unrounded)            (may be NULL)
 * out_qt   [ntime x nsets] column-major outlet series `sink` (Optim:186-196) (may be NULL)
 * With nsets = 1 this is a drop-in body for OFUN, so rtop::sceua (Optim:225-229) works unchanged. */
int gr2m_objective_batch(gr2m_basin *b, const double *params, int nsets, const double *qobs,
                         int warmup, int which, int flags,
                         double *out_of, double *out_crit, double *out_qt);

/* Same, with params/out buffers already in device memory (params_dev, out_*_dev are device
 * pointers; qobs stays a host pointer and is cached on the device).  Asynchronous on the handle's
 * stream; used to keep a population-based optimiser entirely on the GPU. */
int gr2m_objective_batch_dev(gr2m_basin *b, const double *params_dev, int nsets, const double *qobs,
                             int warmup, int which, int flags,
                             double *out_of_dev, double *out_crit_dev, double *out_qt_dev);

/* milliseconds spent in the kernels of the last call on this handle (CUDA events on its stream) */
int gr2m_basin_last_kernel_ms(gr2m_basin *b, float *ms);

/* ---- routing: Weighted Flow Accumulation ----------------------------------------------------
 * wfa_plan_create: flowdir [nrow x ncol] ROW-major int16 in TauDEM D8 codes 1=E 2=NE 3=N 4=NW 5=W
 * 6=SW 7=S 8=SE (FlowDir.tif of Routing:91-93); any other value is nodata.  Computes, on the device
 * and once: receiver, in-degree, Kahn levels, level order and edge contamination. */
int wfa_plan_create(const int16_t *flowdir, int nrow, int ncol, int flags, wfa_plan **out);
int wfa_plan_create_dev(const int16_t *flowdir_dev, int nrow, int ncol, int flags, wfa_plan **out);
int wfa_plan_destroy(wfa_plan *p);
int wfa_plan_set_stream(wfa_plan *p, void *cuda_stream);
/* sizes: number of levels, number of cells in the level order, number of cells */
int wfa_plan_info(const wfa_plan *p, int *nlevels, int *norder, int64_t *ncell);
/* per-outlet-basin regrouping of the upper levels: cut level, number of basins (trees), cells above the
 * cut, and whether the sweep uses it (it needs >= 32 basins, none holding more than 1/8 of those cells) */
int wfa_plan_basin_info(const wfa_plan *p, int *cut_level, int *nbasins, int *nupper, int *enabled);
/* River reaches of the same upper network: maximal chains of cells with one upper inflow each, swept by one lane
 * group per reach, reaches grouped into `nreach_levels` dependency levels (one launch each);
 * `critical_cells` = sum over those levels of the longest reach, the sweep's serial chain in cells. */
int wfa_plan_reach_info(const wfa_plan *p, int *nreaches, int *nreach_levels, int *critical_cells, int *enabled);
/* chain segments of the levels below the cut (csrc/wfa.cu, k_accum_segments): segments, cells they cover, first level
 * not covered, and whether the default sweep uses them */
int wfa_plan_segment_info(const wfa_plan *p, long long *nsegments, long long *ncells, int *levels, int *enabled);
/* Copy the integer topology back (any pointer may be NULL): receiver/indeg/level [ncell] int32,
 * inmask/contam [ncell] uint8, order [norder], level_off [nlevels+1]. */
int wfa_plan_export(const wfa_plan *p, int32_t *receiver, uint8_t *inmask, int32_t *indeg,
                    int32_t *level, int32_t *order, int32_t *level_off, uint8_t *contam);

/* The month loop of Routing_GR2MSemiDistr for all months at once.
 *   QS        [ntime x nsub] column-major (Model$QS, Routing:57-74)
 *   src_cell  [nsub] 0-based row-major DEM cell of each subbasin's point on surface (Routing:79-81)
 *   poly_off  [nsub+1], poly_cells[]: 0-based cells with coverage fraction 1 per polygon (Routing:119-121)
 *   QR        [ntime x nsub] column-major, UNROUNDED per-polygon max (na.rm; -Inf if empty);
 *             the wrapper applies round(,1) (Routing:131). */
int wfa_route(wfa_plan *p, int ntime, int nsub, const double *QS, const int32_t *src_cell,
              const int32_t *poly_off, const int32_t *poly_cells, int flags, double *QR);

/* Month-invariant routing operator (same inputs and result as wfa_route, split in two): everything that
 * depends on the D8 grid, the source cells and the polygons only is computed once; applying it costs
 * O(confluences of source-carrying paths) per column instead of O(grid cells).  QS and QR are
 * [ncols x nsub] column-major; the columns are months, or months x parameter sets alike.  `flags` takes
 * WFA_F32 and WFA_NO_CONTAMINATION.  Results equal the dense sweep's bit for bit (a zero may differ in sign).
 * A router borrows its plan's topology and stream: destroy it before the plan. */
typedef struct wfa_router wfa_router;
int wfa_router_create(wfa_plan *p, int nsub, const int32_t *src_cell, const int32_t *poly_off,
                      const int32_t *poly_cells, int flags, wfa_router **out);
void wfa_router_destroy(wfa_router *r);
/* nnodes: sources + confluences; nlevels: evaluation levels (launches per apply); ncandidates: summed
 * length of the per-polygon node lists; nactive: grid cells at or below a source. */
int wfa_router_info(const wfa_router *r, int *nnodes, int *nlevels, int *ncandidates, long long *nactive);
int wfa_router_apply(wfa_router *r, int ncols, const double *QS, double *QR);           /* host buffers */
int wfa_router_apply_dev(wfa_router *r, int ncols, const double *QS_dev, double *QR_dev);
/* Which subbasins feed polygon `gauge` (0-based): the routed discharge of a polygon is the maximum over its candidate
 * nodes of the sum of QS over the node's upstream subbasins (Routing:119-123 on the operator's node forest).  With
 * QS >= 0 only the candidates that do not descend from another candidate matter; their upstream sets are disjoint.
 * *nmax = number of those; owner[s] (nsub entries) = index of the one whose set holds subbasin s, or -1 if s does
 * not reach the polygon.  gr2m_objective_routed_batch uses it to skip the routing altogether (see there). */
int wfa_router_gauge_upstream(wfa_router *r, int gauge, int *nmax, int32_t *owner);

/* Routed objective (north_star's simulate + route + objective loop): the reference's chain Run_GR2MSemiDistr ->
 * Routing_GR2MSemiDistr -> hydroGOF for nsets parameter sets at once.  Per set: QS of every subbasin rounded to 0.1
 * (Run:229-235), routed by the operator (Routing:100-123; months x sets are its columns), QR of polygon `gauge`
 * (0-based) rounded to 0.1 (Routing:131), scored against qobs as in Optim:199-212.  GR2M_SINK_UNROUNDED skips both
 * roundings.  Outputs as gr2m_objective_batch (out_qt: the gauge series, [nsets][ntime]).  The router must be FP64 and
 * built for the basin's subbasins in the same order.
 * With one region and at most four maximal candidates (wfa_router_gauge_upstream; a watershed polygon has one) nothing
 * is materialised: the recursion kernel sums the rounded QS of the upstream subbasins directly, one launch per
 * candidate, and subbasins that do not reach the gauge are not simulated.  Otherwise QS is written, the operator
 * applied and the gauge row read back; environment GR2M_ROUTED_MATERIALISE=1 forces that path (cross-check). */
int gr2m_objective_routed_batch(gr2m_basin *b, wfa_router *r, const double *params, int nsets, const double *qobs,
                                int gauge, int warmup, int which, int flags, double *out_of, double *out_crit,
                                double *out_qt); /* device, async */

/* Dense AreaD8 -wg for nmonth weight grids: W, A are [nmonth][ncell] (month planes, row-major
 * cells).  Nodata / contaminated cells come back as NaN (TauDEM -1 -> NA after raster()). */
int wfa_accumulate(wfa_plan *p, int nmonth, const double *W, int flags, double *A);

/* Device-resident batched sweep, in place: WA_dev is [ncell][ld] (cell-major, `ld` >= nmonth values
 * per cell, FP64 or FP32 per WFA_F32) holding weights on entry and accumulations on exit;
 * contaminated cells are NOT masked here (see wfa_mask_dev).  Asynchronous on the plan's stream. */
int wfa_accumulate_dev(wfa_plan *p, int nmonth, int ld, void *WA_dev, int flags);
int wfa_mask_dev(wfa_plan *p, int nmonth, int ld, void *WA_dev, int flags);
int wfa_plan_last_kernel_ms(wfa_plan *p, float *ms);
/* split of the last sweep: wide levels (one launch each) and the thin tail */
int wfa_plan_last_phase_ms(wfa_plan *p, float *wide_ms, float *tail_ms);

/* ---- routing preprocessing: depression fill + D8 flow directions on the device ----------------
 * Stands in for "pitremove" + "D8Flowdir" of Routing:88-93 (once per call).  dem, valid: [nrow x ncol]
 * row-major, valid != 0 marks cells with data; flowdir (out): TauDEM codes 1..8, -32768 = nodata;
 * filled (out, may be NULL): pit-filled elevations, NaN at invalid cells.  The flat-resolution rule is
 * this library's own (see csrc/d8prep.cu); any D8 grid, e.g. TauDEM's, can be given to wfa_plan_create. */
int wfa_flowdir_from_dem(const double *dem, const uint8_t *valid, int nrow, int ncol, int16_t *flowdir,
                         double *filled);
/* Same with every array already in device memory (blocking; the result can go straight to wfa_plan_create_dev). */
int wfa_flowdir_from_dem_dev(const double *dem_dev, const uint8_t *valid_dev, int nrow, int ncol, int16_t *flowdir_dev,
                             double *filled_dev);

#ifdef __cplusplus
}
#endif
#endif /* GR2M_B200_H */
