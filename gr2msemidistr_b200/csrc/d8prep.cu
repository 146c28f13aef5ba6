// d8prep.cu -- depression fill + D8 flow directions on the device (sm_100a).
//
// Replaces, for Routing_GR2MSemiDistr, the two TauDEM preprocessing processes
// "pitremove -z dem.tif -fel CONDEM.tif" and "D8Flowdir -p FlowDir.tif -fel CONDEM.tif"
// (R/Routing_GR2MSemiDistr.R:88-93).  Runs once per call; not on the per-month path.  The rule is
// this package's own (TauDEM's flat resolution is not reproduced) and is order-independent, so this
// parallel version and the CPU reference (oracle/d8prep.py) agree bit for bit:
//   fill      filled[c] = max(z[c], min over neighbours filled[n]) relaxed from +inf to its greatest
//             fixed point, outlets (valid cells next to invalid / off-grid ones) pinned at z;
//   descent   steepest strictly lower neighbour on the filled surface (drop / distance, first k on ties);
//   outlets   without a lower neighbour point at their first invalid / off-grid neighbour;
//   flats     level-synchronous BFS through equal filled elevation towards the nearest resolved cell.
#include "common.cuh"

#include <math.h>

namespace {

__constant__ int p_dx[9] = {0, 1, 1, 0, -1, -1, -1, 0, 1};
__constant__ int p_dy[9] = {0, 0, -1, -1, -1, 0, 1, 1, 1};
constexpr int TILE = 32, INNER = 48;
constexpr double SQRT2 = 1.4142135623730951;

__global__ void k_d8_init(const double *__restrict__ z, const uint8_t *__restrict__ valid, int nrow, int ncol,
                          double *__restrict__ F, uint8_t *__restrict__ edge)
{
    int c = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
    if (c >= ncol) return;
    size_t i = (size_t)r * ncol + c;
    if (!valid[i]) { F[i] = INFINITY; edge[i] = 0; return; }
    bool e = false;
    for (int k = 1; k <= 8; ++k) {
        int r2 = r + p_dy[k], c2 = c + p_dx[k];
        if (r2 < 0 || r2 >= nrow || c2 < 0 || c2 >= ncol || !valid[(size_t)r2 * ncol + c2]) e = true;
    }
    edge[i] = e;
    F[i] = e ? z[i] : INFINITY;
}

// One 32 x 32 tile per CTA, relaxed INNER times in shared memory with its halo frozen.
__global__ void __launch_bounds__(TILE * TILE) k_d8_fill(const double *__restrict__ z, const uint8_t *__restrict__ valid,
                                                       const uint8_t *__restrict__ edge, int nrow, int ncol, double *F,
                                                       int *changed)
{
    __shared__ double s[TILE + 2][TILE + 2];
    __shared__ int s_changed;
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int r0 = blockIdx.y * TILE, c0 = blockIdx.x * TILE;
    if (tx == 0 && ty == 0) s_changed = 0;
    for (int idx = ty * TILE + tx; idx < (TILE + 2) * (TILE + 2); idx += TILE * TILE) {
        int rr = idx / (TILE + 2), cc = idx % (TILE + 2);
        int r = r0 + rr - 1, c = c0 + cc - 1;
        bool in = r >= 0 && r < nrow && c >= 0 && c < ncol;
        s[rr][cc] = (in && valid[(size_t)r * ncol + c]) ? F[(size_t)r * ncol + c] : INFINITY;
    }
    __syncthreads();
    const int r = r0 + ty, c = c0 + tx;
    const bool in = r < nrow && c < ncol;
    const size_t i = in ? (size_t)r * ncol + c : 0;
    const bool active = in && valid[i] && !edge[i];
    const double zc = active ? z[i] : 0.0;
    bool mine = false;
    for (int it = 0; it < INNER; ++it) {
        if (active) {
            double m = INFINITY;
#pragma unroll
            for (int k = 1; k <= 8; ++k) m = fmin(m, s[ty + 1 + p_dy[k]][tx + 1 + p_dx[k]]);
            double nf = fmax(zc, m);
            if (nf < s[ty + 1][tx + 1]) { s[ty + 1][tx + 1] = nf; mine = true; }
        }
        __syncthreads();
    }
    if (mine) { F[i] = s[ty + 1][tx + 1]; s_changed = 1; }
    __syncthreads();
    if (tx == 0 && ty == 0 && s_changed) *changed = 1;
}

__global__ void k_d8_descent(const double *__restrict__ F, const uint8_t *__restrict__ valid,
                             const uint8_t *__restrict__ edge, int nrow, int ncol, int16_t *__restrict__ code,
                             int32_t *__restrict__ dist)
{
    int c = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
    if (c >= ncol) return;
    size_t i = (size_t)r * ncol + c;
    if (!valid[i]) { code[i] = 0; dist[i] = -1; return; }
    double best = 0.0, f = F[i];
    int bk = 0, first_out = 0;
    for (int k = 1; k <= 8; ++k) {
        int r2 = r + p_dy[k], c2 = c + p_dx[k];
        bool ok = r2 >= 0 && r2 < nrow && c2 >= 0 && c2 < ncol && valid[(size_t)r2 * ncol + c2];
        if (!ok) { if (!first_out) first_out = k; continue; }
        double s = __ddiv_rn(__dsub_rn(f, F[(size_t)r2 * ncol + c2]), (k & 1) ? 1.0 : SQRT2);
        if (s > best) { best = s; bk = k; }
    }
    if (bk == 0 && edge[i]) bk = first_out;
    code[i] = (int16_t)bk;
    dist[i] = bk > 0 ? 0 : -1;
}

__global__ void k_d8_flat_step(const double *__restrict__ F, const uint8_t *__restrict__ valid, int nrow, int ncol, int t,
                               int16_t *code, int32_t *dist, int *changed)
{
    int c = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
    if (c >= ncol) return;
    size_t i = (size_t)r * ncol + c;
    if (!valid[i] || dist[i] >= 0) return;
    double f = F[i];
    for (int k = 1; k <= 8; ++k) {
        int r2 = r + p_dy[k], c2 = c + p_dx[k];
        if (r2 < 0 || r2 >= nrow || c2 < 0 || c2 >= ncol) continue;
        size_t j = (size_t)r2 * ncol + c2;
        if (valid[j] && F[j] == f && ((volatile int32_t *)dist)[j] == t - 1) {
            code[i] = (int16_t)k;
            dist[i] = t;      // t != t-1: neighbours assigned in this same step are not mistaken for step t-1
            *changed = 1;
            return;
        }
    }
}

__global__ void k_d8_finish(const uint8_t *__restrict__ valid, size_t n, int16_t *__restrict__ code, double *F)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (!valid[i]) { code[i] = -32768; if (F) F[i] = NAN; }
    else if (code[i] <= 0) code[i] = -32768;
}

}  // namespace

// dem / valid / flowdir / filled are host pointers, or device pointers when on_device
static int flowdir_common(const double *dem, const uint8_t *valid, int nrow, int ncol, int16_t *flowdir, double *filled,
                          bool on_device)
{
    REQUIRE(dem && valid && flowdir, "NULL argument");
    REQUIRE(nrow > 0 && ncol > 0, "nrow and ncol must be positive");
    int rc = gr2m_check_device();
    if (rc) return rc;
    CK(cudaSetDevice(gr2m_current_device()));
    const size_t n = (size_t)nrow * ncol;
    DevBuf bz, bF, bv, be, bc, bd, bflag;       // through the handle allocator: block cache, guard-band debug mode
    auto release = [&]() { FreeBatch fb; for (DevBuf *b : {&bz, &bF, &bv, &be, &bc, &bd, &bflag}) b->release(); };
#define DCK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { gr2m_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); release(); return e_ == cudaErrorMemoryAllocation ? GR2M_ERR_NOMEM : GR2M_ERR_CUDA; } } while (0)
#define DRC(call) do { int r_ = (call); if (r_) { release(); return r_; } } while (0)
    if (!on_device) { DRC(bz.ensure(n * 8)); DRC(bv.ensure(n)); DRC(bc.ensure(n * 2)); }
    if (!on_device || !filled) DRC(bF.ensure(n * 8));
    DRC(be.ensure(n)); DRC(bd.ensure(n * 4)); DRC(bflag.ensure(sizeof(int)));
    const double *dz = on_device ? dem : bz.as<double>();
    const uint8_t *dv = on_device ? valid : bv.as<uint8_t>();
    double *dF = (on_device && filled) ? filled : bF.as<double>();
    uint8_t *de = be.as<uint8_t>();
    int16_t *dc = on_device ? flowdir : bc.as<int16_t>();
    int32_t *dd = bd.as<int32_t>();
    int *dflag = bflag.as<int>();
    if (!on_device) {
        DCK(cudaMemcpy(bz.p, dem, n * 8, cudaMemcpyHostToDevice));
        DCK(cudaMemcpy(bv.p, valid, n, cudaMemcpyHostToDevice));
    }
    dim3 rowgrid(ceil_div(ncol, 256), nrow);
    gr2m_count_launch(), k_d8_init<<<rowgrid, 256>>>(dz, dv, nrow, ncol, dF, de);
    DCK(cudaGetLastError());
    dim3 tgrid(ceil_div(ncol, TILE), ceil_div(nrow, TILE)), tblk(TILE, TILE);
    int flag = 1;
    for (long long it = 0; flag && it < (long long)nrow + ncol + 16; ++it) {   // each pass moves >= INNER cells
        DCK(cudaMemset(dflag, 0, sizeof(int)));
        gr2m_count_launch(), k_d8_fill<<<tgrid, tblk>>>(dz, dv, de, nrow, ncol, dF, dflag);
        DCK(cudaGetLastError());
        DCK(cudaMemcpy(&flag, dflag, sizeof(int), cudaMemcpyDeviceToHost));
    }
    gr2m_count_launch(), k_d8_descent<<<rowgrid, 256>>>(dF, dv, de, nrow, ncol, dc, dd);
    DCK(cudaGetLastError());
    flag = 1;
    for (int t = 1; flag && t < nrow * ncol + 2; ++t) {
        DCK(cudaMemset(dflag, 0, sizeof(int)));
        gr2m_count_launch(), k_d8_flat_step<<<rowgrid, 256>>>(dF, dv, nrow, ncol, t, dc, dd, dflag);
        DCK(cudaGetLastError());
        DCK(cudaMemcpy(&flag, dflag, sizeof(int), cudaMemcpyDeviceToHost));
    }
    gr2m_count_launch(), k_d8_finish<<<ceil_div((long long)n, 256), 256>>>(dv, n, dc, filled ? dF : nullptr);
    DCK(cudaGetLastError());
    if (!on_device) {
        DCK(cudaMemcpy(flowdir, dc, n * 2, cudaMemcpyDeviceToHost));
        if (filled) DCK(cudaMemcpy(filled, dF, n * 8, cudaMemcpyDeviceToHost));
    } else {
        DCK(cudaDeviceSynchronize());
    }
#undef DCK
#undef DRC
    release();
    return GR2M_OK;
}

extern "C" int wfa_flowdir_from_dem(const double *dem, const uint8_t *valid, int nrow, int ncol, int16_t *flowdir,
                                    double *filled)
{
    return flowdir_common(dem, valid, nrow, ncol, flowdir, filled, false);
}

extern "C" int wfa_flowdir_from_dem_dev(const double *dem_dev, const uint8_t *valid_dev, int nrow, int ncol,
                                        int16_t *flowdir_dev, double *filled_dev)
{
    return flowdir_common(dem_dev, valid_dev, nrow, ncol, flowdir_dev, filled_dev, true);
}
