// router.cu -- month-invariant routing operator.
//
// Routing_GR2MSemiDistr puts one value per subbasin on the DEM (the cell of its point on surface,
// R/Routing_GR2MSemiDistr.R:77-81,109-113), every other weight is zero, and takes per polygon the largest
// accumulation over its fully covered cells (Routing:119-121).  With weights at nsub cells only, AreaD8's
// result changes along a flow path only at a source cell or where two paths that carry sources meet:
//
//   node      = a source cell, or a cell with >= 2 upstream neighbours that have a source above them
//   A[node]   = QS[source at node] (else 0) + sum over the node's child nodes, in AreaD8's k = 1..8 order
//               of the neighbour each child's path arrives through
//   A[cell]   = A[nearest node at or above it] for every other cell below a source, +0 elsewhere
//
// (adding the +0.0 of a source-free neighbour changes nothing, so the k-ordered FP sums are the ones the
// dense sweep makes, bit for bit -- only the sign of a zero can differ).  The node forest, its k-ordered
// child lists, the evaluation levels and, per polygon, the distinct nodes its interior cells see are
// functions of (D8 grid, source cells, polygons) only.  A router holds them; applying it to QS costs
// O((nodes + candidates) x columns) instead of O(grid cells x months), and the columns can be months of
// one run or months x parameter sets of a calibration alike.
#include "common.cuh"
#include "wfa_plan.cuh"

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>
#include <thrust/iterator/counting_iterator.h>
#include <math.h>
#include <stdlib.h>

#include <algorithm>
#include <atomic>
#include <new>
#include <vector>

namespace {

__constant__ int r_dx[9] = {0, 1, 1, 0, -1, -1, -1, 0, 1};
__constant__ int r_dy[9] = {0, 0, -1, -1, -1, 0, 1, 1, 1};

constexpr int32_t NODE_BIT = 0x40000000;

__device__ __forceinline__ bool act_get(const unsigned *act, long long c)
{
    return (act[c >> 2] >> ((c & 3) * 8)) & 1u;
}

// cell -> source index (-1 = none); sources on cells the plan never sweeps (nodata, cycles) carry nothing
__global__ void k_rt_srcmap(const int32_t *__restrict__ src, int nsub, const int32_t *__restrict__ level,
                            int32_t *__restrict__ srcmap)
{
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nsub) return;
    const int32_t c = src[s];
    if (c >= 0 && level[c] >= 0) srcmap[c] = s;
}

// mark every cell at or below a source; a walker stops at the first cell somebody else has marked
__global__ void k_rt_mark(const int32_t *__restrict__ src, int nsub, const int32_t *__restrict__ level,
                          const int32_t *__restrict__ receiver, unsigned *act)
{
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nsub) return;
    int32_t c = src[s];
    while (c >= 0 && level[c] >= 0) {
        const unsigned bit = 1u << ((c & 3) * 8);
        if (atomicOr(&act[c >> 2], bit) & bit) break;
        c = receiver[c];
    }
}

// per marked cell: node if it holds a source or >= 2 marked upstream neighbours
__global__ void k_rt_classify(const int32_t *__restrict__ acells, int nact, int ncol, const uint8_t *__restrict__ inmask,
                              const unsigned *__restrict__ act, const int32_t *__restrict__ srcmap,
                              int32_t *__restrict__ isnode)
{
    int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= nact) return;
    const int32_t c = acells[a];
    unsigned m = inmask[c];
    int nau = 0;
    while (m) {
        int k = __ffs(m);
        m &= m - 1;
        nau += act_get(act, (long long)c + r_dy[k] * ncol + r_dx[k]) ? 1 : 0;
    }
    isnode[a] = (srcmap[c] >= 0 || nau != 1) ? 1 : 0;
}

__global__ void k_rt_nodes(const int32_t *__restrict__ acells, int nact, const int32_t *__restrict__ isnode,
                           const int32_t *__restrict__ rank, const int32_t *__restrict__ srcmap,
                           int32_t *__restrict__ node_cell, int32_t *__restrict__ node_src)
{
    int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= nact || !isnode[a]) return;
    const int v = rank[a] - 1;                    // rank = inclusive scan of isnode
    node_cell[v] = acells[a];
    node_src[v] = srcmap[acells[a]];
}

__global__ void k_rt_seg_init(const int32_t *__restrict__ node_cell, int nn, int32_t *__restrict__ seg)
{
    int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v < nn) seg[node_cell[v]] = v | NODE_BIT;
}

// label the cells between a node and the next node downstream; record that node and the k slot
// (direction, seen from the parent, of the neighbour the path arrives through)
__global__ void k_rt_seg_walk(const int32_t *__restrict__ node_cell, int nn, int ncol,
                              const int32_t *__restrict__ level, const int32_t *__restrict__ receiver, int32_t *seg,
                              int32_t *__restrict__ parent, unsigned long long *__restrict__ ckey)
{
    int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nn) return;
    int32_t prev = node_cell[v], c = receiver[prev];
    int32_t par = -1;
    int kslot = 0;
    while (c >= 0 && level[c] >= 0) {
        const int32_t sv = seg[c];
        if (sv >= 0 && (sv & NODE_BIT)) {
            par = sv & ~NODE_BIT;
            const int dr = prev / ncol - c / ncol, dc = prev % ncol - c % ncol;
            for (int k = 1; k <= 8; ++k)
                if (r_dy[k] == dr && r_dx[k] == dc) kslot = k;
            break;
        }
        seg[c] = v;
        prev = c;
        c = receiver[c];
    }
    parent[v] = par;
    ckey[v] = par >= 0 ? ((unsigned long long)par << 4) | (unsigned long long)kslot : ~0ull;
}

__global__ void k_rt_child_count(const int32_t *__restrict__ parent, int nn, int32_t *__restrict__ cnt)
{
    int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v < nn && parent[v] >= 0) atomicAdd(&cnt[parent[v]], 1);
}

__global__ void k_rt_height_step(const int32_t *__restrict__ parent, int nn, int32_t *h, int *__restrict__ changed)
{
    int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nn || parent[v] < 0) return;
    const int want = ((volatile int32_t *)h)[v] + 1;
    if (atomicMax(&h[parent[v]], want) < want) *changed = 1;
}

__global__ void k_rt_iota(int32_t *__restrict__ a, int n)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = i;
}

__global__ void k_rt_height_hist(const int32_t *__restrict__ h, int nn, int32_t *__restrict__ hist)
{
    int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v < nn) atomicAdd(&hist[h[v]], 1);
}

// polygon interior cells -> (polygon, node) keys for cells below a source; a flag per polygon when it has a
// source-free (zero) interior cell.  Cells the plan never sweeps and contaminated cells are skipped, as in
// k_polygon_max.
__global__ void k_rt_cand(const int32_t *__restrict__ poff, int nsub, const int32_t *__restrict__ pcells, int npc,
                          const int32_t *__restrict__ level, const uint8_t *__restrict__ contam, int use_contam,
                          const unsigned *__restrict__ act, const int32_t *__restrict__ seg,
                          unsigned long long *__restrict__ key, uint8_t *__restrict__ flag, int32_t *__restrict__ haszero)
{
    int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= npc) return;
    int lo = 0, hi = nsub;                        // polygon s with poff[s] <= q < poff[s+1]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (poff[mid] <= q) lo = mid; else hi = mid;
    }
    const int32_t c = pcells[q];
    uint8_t f = 0;
    unsigned long long kk = 0;
    if (level[c] >= 0 && !(use_contam && contam[c])) {
        if (act_get(act, c)) {
            f = 1;
            kk = ((unsigned long long)lo << 32) | (unsigned long long)(seg[c] & ~NODE_BIT);
        } else {
            haszero[lo] = 1;
        }
    }
    key[q] = kk;
    flag[q] = f;
}

__global__ void k_rt_cand_off(const unsigned long long *__restrict__ ukey, int ncand, int nsub, int32_t *__restrict__ coff,
                              int32_t *__restrict__ cidx)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < ncand) cidx[i] = (int32_t)(ukey[i] & 0xffffffffull);
    if (i > nsub) return;
    const unsigned long long want = (unsigned long long)i << 32;   // first key of polygon >= i
    int lo = 0, hi = ncand;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (ukey[mid] < want) lo = mid + 1; else hi = mid;
    }
    coff[i] = lo;
}

// ---- apply ------------------------------------------------------------------------------------------
// nodes of one evaluation level; columns (months, or months x sets) on consecutive threads
template <typename T>
__global__ void k_rt_eval(const int32_t *__restrict__ horder, int v0, int v1, const int32_t *__restrict__ node_src,
                          const int32_t *__restrict__ child_off, const int32_t *__restrict__ child_idx,
                          const double *__restrict__ QS, int ldq, int c0, int ncols, T *val, int ldv)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int v = horder[v0 + blockIdx.y];
    if (t >= ncols) return;
    const int s = node_src[v];
    T acc = s >= 0 ? (T)QS[(size_t)s * ldq + c0 + t] : (T)0;
    for (int q = child_off[v]; q < child_off[v + 1]; ++q) acc += val[(size_t)child_idx[q] * ldv + t];
    val[(size_t)v * ldv + t] = acc;
}

template <typename T>
__global__ void k_rt_polymax(const int32_t *__restrict__ coff, const int32_t *__restrict__ cidx,
                             const int32_t *__restrict__ haszero, int nsub, const T *__restrict__ val, int ldv, int c0,
                             int ncols, double *__restrict__ QR, int ldq)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int s = blockIdx.y;
    if (t >= ncols) return;
    double m = haszero[s] ? 0.0 : -INFINITY;
    for (int q = coff[s]; q < coff[s + 1]; ++q) {
        const double v = (double)val[(size_t)cidx[q] * ldv + t];
        if (v > m) m = v;
    }
    QR[(size_t)s * ldq + c0 + t] = m;
}

}  // namespace

static std::atomic<long long> g_router_uid{0};

struct wfa_router {
    long long uid = ++g_router_uid;               // distinguishes routers that reuse an address (caches keyed on it)
    wfa_plan *plan = nullptr;
    int device = 0, nsub = 0, nn = 0, nh = 0, ncand = 0, flags = 0;
    long long nactive = 0;
    cudaStream_t stream = nullptr;
    DevBuf node_src, child_off, child_idx, horder, cand_off, cand_idx, haszero, qs, val, qr;
    std::vector<int32_t> h_hoff;                  // first node (in horder) of every evaluation level
};

static void router_free(wfa_router *r)
{
    if (!r) return;
    cudaSetDevice(r->device);
    DevBuf *bufs[] = {&r->node_src, &r->child_off, &r->child_idx, &r->horder, &r->cand_off, &r->cand_idx, &r->haszero,
                      &r->qs, &r->val, &r->qr};
    { FreeBatch fb; for (DevBuf *d : bufs) d->release(); }
    delete r;
}

static int router_build(wfa_router *r, const int32_t *src_host, const int32_t *poly_off, const int32_t *poly_cells)
{
    wfa_plan *p = r->plan;
    cudaStream_t st = r->stream;
    const long long n = p->ncell;
    const int nsub = r->nsub, npc = poly_off[nsub];
    const int use_contam = (r->flags & WFA_NO_CONTAMINATION) ? 0 : 1;
    DevBuf src, poff, pcells, srcmap, act, acells, nsel, temp, isnode, rank, node_cell, parent, ckey, ckey2, cval, cval2,
        ccnt, height, hkey2, hist, flag, key, ksel, ksort, kuniq;
    auto cleanup = [&]() {
        DevBuf *t[] = {&src, &poff, &pcells, &srcmap, &act, &acells, &nsel, &temp, &isnode, &rank, &node_cell, &parent,
                       &ckey, &ckey2, &cval, &cval2, &ccnt, &height, &hkey2, &hist, &flag, &key, &ksel, &ksort, &kuniq};
        FreeBatch fb;
        for (DevBuf *d : t) d->release();
    };
#define TCK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { gr2m_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); cleanup(); return e_ == cudaErrorMemoryAllocation ? GR2M_ERR_NOMEM : GR2M_ERR_CUDA; } } while (0)
#define TRC(call) do { int r_ = (call); if (r_) { cleanup(); return r_; } } while (0)
    const int32_t *level = p->level.as<int32_t>(), *receiver = p->receiver.as<int32_t>();
    const int nbs = ceil_div(nsub, 256);
    TRC(src.ensure((size_t)nsub * 4));
    TRC(poff.ensure(((size_t)nsub + 1) * 4));
    TRC(pcells.ensure(std::max<size_t>(4, (size_t)npc * 4)));
    TCK(cudaMemcpyAsync(src.p, src_host, (size_t)nsub * 4, cudaMemcpyHostToDevice, st));
    TCK(cudaMemcpyAsync(poff.p, poly_off, ((size_t)nsub + 1) * 4, cudaMemcpyHostToDevice, st));
    if (npc) TCK(cudaMemcpyAsync(pcells.p, poly_cells, (size_t)npc * 4, cudaMemcpyHostToDevice, st));
    // 1. cells at or below a source
    const size_t actbytes = ((size_t)n + 3) / 4 * 4;
    TRC(srcmap.ensure((size_t)n * 4));
    TRC(act.ensure(actbytes));
    TCK(cudaMemsetAsync(srcmap.p, 0xff, (size_t)n * 4, st));
    TCK(cudaMemsetAsync(act.p, 0, actbytes, st));
    gr2m_count_launch(), k_rt_srcmap<<<nbs, 256, 0, st>>>(src.as<int32_t>(), nsub, level, srcmap.as<int32_t>());
    gr2m_count_launch(), k_rt_mark<<<nbs, 256, 0, st>>>(src.as<int32_t>(), nsub, level, receiver, act.as<unsigned>());
    TCK(cudaGetLastError());
    TRC(acells.ensure(std::max<size_t>(4, (size_t)n * 4)));
    TRC(nsel.ensure(16));
    size_t tb = 0;
    thrust::counting_iterator<int32_t> iota(0);
    TCK(cub::DeviceSelect::Flagged(nullptr, tb, iota, act.as<uint8_t>(), acells.as<int32_t>(), nsel.as<int>(), (int)n, st));
    TRC(temp.ensure(std::max<size_t>(tb, 16)));
    TCK(cub::DeviceSelect::Flagged(temp.p, tb, iota, act.as<uint8_t>(), acells.as<int32_t>(), nsel.as<int>(), (int)n, st));
    int nact = 0;
    TCK(cudaMemcpyAsync(&nact, nsel.p, 4, cudaMemcpyDeviceToHost, st));
    TCK(cudaStreamSynchronize(st));
    r->nactive = nact;
    // 2. nodes
    int nn = 0;
    if (nact > 0) {
        const int nba = ceil_div(nact, 256);
        TRC(isnode.ensure((size_t)nact * 4));
        TRC(rank.ensure((size_t)nact * 4));
        gr2m_count_launch(), k_rt_classify<<<nba, 256, 0, st>>>(acells.as<int32_t>(), nact, p->ncol, p->inmask.as<uint8_t>(), act.as<unsigned>(),
                                           srcmap.as<int32_t>(), isnode.as<int32_t>());
        TCK(cudaGetLastError());
        TCK(cub::DeviceScan::InclusiveSum(nullptr, tb, isnode.as<int32_t>(), rank.as<int32_t>(), nact, st));
        if (tb > temp.cap) TRC(temp.ensure(tb));
        TCK(cub::DeviceScan::InclusiveSum(temp.p, tb, isnode.as<int32_t>(), rank.as<int32_t>(), nact, st));
        TCK(cudaMemcpyAsync(&nn, rank.as<int32_t>() + (nact - 1), 4, cudaMemcpyDeviceToHost, st));
        TCK(cudaStreamSynchronize(st));
    }
    r->nn = nn;
    TRC(r->node_src.ensure(std::max<size_t>(4, (size_t)nn * 4)));
    TRC(r->child_off.ensure(((size_t)nn + 1) * 4));
    TRC(r->child_idx.ensure(std::max<size_t>(4, (size_t)nn * 4)));
    TRC(r->horder.ensure(std::max<size_t>(4, (size_t)nn * 4)));
    r->h_hoff.assign(1, 0);
    r->nh = 0;
    if (nn > 0) {
        const int nba = ceil_div(nact, 256), nbn = ceil_div(nn, 256);
        TRC(node_cell.ensure((size_t)nn * 4));
        gr2m_count_launch(), k_rt_nodes<<<nba, 256, 0, st>>>(acells.as<int32_t>(), nact, isnode.as<int32_t>(), rank.as<int32_t>(),
                                        srcmap.as<int32_t>(), node_cell.as<int32_t>(), r->node_src.as<int32_t>());
        // 3. segments: srcmap becomes seg (cell -> nearest node at or above it)
        int32_t *seg = srcmap.as<int32_t>();
        TCK(cudaMemsetAsync(seg, 0xff, (size_t)n * 4, st));
        TRC(parent.ensure((size_t)nn * 4));
        TRC(ckey.ensure((size_t)nn * 8));
        TRC(ckey2.ensure((size_t)nn * 8));
        TRC(cval.ensure((size_t)nn * 4));
        TRC(ccnt.ensure(((size_t)nn + 1) * 4));
        gr2m_count_launch(), k_rt_seg_init<<<nbn, 256, 0, st>>>(node_cell.as<int32_t>(), nn, seg);
        gr2m_count_launch(), k_rt_seg_walk<<<nbn, 256, 0, st>>>(node_cell.as<int32_t>(), nn, p->ncol, level, receiver, seg, parent.as<int32_t>(),
                                           ckey.as<unsigned long long>());
        TCK(cudaGetLastError());
        // 4. children in (parent, k slot) order
        gr2m_count_launch(), k_rt_iota<<<nbn, 256, 0, st>>>(cval.as<int32_t>(), nn);
        TCK(cub::DeviceRadixSort::SortPairs(nullptr, tb, ckey.as<unsigned long long>(), ckey2.as<unsigned long long>(),
                                            cval.as<int32_t>(), r->child_idx.as<int32_t>(), nn, 0, 64, st));
        if (tb > temp.cap) TRC(temp.ensure(tb));
        TCK(cub::DeviceRadixSort::SortPairs(temp.p, tb, ckey.as<unsigned long long>(), ckey2.as<unsigned long long>(),
                                            cval.as<int32_t>(), r->child_idx.as<int32_t>(), nn, 0, 64, st));
        TCK(cudaMemsetAsync(ccnt.p, 0, ((size_t)nn + 1) * 4, st));
        gr2m_count_launch(), k_rt_child_count<<<nbn, 256, 0, st>>>(parent.as<int32_t>(), nn, ccnt.as<int32_t>());
        TCK(cudaGetLastError());
        TCK(cub::DeviceScan::ExclusiveSum(nullptr, tb, ccnt.as<int32_t>(), r->child_off.as<int32_t>(), nn + 1, st));
        if (tb > temp.cap) TRC(temp.ensure(tb));
        TCK(cub::DeviceScan::ExclusiveSum(temp.p, tb, ccnt.as<int32_t>(), r->child_off.as<int32_t>(), nn + 1, st));
        // 5. evaluation levels: height of every node above its deepest leaf
        TRC(height.ensure((size_t)nn * 4));
        TRC(nsel.ensure(16));
        TCK(cudaMemsetAsync(height.p, 0, (size_t)nn * 4, st));
        int changed = 1;
        while (changed) {
            TCK(cudaMemsetAsync(nsel.p, 0, 4, st));
            for (int it = 0; it < 16; ++it)
                gr2m_count_launch(), k_rt_height_step<<<nbn, 256, 0, st>>>(parent.as<int32_t>(), nn, height.as<int32_t>(), nsel.as<int>());
            TCK(cudaGetLastError());
            TCK(cudaMemcpyAsync(&changed, nsel.p, 4, cudaMemcpyDeviceToHost, st));
            TCK(cudaStreamSynchronize(st));
        }
        TRC(hkey2.ensure((size_t)nn * 4));
        gr2m_count_launch(), k_rt_iota<<<nbn, 256, 0, st>>>(cval.as<int32_t>(), nn);
        TCK(cub::DeviceRadixSort::SortPairs(nullptr, tb, height.as<int32_t>(), hkey2.as<int32_t>(), cval.as<int32_t>(),
                                            r->horder.as<int32_t>(), nn, 0, 32, st));
        if (tb > temp.cap) TRC(temp.ensure(tb));
        TCK(cub::DeviceRadixSort::SortPairs(temp.p, tb, height.as<int32_t>(), hkey2.as<int32_t>(), cval.as<int32_t>(),
                                            r->horder.as<int32_t>(), nn, 0, 32, st));
        int hmax = 0;
        TCK(cudaMemcpyAsync(&hmax, hkey2.as<int32_t>() + (nn - 1), 4, cudaMemcpyDeviceToHost, st));
        TCK(cudaStreamSynchronize(st));
        r->nh = hmax + 1;
        TRC(hist.ensure(((size_t)r->nh) * 4));
        TCK(cudaMemsetAsync(hist.p, 0, (size_t)r->nh * 4, st));
        gr2m_count_launch(), k_rt_height_hist<<<nbn, 256, 0, st>>>(height.as<int32_t>(), nn, hist.as<int32_t>());
        TCK(cudaGetLastError());
        std::vector<int32_t> hh(r->nh);
        TCK(cudaMemcpyAsync(hh.data(), hist.p, (size_t)r->nh * 4, cudaMemcpyDeviceToHost, st));
        TCK(cudaStreamSynchronize(st));
        r->h_hoff.assign(r->nh + 1, 0);
        for (int i = 0; i < r->nh; ++i) r->h_hoff[i + 1] = r->h_hoff[i] + hh[i];
    } else {
        TCK(cudaMemsetAsync(r->child_off.p, 0, 4, st));
    }
    // 6. per polygon: the distinct nodes its interior cells see (+ whether it has a zero cell)
    TRC(r->haszero.ensure((size_t)nsub * 4));
    TRC(r->cand_off.ensure(((size_t)nsub + 1) * 4));
    TCK(cudaMemsetAsync(r->haszero.p, 0, (size_t)nsub * 4, st));
    int ncand = 0;
    if (npc > 0) {
        TRC(key.ensure((size_t)npc * 8));
        TRC(flag.ensure((size_t)npc));
        TRC(ksel.ensure((size_t)npc * 8));
        gr2m_count_launch(), k_rt_cand<<<ceil_div(npc, 256), 256, 0, st>>>(poff.as<int32_t>(), nsub, pcells.as<int32_t>(), npc, level,
                                                      p->contam.as<uint8_t>(), use_contam, act.as<unsigned>(),
                                                      srcmap.as<int32_t>(), key.as<unsigned long long>(),
                                                      flag.as<uint8_t>(), r->haszero.as<int32_t>());
        TCK(cudaGetLastError());
        TCK(cub::DeviceSelect::Flagged(nullptr, tb, key.as<unsigned long long>(), flag.as<uint8_t>(),
                                       ksel.as<unsigned long long>(), nsel.as<int>(), npc, st));
        if (tb > temp.cap) TRC(temp.ensure(tb));
        TCK(cub::DeviceSelect::Flagged(temp.p, tb, key.as<unsigned long long>(), flag.as<uint8_t>(),
                                       ksel.as<unsigned long long>(), nsel.as<int>(), npc, st));
        int nk = 0;
        TCK(cudaMemcpyAsync(&nk, nsel.p, 4, cudaMemcpyDeviceToHost, st));
        TCK(cudaStreamSynchronize(st));
        if (nk > 0) {
            TRC(ksort.ensure((size_t)nk * 8));
            TRC(kuniq.ensure((size_t)nk * 8));
            TCK(cub::DeviceRadixSort::SortKeys(nullptr, tb, ksel.as<unsigned long long>(), ksort.as<unsigned long long>(),
                                               nk, 0, 64, st));
            if (tb > temp.cap) TRC(temp.ensure(tb));
            TCK(cub::DeviceRadixSort::SortKeys(temp.p, tb, ksel.as<unsigned long long>(), ksort.as<unsigned long long>(),
                                               nk, 0, 64, st));
            TCK(cub::DeviceSelect::Unique(nullptr, tb, ksort.as<unsigned long long>(), kuniq.as<unsigned long long>(),
                                          nsel.as<int>(), nk, st));
            if (tb > temp.cap) TRC(temp.ensure(tb));
            TCK(cub::DeviceSelect::Unique(temp.p, tb, ksort.as<unsigned long long>(), kuniq.as<unsigned long long>(),
                                          nsel.as<int>(), nk, st));
            TCK(cudaMemcpyAsync(&ncand, nsel.p, 4, cudaMemcpyDeviceToHost, st));
            TCK(cudaStreamSynchronize(st));
        }
    }
    r->ncand = ncand;
    TRC(r->cand_idx.ensure(std::max<size_t>(4, (size_t)ncand * 4)));
    gr2m_count_launch(), k_rt_cand_off<<<ceil_div(std::max(ncand, nsub + 1), 256), 256, 0, st>>>(kuniq.as<unsigned long long>(), ncand, nsub,
                                                                          r->cand_off.as<int32_t>(),
                                                                          r->cand_idx.as<int32_t>());
    TCK(cudaGetLastError());
    TCK(cudaStreamSynchronize(st));
    cleanup();
#undef TCK
#undef TRC
    return GR2M_OK;
}

extern "C" int wfa_router_create(wfa_plan *p, int nsub, const int32_t *src_cell, const int32_t *poly_off,
                                 const int32_t *poly_cells, int flags, wfa_router **out)
{
    REQUIRE(p && src_cell && poly_off && poly_cells && out, "NULL argument");
    REQUIRE(nsub > 0, "nsub must be positive");
    const long long n = p->ncell;
    REQUIRE(n < (1ll << 30), "grid too large for the routing operator");
    for (int s = 0; s < nsub; ++s) {
        REQUIRE(src_cell[s] >= 0 && src_cell[s] < n, "src_cell out of range (point on surface outside the DEM)");
        REQUIRE(poly_off[s + 1] >= poly_off[s], "poly_off must be non-decreasing");
    }
    REQUIRE(poly_off[0] == 0, "poly_off[0] must be 0");
    const int npc = poly_off[nsub];
    for (int q = 0; q < npc; ++q) REQUIRE(poly_cells[q] >= 0 && poly_cells[q] < n, "poly_cells out of range");
    CK(cudaSetDevice(p->device));
    // duplicates: the last assignment wins, as R's `[<-` (Routing:109-113)
    std::vector<int32_t> src(src_cell, src_cell + nsub);
    {
        std::vector<int> idx(nsub);
        for (int s = 0; s < nsub; ++s) idx[s] = s;
        std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return src_cell[a] < src_cell[b]; });
        for (int i = 0; i + 1 < nsub; ++i)
            if (src_cell[idx[i]] == src_cell[idx[i + 1]]) src[idx[i]] = -1;
    }
    wfa_router *r = new (std::nothrow) wfa_router;
    if (!r) { gr2m_set_error("out of host memory"); return GR2M_ERR_NOMEM; }
    r->plan = p; r->device = p->device; r->nsub = nsub; r->flags = flags; r->stream = p->stream;
    int rc = router_build(r, src.data(), poly_off, poly_cells);
    if (rc) { router_free(r); return rc; }
    *out = r;
    return GR2M_OK;
}

extern "C" void wfa_router_destroy(wfa_router *r) { router_free(r); }

extern "C" int wfa_router_info(const wfa_router *r, int *nnodes, int *nlevels, int *ncandidates, long long *nactive)
{
    REQUIRE(r != nullptr, "router is NULL");
    if (nnodes) *nnodes = r->nn;
    if (nlevels) *nlevels = r->nh;
    if (ncandidates) *ncandidates = r->ncand;
    if (nactive) *nactive = r->nactive;
    return GR2M_OK;
}

template <typename T>
static int router_apply_t(wfa_router *r, int ncols, const double *QS_dev, double *QR_dev)
{
    cudaStream_t st = r->stream;
    // columns in chunks so the node values stay within ~8 GB
    long long fit = (8ll << 30) / ((long long)std::max(r->nn, 1) * (long long)sizeof(T));
    int chunk = (int)std::min<long long>(ncols, std::max<long long>(32, fit / 32 * 32));
    int rc;
    if ((rc = r->val.ensure(std::max<size_t>(16, (size_t)r->nn * chunk * sizeof(T))))) return rc;
    T *val = r->val.as<T>();
    for (int c0 = 0; c0 < ncols; c0 += chunk) {
        const int nc = std::min(chunk, ncols - c0);
        const int bx = ceil_div(nc, 128);
        for (int h = 0; h < r->nh; ++h) {
            const int v0 = r->h_hoff[h], v1 = r->h_hoff[h + 1];
            for (int b = v0; b < v1; b += 65535) {
                dim3 grid(bx, std::min(65535, v1 - b));
                gr2m_count_launch(), k_rt_eval<T><<<grid, 128, 0, st>>>(r->horder.as<int32_t>(), b, v1, r->node_src.as<int32_t>(),
                                                   r->child_off.as<int32_t>(), r->child_idx.as<int32_t>(), QS_dev, ncols,
                                                   c0, nc, val, chunk);
            }
        }
        for (int b = 0; b < r->nsub; b += 65535) {
            dim3 grid(bx, std::min(65535, r->nsub - b));
            gr2m_count_launch(), k_rt_polymax<T><<<grid, 128, 0, st>>>(r->cand_off.as<int32_t>() + b, r->cand_idx.as<int32_t>(),
                                                  r->haszero.as<int32_t>() + b, r->nsub - b, val, chunk, c0, nc,
                                                  QR_dev + (size_t)b * ncols, ncols);
        }
        CK(cudaGetLastError());
    }
    return GR2M_OK;
}

extern "C" int wfa_router_apply_dev(wfa_router *r, int ncols, const double *QS_dev, double *QR_dev)
{
    REQUIRE(r && QS_dev && QR_dev, "NULL argument");
    REQUIRE(ncols > 0, "ncols must be positive");
    CK(cudaSetDevice(r->device));
    r->stream = r->plan->stream;
    return (r->flags & WFA_F32) ? router_apply_t<float>(r, ncols, QS_dev, QR_dev)
                                : router_apply_t<double>(r, ncols, QS_dev, QR_dev);
}

// for gr2m_objective_routed_batch (gr2m.cu): number of subbasins, and "wait until the last apply is done"
int wfa_router_nsub(const wfa_router *r) { return r ? r->nsub : -1; }
int wfa_router_is_f32(const wfa_router *r) { return r && (r->flags & WFA_F32) ? 1 : 0; }
// For gr2m_objective_routed_batch's fused path.  The routed discharge of one polygon is max over its candidate nodes of
// the node's accumulated value = the sum of QS over the subbasins upstream of the node, so for a FIXED gauge the whole
// routing collapses to weighted outlet sums.  In the node forest two nodes' upstream sets are nested or disjoint, and QS
// is never negative, so only the candidates that do not descend from another candidate matter ("maximal" ones), and
// their upstream sets are disjoint: owner[s] = index of the maximal candidate whose upstream set holds subbasin s, or -1.
// Host-side walk over a copy of the forest (a few integers per node; done once per router and gauge by the caller).
// *nmax = number of maximal candidates (0: the polygon has none -- the caller keeps the materialised path).
int wfa_router_gauge_owner(wfa_router *r, int gauge, int *nmax, int32_t *owner_host)
{
    REQUIRE(r && nmax && owner_host, "NULL argument");
    REQUIRE(gauge >= 0 && gauge < r->nsub, "gauge polygon index out of range");
    CK(cudaSetDevice(r->device));
    const int nn = r->nn;
    for (int s = 0; s < r->nsub; ++s) owner_host[s] = -1;
    *nmax = 0;
    if (nn <= 0) return GR2M_OK;
    std::vector<int32_t> nsrc(nn), coff(nn + 1), cidx(nn), goff(2);
    CK(cudaStreamSynchronize(r->plan->stream));
    CK(cudaMemcpy(nsrc.data(), r->node_src.p, (size_t)nn * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(coff.data(), r->child_off.p, ((size_t)nn + 1) * 4, cudaMemcpyDeviceToHost));
    if (coff[nn] > 0) CK(cudaMemcpy(cidx.data(), r->child_idx.p, (size_t)coff[nn] * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(goff.data(), r->cand_off.as<int32_t>() + gauge, 8, cudaMemcpyDeviceToHost));
    const int nc = goff[1] - goff[0];
    if (nc <= 0) return GR2M_OK;
    std::vector<int32_t> cand(nc);
    CK(cudaMemcpy(cand.data(), r->cand_idx.as<int32_t>() + goff[0], (size_t)nc * 4, cudaMemcpyDeviceToHost));
    std::sort(cand.begin(), cand.end());
    cand.erase(std::unique(cand.begin(), cand.end()), cand.end());
    std::vector<int32_t> parent(nn, -1);
    for (int v = 0; v < nn; ++v)
        for (int q = coff[v]; q < coff[v + 1]; ++q) parent[cidx[q]] = v;
    std::vector<char> iscand(nn, 0);
    for (int c : cand) iscand[c] = 1;
    std::vector<int32_t> stack;
    int k = 0;
    for (int c : cand) {
        bool maximal = true;
        for (int v = parent[c]; v >= 0; v = parent[v])
            if (iscand[v]) { maximal = false; break; }
        if (!maximal) continue;
        stack.assign(1, c);
        while (!stack.empty()) {
            const int v = stack.back();
            stack.pop_back();
            if (nsrc[v] >= 0) owner_host[nsrc[v]] = k;
            for (int q = coff[v]; q < coff[v + 1]; ++q) stack.push_back(cidx[q]);
        }
        ++k;
    }
    *nmax = k;
    return GR2M_OK;
}

long long wfa_router_uid(const wfa_router *r) { return r ? r->uid : 0; }

extern "C" int wfa_router_gauge_upstream(wfa_router *r, int gauge, int *nmax, int32_t *owner)
{
    return wfa_router_gauge_owner(r, gauge, nmax, owner);
}

int wfa_router_wait(wfa_router *r)
{
    REQUIRE(r != nullptr, "router is NULL");
    CK(cudaSetDevice(r->device));
    CK(cudaStreamSynchronize(r->plan->stream));
    return GR2M_OK;
}

extern "C" int wfa_router_apply(wfa_router *r, int ncols, const double *QS, double *QR)
{
    REQUIRE(r && QS && QR, "NULL argument");
    REQUIRE(ncols > 0, "ncols must be positive");
    CK(cudaSetDevice(r->device));
    r->stream = r->plan->stream;
    const size_t bytes = (size_t)r->nsub * ncols * 8;
    int rc;
    if ((rc = r->qs.ensure(bytes))) return rc;
    if ((rc = r->qr.ensure(bytes))) return rc;
    CK(cudaMemcpyAsync(r->qs.p, QS, bytes, cudaMemcpyHostToDevice, r->stream));
    if ((rc = wfa_router_apply_dev(r, ncols, r->qs.as<double>(), r->qr.as<double>()))) return rc;
    CK(cudaMemcpyAsync(QR, r->qr.p, bytes, cudaMemcpyDeviceToHost, r->stream));
    CK(cudaStreamSynchronize(r->stream));
    return GR2M_OK;
}
