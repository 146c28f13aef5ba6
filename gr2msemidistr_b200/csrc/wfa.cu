// wfa.cu -- Weighted Flow Accumulation (TauDEM "AreaD8 -wg" semantics) on a D8 grid, all months
// batched, and its C ABI (sm_100a).  Replaces the month loop of R/Routing_GR2MSemiDistr.R:100-123.
//
// Plan (month-invariant, built once on the device, integer, deterministic):
//   receiver[c], inmask[c] (bit k-1: neighbour k drains into c), indeg[c], level[c] (Kahn frontier
//   index), order[] = cells sorted by (level, cell id), level_off[], contam[c] (TauDEM edge
//   contamination, which depends on topology only).
// Sweep: accumulation buffer A[cell][ld] is cell-major with the month batch contiguous, so one
// sub-warp group handles one cell with the months on its lanes: every upstream row is one coalesced
// read, and a cell's sum is taken in TauDEM's fixed k = 1..8 neighbour order -- deterministic with
// no atomics and no cross-lane reduction.  Cells of one level are independent; levels are swept in
// order: the wide levels below a cut as one launch each (bandwidth-bound), the levels above it per
// outlet basin -- the upper network is a forest of independent trees, one CTA walks one tree with
// block barriers only (k_accum_components*).  When the forest does not split into enough trees the
// upper levels are swept grid-wide instead: one launch per level, then a single resident CTA for the
// long thin tail (WFA_LEVEL_SYNC), or a ticketed flag-driven dataflow kernel (WFA_DATAFLOW).
#include "common.cuh"
#include "wfa_plan.cuh"

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <stdlib.h>

#include <algorithm>
#include <new>
#include <vector>

namespace {

__constant__ int c_dx[9] = {0, 1, 1, 0, -1, -1, -1, 0, 1};
__constant__ int c_dy[9] = {0, 0, -1, -1, -1, 0, 1, 1, 1};

__device__ __forceinline__ bool d8_valid(int d) { return d >= 1 && d <= 8; }

// ---- plan construction --------------------------------------------------------------------------
__global__ void k_topology(const int16_t *__restrict__ dir, int nrow, int ncol, int32_t *__restrict__ receiver,
                           uint8_t *__restrict__ inmask, int32_t *__restrict__ indeg, int32_t *__restrict__ cnt,
                           uint8_t *__restrict__ contam, int32_t *__restrict__ level)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)nrow * ncol) return;
    int r = (int)(i / ncol), c = (int)(i % ncol);
    int d = dir[i];
    level[i] = -1;
    if (!d8_valid(d)) {
        receiver[i] = -1; inmask[i] = 0; indeg[i] = 0; cnt[i] = -1; contam[i] = 0;
        return;
    }
    int rr = r + c_dy[d], cc = c + c_dx[d];
    int32_t rec = -1;
    if (rr >= 0 && rr < nrow && cc >= 0 && cc < ncol && d8_valid(dir[(long long)rr * ncol + cc]))
        rec = (int32_t)((long long)rr * ncol + cc);
    unsigned m = 0, edge = 0;
#pragma unroll
    for (int k = 1; k <= 8; ++k) {
        int r2 = r + c_dy[k], c2 = c + c_dx[k];
        if (r2 < 0 || r2 >= nrow || c2 < 0 || c2 >= ncol) { edge = 1; continue; }
        int dn = dir[(long long)r2 * ncol + c2];
        if (!d8_valid(dn)) { edge = 1; continue; }
        if (dn - k == 4 || dn - k == -4) m |= 1u << (k - 1);
    }
    receiver[i] = rec; inmask[i] = (uint8_t)m; indeg[i] = __popc(m); cnt[i] = __popc(m);
    contam[i] = (uint8_t)edge;   // seeded with the cell's own edge flag; upstream OR-ed in while peeling
}

__global__ void k_frontier_init(const int32_t *__restrict__ cnt, long long n, int32_t *__restrict__ frontier,
                                int *__restrict__ count)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && cnt[i] == 0) frontier[atomicAdd(count, 1)] = (int32_t)i;
}

__device__ __forceinline__ void peel_cell(int32_t c, int lev, int ncol, const int32_t *__restrict__ receiver,
                                          const uint8_t *__restrict__ inmask, int32_t *cnt, uint8_t *contam,
                                          int32_t *level, int32_t *next, int *next_count)
{
    level[c] = lev;
    unsigned m = inmask[c], con = contam[c];
    while (m) {
        int k = __ffs(m);   // 1-based bit index == neighbour k
        m &= m - 1;
        con |= ((volatile uint8_t *)contam)[c + c_dy[k] * ncol + c_dx[k]];
    }
    contam[c] = (uint8_t)con;
    int32_t d = receiver[c];
    if (d >= 0 && atomicSub(&cnt[d], 1) == 1) next[atomicAdd(next_count, 1)] = d;
}

// one Kahn frontier per launch (wide frontiers)
__global__ void k_peel(const int32_t *__restrict__ cur, int ncur, int lev, int ncol,
                       const int32_t *__restrict__ receiver, const uint8_t *__restrict__ inmask, int32_t *cnt,
                       uint8_t *contam, int32_t *level, int32_t *next, int *next_count)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < ncur) peel_cell(cur[i], lev, ncol, receiver, inmask, cnt, contam, level, next, next_count);
}

// many thin frontiers inside one resident CTA; stops when the frontier is empty or wider than `limit`
// state[0] = level, state[1] = frontier size, state[2] = which buffer holds the frontier (0/1)
__global__ void __launch_bounds__(1024) k_peel_thin(int32_t *buf0, int32_t *buf1, int *state, int limit, int ncol,
                                                    const int32_t *__restrict__ receiver,
                                                    const uint8_t *__restrict__ inmask, int32_t *cnt, uint8_t *contam,
                                                    int32_t *level)
{
    __shared__ int s_next;
    int lev = state[0], n = state[1], which = state[2];
    while (n > 0 && n <= limit) {
        int32_t *cur = which ? buf1 : buf0, *nxt = which ? buf0 : buf1;
        if (threadIdx.x == 0) s_next = 0;
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += blockDim.x)
            peel_cell(cur[i], lev, ncol, receiver, inmask, cnt, contam, level, nxt, &s_next);
        __threadfence();
        __syncthreads();
        n = s_next;
        which ^= 1;
        ++lev;
        __syncthreads();
    }
    if (threadIdx.x == 0) { state[0] = lev; state[1] = n; state[2] = which; }
}

__global__ void k_sort_keys(const int32_t *__restrict__ level, long long n, int nlevels, uint32_t *__restrict__ keys,
                            int32_t *__restrict__ vals, uint8_t *__restrict__ contam)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int l = level[i];
    if (l < 0) contam[i] = 0;   // never evaluated (nodata, on or behind a cycle): no contamination state
    keys[i] = l < 0 ? (uint32_t)nlevels : (uint32_t)l;
    vals[i] = (int32_t)i;
}

__global__ void k_level_offsets(const uint32_t *__restrict__ keys, long long n, int nlevels,
                                int32_t *__restrict__ level_off, const int32_t *__restrict__ order,
                                const uint8_t *__restrict__ inmask, uint8_t *__restrict__ omask)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    omask[i] = inmask[order[i]];   // in-flow mask in sweep order: read coalesced next to order[]
    uint32_t k = keys[i];
    if (k >= (uint32_t)nlevels) {
        if (i == 0 || keys[i - 1] < (uint32_t)nlevels) level_off[nlevels] = (int32_t)i;
        return;
    }
    if (i == 0 || keys[i - 1] != k) level_off[k] = (int32_t)i;
    if (i == n - 1) level_off[nlevels] = (int32_t)n;
}

// tailmask[c]: bit k-1 set iff upstream neighbour k of c lies in the tail (level >= tail_level)
__global__ void k_tailmask(const int32_t *__restrict__ order, int begin, int end, int ncol, int tail_level,
                           const uint8_t *__restrict__ inmask, const int32_t *__restrict__ level,
                           uint8_t *__restrict__ tailmask)
{
    int i = begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= end) return;
    int32_t c = order[i];
    unsigned m = inmask[c], tm = 0;
    while (m) {
        int k = __ffs(m);
        m &= m - 1;
        if (level[c + c_dy[k] * ncol + c_dx[k]] >= tail_level) tm |= 1u << (k - 1);
    }
    tailmask[c] = (uint8_t)tm;
}

// ---- sweep --------------------------------------------------------------------------------------
// One group of G lanes per cell, months j = g, g+G, ... on the lanes.  The cell's own row and its
// first upstream row are requested together (two independent loads in flight per lane); the rarer
// further upstream rows follow.  Additions run in TauDEM's k = 1..8 order, so the sum is deterministic.
// CG: read through L2 only (rows written by other warps during the same launch).
template <typename T, bool CG>
__device__ __forceinline__ T ldv(const T *p)
{
    return CG ? __ldcg(p) : *p;
}

template <typename T>
__device__ __forceinline__ const T *up_row(const T *A, int32_t c, int k, int ncol, int ld)
{
    return A + (size_t)(c + c_dy[k] * ncol + c_dx[k]) * ld;
}

template <typename T, int G, bool CG>
__device__ __forceinline__ void accum_cell(int32_t c, int g, int ncol, unsigned m, T *A, int ld, int nmonth)
{
    T *row = A + (size_t)c * ld;
    const T *u0 = up_row(A, c, __ffs(m), ncol, ld);
    const unsigned rest = m & (m - 1);
    for (int j = g; j < nmonth; j += G) {
        T own = ldv<T, CG>(row + j), a0 = ldv<T, CG>(u0 + j);
        T acc = own + a0;
        unsigned mm = rest;
        while (mm) {
            acc += ldv<T, CG>(up_row(A, c, __ffs(mm), ncol, ld) + j);
            mm &= mm - 1;
        }
        row[j] = acc;
    }
}

// two independent cells of one level interleaved: four loads in flight per lane
template <typename T, int G, bool CG>
__device__ __forceinline__ void accum_cell2(int32_t c0, unsigned m0, int32_t c1, unsigned m1, int g, int ncol, T *A,
                                            int ld, int nmonth)
{
    T *row0 = A + (size_t)c0 * ld, *row1 = A + (size_t)c1 * ld;
    const T *u0 = up_row(A, c0, __ffs(m0), ncol, ld), *u1 = up_row(A, c1, __ffs(m1), ncol, ld);
    const unsigned rest0 = m0 & (m0 - 1), rest1 = m1 & (m1 - 1);
    for (int j = g; j < nmonth; j += G) {
        T own0 = ldv<T, CG>(row0 + j), a0 = ldv<T, CG>(u0 + j);
        T own1 = ldv<T, CG>(row1 + j), a1 = ldv<T, CG>(u1 + j);
        T acc0 = own0 + a0, acc1 = own1 + a1;
        unsigned mm = rest0;
        while (mm) {
            acc0 += ldv<T, CG>(up_row(A, c0, __ffs(mm), ncol, ld) + j);
            mm &= mm - 1;
        }
        mm = rest1;
        while (mm) {
            acc1 += ldv<T, CG>(up_row(A, c1, __ffs(mm), ncol, ld) + j);
            mm &= mm - 1;
        }
        row0[j] = acc0;
        row1[j] = acc1;
    }
}

// ---- chain segments ---------------------------------------------------------------------------------
// Most cells below the cut have exactly one upstream cell that is not a source (level 0): their sum is the
// predecessor's sum plus their own weight (plus source weights).  A level sweep re-reads that predecessor row from
// DRAM one launch after it was written: 8 of the 24 bytes a non-source cell moves per month.  Such cells are
// ABSORBED into the lane group that computes their predecessor: a segment is a head cell (computed like in
// k_accum_level) followed by up to SEG_K - 1 absorbed cells along the flow path, the running sum staying in
// registers.  Segments are cut where the level index is 1 modulo SEG_K, so a segment never exceeds SEG_K cells and
// needs no position bookkeeping; a segment is launched with its head's level, which is never later than the level
// any of its cells had, so every dependency of every later launch is still met.  Additions keep TauDEM's k order
// (the predecessor's value enters at its own k), so the sums are the same bits as the level sweep's.
constexpr int SEG_K = 2;
constexpr int SEG_PF = 8192;      // descriptors are prefetched into L2 this many segments ahead
struct SegDesc {
    int32_t head;
    uint8_t len, hmask;
    uint16_t step[SEG_K - 1];      // per absorbed cell: (flow direction of the previous cell - 1) | in-flow mask << 3
};
static_assert(sizeof(SegDesc) % 4 == 0 && sizeof(SegDesc) >= 4 + 2 * SEG_K, "SegDesc is read as 32-bit words");
constexpr int SEG_NW = (int)(sizeof(SegDesc) / 4);

// absorbable: level in [2, seg_levels), exactly one upstream cell above level 0, and not at a segment boundary
__device__ __forceinline__ bool seg_absorbed(int32_t c, int lev, int seg_levels, int ncol, const uint8_t *__restrict__ inmask,
                                             const int32_t *__restrict__ level)
{
    if (lev < 2 || lev >= seg_levels || (lev - 1) % SEG_K == 0) return false;
    unsigned m = inmask[c];
    int nhigh = 0;
    while (m) {
        int k = __ffs(m);
        m &= m - 1;
        nhigh += level[c + c_dy[k] * ncol + c_dx[k]] >= 1;
    }
    return nhigh == 1;
}

// flag[i] = 1 iff order[i] (i in [begin, end)) starts a segment
__global__ void k_seg_flag(const int32_t *__restrict__ order, int begin, int end, int seg_levels, int ncol,
                           const uint8_t *__restrict__ inmask, const int32_t *__restrict__ level, int32_t *__restrict__ flag)
{
    int i = begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= end) return;
    const int32_t c = order[i];
    flag[i - begin] = seg_absorbed(c, level[c], seg_levels, ncol, inmask, level) ? 0 : 1;
}

// heads write their descriptor at the slot the scan gave them
__global__ void k_seg_build(const int32_t *__restrict__ order, int begin, int end, int seg_levels, int ncol,
                            const uint8_t *__restrict__ inmask, const int32_t *__restrict__ level,
                            const int32_t *__restrict__ receiver, const int32_t *__restrict__ slot,
                            SegDesc *__restrict__ segs, unsigned long long *ncells)
{
    int i = begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= end) return;
    if (slot[i - begin] == (i > begin ? slot[i - begin - 1] : 0)) return;      // inclusive scan of the head flags: no step, no head
    int32_t c = order[i];
    SegDesc d;
    d.head = c; d.hmask = inmask[c];
    int len = 1;
    for (int s = 0; s < SEG_K - 1; ++s) d.step[s] = 0;
    while (len < SEG_K) {
        const int32_t r = receiver[c];
        if (r < 0) break;
        const int lr = level[r];
        if (!seg_absorbed(r, lr, seg_levels, ncol, inmask, level)) break;
        // the cell r absorbs from is its only upstream cell above level 0: that must be c (c is above level 0)
        const int dr = r / ncol - c / ncol, dc = r % ncol - c % ncol;
        int kd = 0;
        for (int k = 1; k <= 8; ++k) if (c_dy[k] == dr && c_dx[k] == dc) kd = k;
        d.step[len - 1] = (uint16_t)((kd - 1) | (inmask[r] << 3));
        c = r;
        ++len;
    }
    d.len = (uint8_t)len;
    segs[slot[i - begin] - 1] = d;
    atomicAdd(ncells, (unsigned long long)len);
}

// One lane group per segment, months on the lanes.  Everything that does not depend on earlier launches (the
// descriptor, the weights of all cells of the segment, the source rows flowing into its absorbed cells) is requested
// before griddepcontrol.wait, all loads first and the additions afterwards so that a lane has up to 3 SEG_K rows in
// flight; only the head's upstream rows wait for the previous level.
template <typename T, int G>
__global__ void __launch_bounds__(256, 8) k_accum_segments(const SegDesc *__restrict__ segs, int begin, int end, int ncol, T *A,
                                                           int ld, int nmonth)
{
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int gid = (blockIdx.x * blockDim.x + threadIdx.x) / G, g = threadIdx.x % G;
    const int si = begin + gid;
    const bool live = si < end;
    // descriptors are read once, in order: pull the ones a later block of this launch will want into L2 now
    if (threadIdx.x == 0 && si + SEG_PF < end) asm volatile("prefetch.global.L2 [%0];" ::"l"(segs + si + SEG_PF));
    int32_t cell[SEG_K];
    unsigned st[SEG_K];          // per absorbed cell: (kd - 1) | in-flow mask << 3; st[0]: the head's in-flow mask
    int len = 0;
    if (live) {
        const int32_t *w = reinterpret_cast<const int32_t *>(segs + si);
        unsigned wd[SEG_NW];                               // head | len, hmask, step[0] | step[1], step[2] | ...
#pragma unroll
        for (int q = 0; q < SEG_NW; ++q) wd[q] = (unsigned)__ldg(w + q);
        cell[0] = (int32_t)wd[0];
        len = wd[1] & 0xff;
        st[0] = (wd[1] >> 8) & 0xff;
#pragma unroll
        for (int i = 1; i < SEG_K; ++i) st[i] = (i & 1) ? wd[(i + 1) / 2] >> 16 : wd[(i + 2) / 2] & 0xffff;
#pragma unroll
        for (int i = 1; i < SEG_K; ++i) {
            const int kd = (st[i] & 7) + 1;
            cell[i] = cell[i - 1] + c_dy[kd] * ncol + c_dx[kd];
        }
    }
    for (int jb = 0; jb < nmonth; jb += G) {
        const int j = jb + g;
        const bool on = live && j < nmonth;
        // phase 1, loads only: own weights; per absorbed cell the first source row below the predecessor's k and the
        // first one above it (sources are level-0 cells: their rows are never written)
        T low[SEG_K], s1[SEG_K], h1[SEG_K];
#pragma unroll
        for (int i = 0; i < SEG_K; ++i) {
            low[i] = T(0);
            if (on && i < len) low[i] = __ldcs(A + (size_t)cell[i] * ld + j);
        }
#pragma unroll
        for (int i = 1; i < SEG_K; ++i) {
            s1[i] = T(0); h1[i] = T(0);
            if (on && i < len) {
                const unsigned kp = (((st[i] & 7) + 4) & 7) + 1, m = st[i] >> 3;      // predecessor seen from this cell
                const unsigned lo = m & ((1u << (kp - 1)) - 1), hi = m & ~((1u << kp) - 1);
                if (lo) s1[i] = __ldcs(up_row(A, cell[i], __ffs(lo), ncol, ld) + j);
                if (hi) h1[i] = __ldcs(up_row(A, cell[i], __ffs(hi), ncol, ld) + j);
            }
        }
        if (jb == 0) asm volatile("griddepcontrol.wait;" ::: "memory");
        if (!on) continue;
        // the head like any cell of a level sweep (its upstream rows were written by earlier launches) ...
        T acc = low[0];
        {
            // up to three upstream rows requested together (a confluence has two or three), added in k order
            unsigned mm = st[0];
            const unsigned m0 = mm; mm &= mm - 1;
            const unsigned m1 = mm; mm &= mm - 1;
            const unsigned m2 = mm; mm &= mm - 1;
            T a0 = T(0), a1 = T(0), a2 = T(0);
            if (m0) a0 = __ldcg(up_row(A, cell[0], __ffs(m0), ncol, ld) + j);
            if (m1) a1 = __ldcg(up_row(A, cell[0], __ffs(m1), ncol, ld) + j);
            if (m2) a2 = __ldcg(up_row(A, cell[0], __ffs(m2), ncol, ld) + j);
            if (m0) acc += a0;
            if (m1) acc += a1;
            if (m2) acc += a2;
            while (mm) {
                acc += __ldcg(up_row(A, cell[0], __ffs(mm), ncol, ld) + j);
                mm &= mm - 1;
            }
        }
        A[(size_t)cell[0] * ld + j] = acc;
        // ... then the running sum down the segment, additions in k order: sources below the predecessor, the
        // predecessor (in registers), sources above it
#pragma unroll
        for (int i = 1; i < SEG_K; ++i) {
            if (i < len) {
                const unsigned kp = (((st[i] & 7) + 4) & 7) + 1, m = st[i] >> 3;
                unsigned lo = m & ((1u << (kp - 1)) - 1), hi = m & ~((1u << kp) - 1);
                T v = low[i];
                if (lo) {
                    v += s1[i];
                    lo &= lo - 1;
                    while (lo) {                 // a second source below the predecessor: rare
                        v += __ldcs(up_row(A, cell[i], __ffs(lo), ncol, ld) + j);
                        lo &= lo - 1;
                    }
                }
                acc = v + acc;
                if (hi) {
                    acc += h1[i];
                    hi &= hi - 1;
                    while (hi) {
                        acc += __ldcs(up_row(A, cell[i], __ffs(hi), ncol, ld) + j);
                        hi &= hi - 1;
                    }
                }
                A[(size_t)cell[i] * ld + j] = acc;
            }
        }
    }
}

// one wide level per launch, two cells per lane group (cells of a level never feed each other).
// Launched with programmatic stream serialization: the CTAs of level l+1 may become resident while level
// l drains; they read their topology and pull their own rows (weights, never written by another level)
// into L2, and only then wait for level l to complete and be visible (griddepcontrol.wait).
template <typename T, int G>
__global__ void __launch_bounds__(256) k_accum_level(const int32_t *__restrict__ order, int begin, int end, int ncol,
                                                     const uint8_t *__restrict__ omask, T *A, int ld, int nmonth)
{
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int gid = (blockIdx.x * blockDim.x + threadIdx.x) / G, g = threadIdx.x % G;
    const int i = begin + 2 * gid;
    int32_t c0 = -1, c1 = -1;
    unsigned m0 = 0, m1 = 0;
    if (i < end) {
        c0 = order[i];
        m0 = omask[i];
        if (i + 1 < end) { c1 = order[i + 1]; m1 = omask[i + 1]; }
        const int rowbytes = nmonth * (int)sizeof(T);
        if (g * 128 < rowbytes) {
            if (m0) asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char *>(A + (size_t)c0 * ld) + g * 128));
            if (m1) asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char *>(A + (size_t)c1 * ld) + g * 128));
        }
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (i >= end) return;
    if (m0 && m1) accum_cell2<T, G, true>(c0, m0, c1, m1, g, ncol, A, ld, nmonth);
    else if (m0) accum_cell<T, G, true>(c0, g, ncol, m0, A, ld, nmonth);
    else if (m1) accum_cell<T, G, true>(c1, g, ncol, m1, A, ld, nmonth);
}

// thin levels [lev0, lev1) inside one resident CTA with a block barrier between levels.  The topology
// reads of the NEXT level (offsets, first cell and mask per lane group) are issued before the current
// level is computed, and that cell's rows are prefetched into L2, so each level costs about one L2
// round trip plus the barrier instead of a chain of DRAM latencies.
template <typename T, int G>
__global__ void __launch_bounds__(1024) k_accum_thin(const int32_t *__restrict__ order,
                                                     const int32_t *__restrict__ level_off, int lev0, int lev1,
                                                     int ncol, const uint8_t *__restrict__ omask, T *A, int ld,
                                                     int nmonth)
{
    const int ngroups = blockDim.x / G, gid = threadIdx.x / G, g = threadIdx.x % G;
    int b = level_off[lev0], e = level_off[lev0 + 1];
    int e2 = lev0 + 2 <= lev1 ? level_off[lev0 + 2] : e;
    int32_t cn = -1;
    unsigned mn = 0;
    if (b + gid < e) { cn = order[b + gid]; mn = omask[b + gid]; }
    for (int l = lev0; l < lev1; ++l) {
        const int32_t c = cn;
        const unsigned m = mn;
        // next level: offsets two ahead, first cell of this group, L2 prefetch of its rows
        const int e3 = l + 3 <= lev1 ? level_off[l + 3] : e2;
        cn = -1; mn = 0;
        if (e + gid < e2) { cn = order[e + gid]; mn = omask[e + gid]; }
        if (c >= 0 && m) accum_cell<T, G, true>(c, g, ncol, m, A, ld, nmonth);
        for (int i = b + gid + ngroups; i < e; i += ngroups) {
            const int32_t ci = order[i];
            const unsigned mi = omask[i];
            if (mi) accum_cell<T, G, true>(ci, g, ncol, mi, A, ld, nmonth);
        }
        if (cn >= 0 && g < nmonth) {
            asm volatile("prefetch.global.L2 [%0];" ::"l"(A + (size_t)cn * ld + g));
            unsigned mm = mn;
            while (mm) {
                asm volatile("prefetch.global.L2 [%0];" ::"l"(up_row(A, cn, __ffs(mm), ncol, ld) + g));
                mm &= mm - 1;
            }
        }
        __syncthreads();
        b = e; e = e2; e2 = e3;
    }
}


// ---- upper forest by outlet basin ---------------------------------------------------------------
// Above a cut level the network is a forest of independent trees (one per outlet / sink).  Cells of
// the upper forest are regrouped by tree: corder[] = cells sorted by (root, level, id), so one CTA
// sweeps one tree level by level with only block barriers, and many trees run concurrently.
__global__ void k_root_init(const int32_t *__restrict__ order, int ub, int nu, const int32_t *__restrict__ receiver,
                            int32_t *__restrict__ root)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nu) return;
    int32_t c = order[ub + i], r = receiver[c];
    root[c] = r >= 0 ? r : c;
}

__global__ void k_root_jump(const int32_t *__restrict__ order, int ub, int nu, int32_t *root)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nu) return;
    int32_t c = order[ub + i];
    int32_t r = ((volatile int32_t *)root)[c];
    root[c] = ((volatile int32_t *)root)[r];      // racing readers only ever see an ancestor
}

__global__ void k_root_keys(const int32_t *__restrict__ order, int ub, int nu, const int32_t *__restrict__ root,
                            uint32_t *__restrict__ keys, int32_t *__restrict__ vals)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nu) return;
    int32_t c = order[ub + i];
    keys[i] = (uint32_t)root[c];
    vals[i] = c;
}

// cmask, level-start flags; component starts are appended to a list (any order)
__global__ void k_comp_flags(const int32_t *__restrict__ corder, const uint32_t *__restrict__ ckeys, int nu,
                             const int32_t *__restrict__ level, const uint8_t *__restrict__ inmask,
                             uint8_t *__restrict__ cmask, uint8_t *__restrict__ lstart, int32_t *__restrict__ comp_beg,
                             int *__restrict__ ncomp)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nu) return;
    int32_t c = corder[i];
    cmask[i] = inmask[c];
    bool newc = i == 0 || ckeys[i] != ckeys[i - 1];
    bool newl = newc || level[c] != level[corder[i - 1]];
    lstart[i] = newl ? (newc ? 2 : 1) : 0;
    if (newc) comp_beg[atomicAdd(ncomp, 1)] = i;
}

// upstream cell ids of every upper cell in k = 1..8 order (CSR): the sweep reads ready-made row indices
__global__ void k_comp_upcount(const uint8_t *__restrict__ cmask, int nu, int32_t *__restrict__ cnt)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nu) cnt[i] = __popc(cmask[i]);
}

// bit 31 of an entry marks a "fresh" upstream cell: one level below c, i.e. written in the level
// immediately before c's -- the only rows that sit on the sweep's latency chain
constexpr uint32_t UP_FRESH = 0x80000000u;
constexpr uint32_t UP_ID = 0x7fffffffu;

__global__ void k_comp_upfill(const int32_t *__restrict__ corder, const uint8_t *__restrict__ cmask, int nu, int ncol,
                              const int32_t *__restrict__ level, const int32_t *__restrict__ cup_off,
                              uint32_t *__restrict__ cup_idx)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nu) return;
    int32_t c = corder[i];
    const int lc = level[c];
    unsigned m = cmask[i];
    int q = cup_off[i];
    while (m) {
        int k = __ffs(m);
        m &= m - 1;
        int32_t u = c + c_dy[k] * ncol + c_dx[k];
        cup_idx[q++] = (uint32_t)u | (level[u] == lc - 1 ? UP_FRESH : 0u);
    }
}

// cnext[i] (defined at level starts): position of the next level start or component start
__global__ void k_comp_next(const uint8_t *__restrict__ lstart, int nu, int32_t *__restrict__ cnext)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nu || !lstart[i]) return;
    int j = i + 1;
    while (j < nu && !lstart[j]) ++j;
    cnext[i] = j;
}

// CTA = 8 compute warps (256 threads, 256/G lane groups) + 1 prefetch warp.  The compute warps walk
// one basin level by level (named barrier 1 between levels).  The upper phase is a latency chain as
// long as the longest river, so the per-level path is kept short: upstream cells come as ready-made
// ids (CSR cup_off / cup_idx, k order), and each lane group fetches the ids of its first cell of the
// NEXT level while it works on the current one; the prefetch warp runs PF_AHEAD cells ahead through
// the same list and pulls every row the sweep will touch into L2.
constexpr int PF_AHEAD = 384;

struct UpCell {           // a cell with up to 3 upstream ids (k order) preloaded; bit 31 of an id = fresh
    int32_t c, b, n;
    uint32_t u0, u1, u2;
};

__device__ __forceinline__ UpCell load_upcell(const int32_t *__restrict__ corder, const int32_t *__restrict__ cup_off,
                                              const uint32_t *__restrict__ cup_idx, int i, bool valid)
{
    UpCell x;
    x.c = -1; x.b = 0; x.n = 0; x.u0 = x.u1 = x.u2 = 0;
    if (valid) {
        x.c = corder[i];
        x.b = cup_off[i];
        x.n = cup_off[i + 1] - x.b;
        x.u0 = cup_idx[x.b];                       // n >= 1 above level 0
        x.u1 = x.n > 1 ? cup_idx[x.b + 1] : 0;
        x.u2 = x.n > 2 ? cup_idx[x.b + 2] : 0;
    }
    return x;
}

template <typename T, int G>
__device__ __forceinline__ void accum_upcell(const UpCell &x, const uint32_t *__restrict__ cup_idx, int g, T *A, int ld,
                                             int nmonth)
{
    T *row = A + (size_t)x.c * ld;
    const T *r0 = A + (size_t)(x.u0 & UP_ID) * ld, *r1 = A + (size_t)(x.u1 & UP_ID) * ld,
            *r2 = A + (size_t)(x.u2 & UP_ID) * ld;
    for (int j = g; j < nmonth; j += G) {
        T own = __ldcg(row + j), a0 = __ldcg(r0 + j), a1 = 0, a2 = 0;
        if (x.n > 1) a1 = __ldcg(r1 + j);
        if (x.n > 2) a2 = __ldcg(r2 + j);
        T acc = own + a0;
        if (x.n > 1) acc += a1;
        if (x.n > 2) acc += a2;
        for (int q = 3; q < x.n; ++q) acc += __ldcg(A + (size_t)(cup_idx[x.b + q] & UP_ID) * ld + j);
        row[j] = acc;
    }
}

template <typename T, int G>
__global__ void __launch_bounds__(288) k_accum_components(const int32_t *__restrict__ corder,
                                                          const int32_t *__restrict__ cup_off,
                                                          const uint32_t *__restrict__ cup_idx,
                                                          const int32_t *__restrict__ cnext,
                                                          const int2 *__restrict__ comps, int ncomp, int *ticket,
                                                          T *A, int ld, int nmonth)
{
    __shared__ int s_k, s_pos;
    const int ngroups = 256 / G, gid = threadIdx.x / G, g = threadIdx.x % G;
    const bool prefetcher = threadIdx.x >= 256;
    const int lane = threadIdx.x & 31;
    const int row_bytes = nmonth * (int)sizeof(T);
    for (;;) {
        if (threadIdx.x == 0) {
            s_k = atomicAdd(ticket, 1);
            s_pos = s_k < ncomp ? comps[s_k].x : 0;
        }
        __syncthreads();
        const int k = s_k;
        if (k >= ncomp) return;
        int pos = comps[k].x;
        const int ce = comps[k].y;
        if (prefetcher) {
            for (int i = pos + lane; i < ce; i += 32) {
                while (i > *(volatile int *)&s_pos + PF_AHEAD) __nanosleep(64);
                const char *r = reinterpret_cast<const char *>(A + (size_t)corder[i] * ld);
                for (int bb = 0; bb < row_bytes; bb += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(r + bb));
                for (int q = cup_off[i]; q < cup_off[i + 1]; ++q) {
                    const char *u = reinterpret_cast<const char *>(A + (size_t)(cup_idx[q] & UP_ID) * ld);
                    for (int bb = 0; bb < row_bytes; bb += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(u + bb));
                }
            }
        } else {
            int e = cnext[pos];
            UpCell nx = load_upcell(corder, cup_off, cup_idx, pos + gid, pos + gid < e);
            while (pos < ce) {
                const int e2 = e < ce ? cnext[e] : ce;          // issued before this level's work
                // cells pos+gid, pos+gid+ngroups, ... of this level; the ids of the following cell (or of
                // this group's first cell of the next level) are always requested before the current
                // cell is computed, so topology reads never sit on the critical path
                UpCell cur = nx;
                int i = pos + gid;
                if (cur.c < 0) {
                    nx = load_upcell(corder, cup_off, cup_idx, e + gid, e + gid < e2);
                } else {
                    for (;;) {
                        const int inext = i + ngroups;
                        const bool more = inext < e;
                        const UpCell nn = more ? load_upcell(corder, cup_off, cup_idx, inext, true)
                                               : load_upcell(corder, cup_off, cup_idx, e + gid, e + gid < e2);
                        accum_upcell<T, G>(cur, cup_idx, g, A, ld, nmonth);
                        if (!more) { nx = nn; break; }
                        cur = nn;
                        i = inext;
                    }
                }
                asm volatile("bar.sync 1, 256;" ::: "memory");
                pos = e; e = e2;
                if (threadIdx.x == 0) *(volatile int *)&s_pos = pos;
            }
        }
        __syncthreads();
    }
}

// Same walk for nmonth <= G (every lane owns exactly one month): the chain is taken off the memory
// system.  Each lane group (a) keeps the values of its first cell of the NEXT level in registers --
// own weight and every upstream value that is already final (not fresh) are requested one level
// ahead -- and (b) publishes the row it has just produced in a shared-memory slot tagged with the
// cell id, where the next level finds its fresh upstream rows.  What is left on the critical path of
// a level is a barrier, a tag search, a shared-memory read and an add.
template <typename T, int G>
__global__ void __launch_bounds__(288) k_accum_components_fw(const int32_t *__restrict__ corder,
                                                             const int32_t *__restrict__ cup_off,
                                                             const uint32_t *__restrict__ cup_idx,
                                                             const int32_t *__restrict__ cnext,
                                                             const int2 *__restrict__ comps, int ncomp, int *ticket,
                                                             T *A, int ld, int nmonth)
{
    constexpr int NG = 256 / G;
    __shared__ int s_k, s_pos;
    __shared__ int32_t fw_cell[2][NG];
    __shared__ T fw_row[2][NG][G];
    const int gid = threadIdx.x / G, g = threadIdx.x % G;
    const bool prefetcher = threadIdx.x >= 256;
    const bool live = g < nmonth;
    const int lane = threadIdx.x & 31;
    const int row_bytes = nmonth * (int)sizeof(T);
    for (;;) {
        if (threadIdx.x == 0) {
            s_k = atomicAdd(ticket, 1);
            s_pos = s_k < ncomp ? comps[s_k].x : 0;
        }
        if (threadIdx.x < 2 * NG) (&fw_cell[0][0])[threadIdx.x] = -1;
        __syncthreads();
        const int k = s_k;
        if (k >= ncomp) return;
        int pos = comps[k].x;
        const int ce = comps[k].y;
        if (prefetcher) {
            for (int i = pos + lane; i < ce; i += 32) {
                while (i > *(volatile int *)&s_pos + PF_AHEAD) __nanosleep(64);
                const char *r = reinterpret_cast<const char *>(A + (size_t)corder[i] * ld);
                for (int bb = 0; bb < row_bytes; bb += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(r + bb));
                for (int q = cup_off[i]; q < cup_off[i + 1]; ++q) {
                    const char *u = reinterpret_cast<const char *>(A + (size_t)(cup_idx[q] & UP_ID) * ld);
                    for (int bb = 0; bb < row_bytes; bb += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(u + bb));
                }
            }
        } else {
            int e = cnext[pos], par = 0;
            UpCell nx = load_upcell(corder, cup_off, cup_idx, pos + gid, pos + gid < e);
            T pown = 0, pv0 = 0, pv1 = 0, pv2 = 0;       // values of nx requested ahead (final ones only)
            if (nx.c >= 0 && live) {
                pown = __ldcg(A + (size_t)nx.c * ld + g);
                if (!(nx.u0 & UP_FRESH)) pv0 = __ldcg(A + (size_t)(nx.u0 & UP_ID) * ld + g);
                if (nx.n > 1 && !(nx.u1 & UP_FRESH)) pv1 = __ldcg(A + (size_t)(nx.u1 & UP_ID) * ld + g);
                if (nx.n > 2 && !(nx.u2 & UP_FRESH)) pv2 = __ldcg(A + (size_t)(nx.u2 & UP_ID) * ld + g);
            }
            while (pos < ce) {
                const int e2 = e < ce ? cnext[e] : ce;
                const UpCell cur = nx;
                const T own = pown, v0 = pv0, v1 = pv1, v2 = pv2;
                // ids of this group's first cell of the next level (values follow after this level's work)
                nx = load_upcell(corder, cup_off, cup_idx, e + gid, e + gid < e2);
                const int32_t *tags = fw_cell[par ^ 1];
                if (cur.c >= 0) {
                    // fresh upstream rows: shared-memory slot of the previous level if one holds them
                    auto fetch = [&](uint32_t u, T ahead) -> T {
                        if (!(u & UP_FRESH)) return ahead;
                        const int32_t id = (int32_t)(u & UP_ID);
                        int slot = -1;
#pragma unroll
                        for (int q = 0; q < NG; ++q)
                            if (tags[q] == id) slot = q;
                        if (slot >= 0) return fw_row[par ^ 1][slot][g];
                        return live ? __ldcg(A + (size_t)id * ld + g) : (T)0;
                    };
                    T acc = own + fetch(cur.u0, v0);
                    if (cur.n > 1) acc += fetch(cur.u1, v1);
                    if (cur.n > 2) acc += fetch(cur.u2, v2);
                    for (int q = 3; q < cur.n; ++q) {
                        const uint32_t u = cup_idx[cur.b + q];
                        acc += fetch(u, live && !(u & UP_FRESH) ? __ldcg(A + (size_t)(u & UP_ID) * ld + g) : (T)0);
                    }
                    if (live) A[(size_t)cur.c * ld + g] = acc;
                    fw_row[par][gid][g] = acc;
                }
                if (g == 0) fw_cell[par][gid] = cur.c;
                // remaining cells of this level (only when it is wider than the CTA has lane groups)
                for (int i = pos + gid + NG; i < e; i += NG)
                    accum_upcell<T, G>(load_upcell(corder, cup_off, cup_idx, i, true), cup_idx, g, A, ld, nmonth);
                // request next level's final values now; they land while the barrier is crossed
                pown = pv0 = pv1 = pv2 = 0;
                if (nx.c >= 0 && live) {
                    pown = __ldcg(A + (size_t)nx.c * ld + g);
                    if (!(nx.u0 & UP_FRESH)) pv0 = __ldcg(A + (size_t)(nx.u0 & UP_ID) * ld + g);
                    if (nx.n > 1 && !(nx.u1 & UP_FRESH)) pv1 = __ldcg(A + (size_t)(nx.u1 & UP_ID) * ld + g);
                    if (nx.n > 2 && !(nx.u2 & UP_FRESH)) pv2 = __ldcg(A + (size_t)(nx.u2 & UP_ID) * ld + g);
                }
                asm volatile("bar.sync 1, 256;" ::: "memory");
                pos = e; e = e2; par ^= 1;
                if (threadIdx.x == 0) *(volatile int *)&s_pos = pos;
            }
        }
        __syncthreads();
    }
}


// ---- upper network by river reach -----------------------------------------------------------------
// Above the cut most cells have exactly one upstream cell that is itself above the cut: they form
// reaches (maximal chains between confluences of upper branches).  Along a reach the only value on
// the dependency chain is the predecessor's, which the sweeping warp keeps in a register; everything
// else a cell adds is final before the reach starts.  Reaches are ordered by reach level (0 = fed from
// below the cut only, else 1 + max over the reaches that join at its head); one launch sweeps all
// reaches of a level, one lane group per reach.  The latency chain of the upper phase thus shrinks from
// "longest river in LEVELS x a barrier and an L2 round trip" to "sum over reach levels of the longest
// reach in CELLS x a few dependent adds".
constexpr uint32_t UP_PRED = 0x80000000u;     // rup_idx entry: the in-reach predecessor

__global__ void k_run_upos(const int32_t *__restrict__ uorder, int nu, int32_t *__restrict__ upos)
{
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < nu) upos[uorder[j]] = j;
}

// pred[j] = position of the only upper upstream cell, or -1 (reach head); npu[j] = number of upper upstream cells
__global__ void k_run_pred(const int32_t *__restrict__ uorder, int nu, int ncol, int cut,
                           const uint8_t *__restrict__ inmask, const int32_t *__restrict__ level,
                           const int32_t *__restrict__ upos, int32_t *__restrict__ pred, uint8_t *__restrict__ npu)
{
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nu) return;
    int32_t c = uorder[j];
    unsigned m = inmask[c];
    int cnt = 0, last = -1;
    while (m) {
        int k = __ffs(m);
        m &= m - 1;
        int32_t u = c + c_dy[k] * ncol + c_dx[k];
        if (level[u] >= cut) { ++cnt; last = upos[u]; }
    }
    npu[j] = (uint8_t)cnt;
    pred[j] = cnt == 1 ? last : -1;
}

__global__ void k_run_rank_init(const int32_t *__restrict__ pred, int nu, int32_t *__restrict__ head,
                                int32_t *__restrict__ dist)
{
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nu) return;
    head[j] = pred[j] < 0 ? j : pred[j];
    dist[j] = pred[j] < 0 ? 0 : 1;
}

// pointer jumping (ping-pong buffers): head -> head of head, dist accumulates
__global__ void k_run_rank_step(const int32_t *__restrict__ hs, const int32_t *__restrict__ ds, int nu,
                                int32_t *__restrict__ hd, int32_t *__restrict__ dd, int *__restrict__ changed)
{
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nu) return;
    int32_t h = hs[j], hh = hs[h];
    hd[j] = hh;
    dd[j] = ds[j] + (hh != h ? ds[h] : 0);
    if (hh != h) *changed = 1;
}

// reach level of the heads by monotone relaxation (values only grow; fixed point = the definition)
__global__ void k_run_level_step(const int32_t *__restrict__ uorder, int nu, int ncol, int cut,
                                 const uint8_t *__restrict__ inmask, const int32_t *__restrict__ level,
                                 const int32_t *__restrict__ upos, const int32_t *__restrict__ head,
                                 const uint8_t *__restrict__ npu, int32_t *rl, int *__restrict__ changed)
{
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nu || npu[j] < 2) return;             // heads fed by upper reaches only
    int32_t c = uorder[j];
    unsigned m = inmask[c];
    int best = 0;
    while (m) {
        int k = __ffs(m);
        m &= m - 1;
        int32_t u = c + c_dy[k] * ncol + c_dx[k];
        if (level[u] >= cut) best = max(best, 1 + ((volatile int32_t *)rl)[head[upos[u]]]);
    }
    if (best > rl[j]) { rl[j] = best; *changed = 1; }
}

__global__ void k_run_keys(const int32_t *__restrict__ head, const int32_t *__restrict__ dist,
                           const int32_t *__restrict__ rl, int nu, unsigned long long *__restrict__ keys,
                           int32_t *__restrict__ vals, int *__restrict__ overflow)
{
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nu) return;
    int32_t h = head[j];
    if (dist[j] >= (1 << 20)) *overflow = 1;     // a reach of a million cells does not fit the key
    keys[j] = ((unsigned long long)rl[h] << 48) | ((unsigned long long)h << 20) | (unsigned long long)dist[j];
    vals[j] = j;
}

// after the sort: cells in sweep order, reach starts flagged, upstream counts
__global__ void k_run_gather(const int32_t *__restrict__ sj, int nu, const int32_t *__restrict__ uorder,
                             const int32_t *__restrict__ dist, const uint8_t *__restrict__ inmask,
                             int32_t *__restrict__ rcell, int32_t *__restrict__ isstart, int32_t *__restrict__ cnt)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nu) return;
    int j = sj[i];
    int32_t c = uorder[j];
    rcell[i] = c;
    isstart[i] = dist[j] == 0;
    cnt[i] = __popc(inmask[c]);
}

__global__ void k_run_starts(const int32_t *__restrict__ isstart, const int32_t *__restrict__ rank, int nu,
                             int32_t *__restrict__ run_start)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nu && isstart[i]) run_start[rank[i] - 1] = i;      // rank = inclusive scan of isstart
}

__global__ void k_run_upfill(const int32_t *__restrict__ sj, int nu, int ncol, const int32_t *__restrict__ rcell,
                             const uint8_t *__restrict__ inmask, const int32_t *__restrict__ pred,
                             const int32_t *__restrict__ uorder, const int32_t *__restrict__ rup_off,
                             uint32_t *__restrict__ rup_idx)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nu) return;
    int32_t c = rcell[i];
    int pj = pred[sj[i]];
    int32_t pc = pj >= 0 ? uorder[pj] : -1;
    unsigned m = inmask[c];
    int q = rup_off[i];
    while (m) {
        int k = __ffs(m);
        m &= m - 1;
        int32_t u = c + c_dy[k] * ncol + c_dx[k];
        rup_idx[q++] = (uint32_t)u | (u == pc ? UP_PRED : 0u);
    }
}

// reach level of every reach (run) and the first reach of every level
__global__ void k_run_levels(const int32_t *__restrict__ run_start, int nr, const int32_t *__restrict__ sj,
                             const int32_t *__restrict__ rl, int32_t *__restrict__ run_rl)
{
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < nr) run_rl[r] = rl[sj[run_start[r]]];
}

// rows a cell contributes to the reach stream: none for a reach head, else 1 (own weight + the side
// inflows that precede the predecessor in k order) + one per side inflow after the predecessor
__global__ void k_run_rows(const int32_t *__restrict__ isstart, const int32_t *__restrict__ rup_off,
                           const uint32_t *__restrict__ rup_idx, int nu, int32_t *__restrict__ rows)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nu) return;
    int r = 0;
    if (!isstart[i]) {
        const int o = rup_off[i], n = rup_off[i + 1] - o;
        int q = 0;
        while (q < n && !(rup_idx[o + q] & UP_PRED)) ++q;
        r = n - q;                                  // 1 + entries after the predecessor
    }
    rows[i] = r;
}

// rowcell[r] = the cell whose sum is complete after stream row r, or -1
__global__ void k_run_rowcell(const int32_t *__restrict__ row_off, const int32_t *__restrict__ rcell, int nu,
                              int32_t *__restrict__ rowcell)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nu) return;
    const int b = row_off[i], e = row_off[i + 1];
    for (int r = b; r < e; ++r) rowcell[r] = r == e - 1 ? rcell[i] : -1;
}

__global__ void k_run_inv(const int32_t *__restrict__ sj, int nu, int32_t *__restrict__ inv)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nu) inv[sj[i]] = i;
}

// per reach, one 80-byte record: first/last stream row, head cell, number of head inflows; the inflow
// cells; and for every inflow the reach whose last cell it is (-1: below the cut, final before the sweep)
constexpr int RINFO_VEC = 5;
__global__ void k_run_pack(const int32_t *__restrict__ run_start, int nr, const int32_t *__restrict__ rcell,
                           const int32_t *__restrict__ rup_off, const uint32_t *__restrict__ rup_idx,
                           const int32_t *__restrict__ row_off, const int32_t *__restrict__ level, int cut,
                           const int32_t *__restrict__ upos, const int32_t *__restrict__ inv,
                           const int32_t *__restrict__ rank, int4 *__restrict__ rinfo)
{
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nr) return;
    const int i0 = run_start[r], i1 = run_start[r + 1];
    const int o = rup_off[i0], n = rup_off[i0 + 1] - o;
    int u[8], d[8];
    for (int q = 0; q < 8; ++q) {
        u[q] = 0; d[q] = -1;
        if (q < n) {
            u[q] = (int)(rup_idx[o + q] & UP_ID);
            if (level[u[q]] >= cut) d[q] = rank[inv[upos[u[q]]]] - 1;     // rank = inclusive scan of reach starts
        }
    }
    int4 *out = rinfo + (size_t)RINFO_VEC * r;
    out[0] = make_int4(row_off[i0], row_off[i1], rcell[i0], n);
    out[1] = make_int4(u[0], u[1], u[2], u[3]);
    out[2] = make_int4(u[4], u[5], u[6], u[7]);
    out[3] = make_int4(d[0], d[1], d[2], d[3]);
    out[4] = make_int4(d[4], d[5], d[6], d[7]);
}

// Phase 1 of the reach sweep, fully parallel: everything a non-head cell adds besides its predecessor is
// final once the levels below the cut are done.  Write it as a dense stream in sweep order, so that the
// serial phase reads consecutive memory (no index chasing, no TLB misses on its latency chain):
//   row_off[i]      : own + side inflows before the predecessor (summed in k order, as AreaD8 does)
//   row_off[i]+1... : the side inflows after the predecessor, one row each (their order matters)
// A running sum over the rows of a reach then reproduces the k-ordered additions bit for bit, since
// (own + a) + chain == chain + (own + a) in IEEE arithmetic.
// A lane group takes G consecutive cells: the sweep metadata is read one cell per lane (coalesced) and
// handed round by shuffles, so that only the row loads themselves see memory latency.
template <typename T, int G>
__global__ void __launch_bounds__(256) k_reach_gather(const int32_t *__restrict__ rcell,
                                                      const int32_t *__restrict__ rup_off,
                                                      const uint32_t *__restrict__ rup_idx,
                                                      const int32_t *__restrict__ row_off, int nu,
                                                      const T *__restrict__ A, T *__restrict__ S, int ld, int lds,
                                                      int nmonth)
{
    const int g = threadIdx.x % G;
    const unsigned gmask = G == 32 ? 0xffffffffu : (((1u << G) - 1u) << ((threadIdx.x % 32) / G * G));
    for (long long ib = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / G * G; ib < nu;
         ib += (long long)gridDim.x * blockDim.x) {
    int m_rb = 0, m_nr = 0, m_c = 0, m_o = 0, m_n = 0;
    uint32_t m_e0 = UP_PRED, m_e1 = UP_PRED, m_e2 = UP_PRED;
    if (ib + g < nu) {
        const int i = (int)ib + g;
        m_rb = row_off[i]; m_nr = row_off[i + 1] - m_rb;
        m_c = rcell[i]; m_o = rup_off[i]; m_n = rup_off[i + 1] - m_o;
    }
    if (m_nr > 0) {
        m_e0 = rup_idx[m_o];
        if (m_n > 1) m_e1 = rup_idx[m_o + 1];
        if (m_n > 2) m_e2 = rup_idx[m_o + 2];
    }
    const int cnt = (int)min((long long)G, nu - ib);
#pragma unroll 4
    for (int s = 0; s < cnt; ++s) {
        const int nr = __shfl_sync(gmask, m_nr, s, G);
        const int rb = __shfl_sync(gmask, m_rb, s, G);
        const int c = __shfl_sync(gmask, m_c, s, G);
        const int n = __shfl_sync(gmask, m_n, s, G);
        const int o = __shfl_sync(gmask, m_o, s, G);
        const uint32_t e0 = __shfl_sync(gmask, m_e0, s, G);
        const uint32_t e1 = __shfl_sync(gmask, m_e1, s, G);
        const uint32_t e2 = __shfl_sync(gmask, m_e2, s, G);
        if (nr == 0) continue;                    // reach head: summed from A by the sweep itself
        for (int j = g; j < nmonth; j += G) {
            T acc = __ldcs(A + (size_t)c * ld + j);
            const T v0 = (e0 & UP_PRED) ? (T)0 : __ldcs(A + (size_t)e0 * ld + j);
            const T v1 = (e1 & UP_PRED) ? (T)0 : __ldcs(A + (size_t)e1 * ld + j);
            const T v2 = (e2 & UP_PRED) ? (T)0 : __ldcs(A + (size_t)e2 * ld + j);
            int r = rb;
            bool after = false;
            auto put = [&](uint32_t e, T v) {
                if (e & UP_PRED) after = true;
                else if (!after) acc += v;
                else __stcs(S + (size_t)(++r) * lds + j, v);
            };
            put(e0, v0);
            if (n > 1) put(e1, v1);
            if (n > 2) put(e2, v2);
            for (int q = 3; q < n; ++q) {
                const uint32_t e = rup_idx[o + q];
                put(e, (e & UP_PRED) ? (T)0 : __ldcs(A + (size_t)e * ld + j));
            }
            __stcs(S + (size_t)rb * lds + j, acc);
        }
    }
    }
}

// Phase 2: one warp per reach, months on the lanes, the reaches of one reach level per launch.  The head
// is summed from A (its inflows are tails of earlier reaches or lie below the cut); after it the sweep
// is a running sum over the reach's rows of the stream, stored to A wherever a cell completes.  The
// stream (and the matching rowcell entries) arrives through a per-warp ring of bulk async copies, RS_ROWS
// rows per stage and RS_STAGES stages ahead, so the chain sees shared-memory latency only.
constexpr int RS_ROWS = 16, RS_STAGES = 4, RS_WARPS = 2;

template <typename T>
__global__ void __launch_bounds__(32 * RS_WARPS) k_reach_sweep(const int4 *__restrict__ rinfo,
                                                              const int32_t *__restrict__ rowcell, int r0, int r1,
                                                              const T *__restrict__ S, int lds, T *A, int ld,
                                                              int nmonth, int *ticket, unsigned *done, unsigned epoch)
{
    __shared__ alignas(128) T sbuf[RS_WARPS][RS_STAGES][RS_ROWS * 32];
    __shared__ alignas(16) int32_t scell[RS_WARPS][RS_STAGES][RS_ROWS];
    __shared__ alignas(8) uint64_t full[RS_WARPS][RS_STAGES];
    const int w = threadIdx.x / 32, lane = threadIdx.x % 32;
    // dataflow mode (ticket != NULL): all reach levels in one launch.  Reaches are sorted by level, so every
    // reach a head waits for has a smaller index; taking the index from a ticket makes "smaller index"
    // mean "already running", whatever order the CTAs were dispatched in.
    int run = blockIdx.x * RS_WARPS + w;
    if (ticket) {
        if (lane == 0) run = atomicAdd(ticket, 1);
        run = __shfl_sync(0xffffffffu, run, 0);
    }
    run += r0;
    if (run >= r1) return;                        // warps never meet at a block barrier
    if (lane == 0) {
        for (int i = 0; i < RS_STAGES; ++i) mbar_init(&full[w][i], 1);
        mbar_fence_init();
    }
    __syncwarp();
    const int4 *ri = rinfo + (size_t)RINFO_VEC * run;
    const int4 h0 = __ldg(ri), h1 = __ldg(ri + 1), h2 = __ldg(ri + 2);
    const int rbeg = h0.x, rend = h0.y, c0 = h0.z, n0 = h0.w;
    const int up[8] = {h1.x, h1.y, h1.z, h1.w, h2.x, h2.y, h2.z, h2.w};
    int mydep = -1;                               // lane q < 8 watches the reach that ends in inflow q
    if (ticket) {
        const int4 d0 = __ldg(ri + 3), d1 = __ldg(ri + 4);
        const int dep[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
#pragma unroll
        for (int q = 0; q < 8; ++q)
            if (lane == q) mydep = dep[q];
    }
    const int b0 = rbeg / RS_ROWS, nblk = rend > rbeg ? (rend + RS_ROWS - 1) / RS_ROWS - b0 : 0;
    // rows narrower than 32 values are contiguous in the stream: one copy per stage, row pitch lds in
    // shared memory too; wider rows are swept 32 months at a time, one copy per row
    const bool dense = lds <= 32;
    const int spitch = dense ? lds : 32;
    const size_t apitch = (size_t)ld * sizeof(T);
    int done_blk = 0;                             // stream blocks consumed so far over all passes (ring position)
    for (int j0 = 0; j0 < nmonth; j0 += 32) {
        const int j = j0 + lane;
        const bool live = j < nmonth;
        char *Aj = reinterpret_cast<char *>(A + (live ? j : 0));
        const uint32_t segbytes = (uint32_t)min(32, lds - j0) * (uint32_t)sizeof(T);
        auto issue = [&](int k, int pos) {        // stream block b0+k into ring slot pos % RS_STAGES (lane 0 only)
            const int st = pos % RS_STAGES;
            const size_t row = (size_t)(b0 + k) * RS_ROWS;
            mbar_arrive_expect_tx(&full[w][st], RS_ROWS * segbytes + RS_ROWS * 4);
            if (dense) {
                bulk_g2s(&sbuf[w][st][0], S + row * lds, RS_ROWS * segbytes, &full[w][st]);
            } else {
                for (int r = 0; r < RS_ROWS; ++r)
                    bulk_g2s(&sbuf[w][st][r * 32], S + (row + r) * lds + j0, segbytes, &full[w][st]);
            }
            bulk_g2s(&scell[w][st][0], rowcell + row, RS_ROWS * 4, &full[w][st]);
        };
        if (lane == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            for (int k = 0; k < min(nblk, RS_STAGES); ++k) issue(k, done_blk + k);
        }
        // reach head
        if (ticket && j0 == 0) {
            if (mydep >= 0) {
                unsigned seen;
                do {
                    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(done + mydep) : "memory");
                } while (seen != epoch);
            }
            __syncwarp();
        }
        T acc = 0;
        if (live) {
            T v[8];
            acc = __ldcg(A + (size_t)c0 * ld + j);
#pragma unroll
            for (int q = 0; q < 8; ++q) v[q] = q < n0 ? __ldcg(A + (size_t)up[q] * ld + j) : (T)0;
#pragma unroll
            for (int q = 0; q < 8; ++q)
                if (q < n0) acc += v[q];
            A[(size_t)c0 * ld + j] = acc;
        }
        for (int k = 0; k < nblk; ++k) {
            const int pos = done_blk + k, st = pos % RS_STAGES;
            mbar_wait(&full[w][st], (pos / RS_STAGES) & 1);
            const int base = (b0 + k) * RS_ROWS;
            const int lo = max(rbeg - base, 0), hi = min(rend - base, RS_ROWS);
            const T *sb = &sbuf[w][st][lane];
            const int4 *sc = reinterpret_cast<const int4 *>(&scell[w][st][0]);
            auto step = [&](T v, int32_t c) {
                acc += v;
                if (c >= 0 && live) *reinterpret_cast<T *>(Aj + (size_t)c * apitch) = acc;
            };
            if (lo == 0 && hi == RS_ROWS) {
#pragma unroll
                for (int q = 0; q < RS_ROWS / 4; ++q) {
                    const int4 c4 = sc[q];
                    const T v0 = sb[(4 * q + 0) * spitch], v1 = sb[(4 * q + 1) * spitch],
                            v2 = sb[(4 * q + 2) * spitch], v3 = sb[(4 * q + 3) * spitch];
                    step(v0, c4.x); step(v1, c4.y); step(v2, c4.z); step(v3, c4.w);
                }
            } else {
                for (int r = lo; r < hi; ++r) step(sb[r * spitch], scell[w][st][r]);
            }
            __syncwarp();
            if (lane == 0 && k + RS_STAGES < nblk) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                issue(k + RS_STAGES, pos + RS_STAGES);
            }
        }
        done_blk += nblk;
    }
    if (ticket) {                                 // publish: the last cell of this reach is in A
        __syncwarp();
        if (lane == 0) {
            __threadfence();
            asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(done + run), "r"(epoch) : "memory");
        }
    }
}

// dataflow tail: cells order[begin, end) are claimed in order, one ticket per sub-warp group and
// CHUNK cells per ticket; a cell waits only for tail upstream cells (tailmask) to publish their
// epoch in done[], which are always held by earlier tickets => no deadlock.
template <typename T, int G>
__global__ void __launch_bounds__(256) k_accum_dataflow(const int32_t *__restrict__ order, int begin, int end,
                                                        int ncol, const uint8_t *__restrict__ omask,
                                                        const uint8_t *__restrict__ tailmask, T *A, int ld, int nmonth,
                                                        unsigned *done, unsigned epoch, int *ticket)
{
    constexpr int CHUNK = 4;
    const int g = threadIdx.x % G;
    const unsigned gmask = G == 32 ? 0xffffffffu : (((1u << G) - 1u) << ((threadIdx.x % 32) / G * G));
    for (;;) {
        int t0 = 0;
        if (g == 0) t0 = atomicAdd(ticket, CHUNK);
        t0 = __shfl_sync(gmask, t0, (threadIdx.x % 32) / G * G);
        int i0 = begin + t0;
        if (i0 >= end) return;
        int i1 = min(end, i0 + CHUNK);
        for (int i = i0; i < i1; ++i) {
            int32_t c = order[i];
            unsigned m = omask[i], tm = tailmask[c];
            if (g == 0) {
                while (tm) {
                    int k = __ffs(tm);
                    tm &= tm - 1;
                    const volatile unsigned *f = done + (c + c_dy[k] * ncol + c_dx[k]);
                    while (*f != epoch) { }
                }
                __threadfence();
            }
            __syncwarp(gmask);
            if (m) accum_cell<T, G, true>(c, g, ncol, m, A, ld, nmonth);
            __syncwarp(gmask);
            if (g == 0) {
                __threadfence();
                ((volatile unsigned *)done)[c] = epoch;
            }
        }
    }
}

// contaminated / unreached / nodata cells -> NaN
template <typename T>
__global__ void k_mask(const int32_t *__restrict__ level, const uint8_t *__restrict__ contam, int use_contam,
                       long long ncell, T *A, int ld, int nmonth)
{
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long c = idx / nmonth;
    int j = (int)(idx % nmonth);
    if (c >= ncell) return;
    if (level[c] < 0 || (use_contam && contam[c])) A[(size_t)c * ld + j] = (T)NAN;
}

// planes [nmonth][ncell] (double) -> cell-major [ncell][ld] (T) and back
template <typename T>
__global__ void k_planes_to_cells(const double *__restrict__ in, long long ncell, int nmonth, int ld, T *__restrict__ out)
{
    __shared__ double s[32][33];
    long long c0 = (long long)blockIdx.x * 32;
    int m0 = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += 8) {
        long long c = c0 + threadIdx.x;
        int m = m0 + j;
        s[j][threadIdx.x] = (c < ncell && m < nmonth) ? in[(size_t)m * ncell + c] : 0.0;
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += 8) {
        long long c = c0 + j;
        int m = m0 + threadIdx.x;
        if (c < ncell && m < ld) out[(size_t)c * ld + m] = (T)s[threadIdx.x][j];
    }
}

template <typename T>
__global__ void k_cells_to_planes(const T *__restrict__ in, long long ncell, int nmonth, int ld, double *__restrict__ out)
{
    __shared__ double s[32][33];
    long long c0 = (long long)blockIdx.x * 32;
    int m0 = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += 8) {
        long long c = c0 + j;
        int m = m0 + threadIdx.x;
        s[j][threadIdx.x] = (c < ncell && m < nmonth) ? (double)in[(size_t)c * ld + m] : 0.0;
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += 8) {
        long long c = c0 + threadIdx.x;
        int m = m0 + j;
        if (c < ncell && m < nmonth) out[(size_t)m * ncell + c] = s[threadIdx.x][j];
    }
}

// Weight[index$cells] <- QS[i,]  (Routing:109-113); src < 0 marks an overwritten duplicate
// months [t0, t0 + nt) of the ntime-month series go to columns 0..nt-1 of the sweep buffer
template <typename T>
__global__ void k_scatter(const double *__restrict__ QS, const int32_t *__restrict__ src, int nsub, int ntime, int t0,
                          int nt, T *A, int ld)
{
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    int s = idx / nt, t = idx % nt;
    if (s >= nsub || src[s] < 0) return;
    A[(size_t)src[s] * ld + t] = (T)QS[(size_t)s * ntime + t0 + t];
}

// max(values[coverage_fraction == 1], na.rm = TRUE) per polygon and month (Routing:119-121).
// Block = one polygon x 32 months; warps stride over the polygon's cells, then a shared-memory max.
template <typename T>
__global__ void __launch_bounds__(256) k_polygon_max(const int32_t *__restrict__ poff, const int32_t *__restrict__ pcells,
                                                     const int32_t *__restrict__ level, const uint8_t *__restrict__ contam,
                                                     int use_contam, const T *__restrict__ A, int ld, int ntime,
                                                     int t0, int nt, double *__restrict__ QR)
{
    __shared__ double sm[8][32];
    const int s = blockIdx.x, t = blockIdx.y * 32 + (threadIdx.x & 31), w = threadIdx.x >> 5;
    double m = -INFINITY;
    if (t < nt)
        for (int q = poff[s] + w; q < poff[s + 1]; q += 8) {
            int32_t c = pcells[q];
            if (level[c] < 0 || (use_contam && contam[c])) continue;
            double v = (double)A[(size_t)c * ld + t];
            if (v > m) m = v;
        }
    sm[w][threadIdx.x & 31] = m;
    __syncthreads();
    if (w == 0 && t < nt) {
#pragma unroll
        for (int k = 1; k < 8; ++k) m = fmax(m, sm[k][threadIdx.x]);
        QR[(size_t)s * ntime + t0 + t] = m;
    }
}


constexpr int COMP_WIDTH = 16384;     // per-basin sweep starts where every later level is at most this wide
constexpr int THIN_LIMIT = 2048;      // plan peeling: frontiers this small are peeled inside one CTA
constexpr int TAIL_WIDTH = 512;       // dataflow sweep: the tail starts where every later level is at most this wide

}  // namespace


static void plan_free(wfa_plan *p)
{
    if (!p) return;
    cudaSetDevice(p->device);
    DevBuf *bufs[] = {&p->receiver, &p->inmask, &p->omask, &p->indeg, &p->level, &p->order, &p->level_off, &p->contam,
                      &p->tailmask, &p->done, &p->ticket, &p->work, &p->work2, &p->qs, &p->src, &p->poff, &p->pcells,
                      &p->qr, &p->corder, &p->cmask, &p->cnext, &p->comps, &p->cup_off, &p->cup_idx, &p->rcell, &p->rup_off,
                      &p->rup_idx, &p->run_start, &p->row_off, &p->rowcell, &p->rstream, &p->rinfo, &p->rdone, &p->segs};
    { FreeBatch fb; for (DevBuf *d : bufs) d->release(); }
    if (p->ev0) cudaEventDestroy(p->ev0);
    if (p->ev1) cudaEventDestroy(p->ev1);
    if (p->evm) cudaEventDestroy(p->evm);
    if (p->own_stream) cudaStreamDestroy(p->own_stream);
    delete p;
}

static int plan_build_components(wfa_plan *p);
static int plan_build_runs(wfa_plan *p);
static int plan_build_segments(wfa_plan *p);

static int plan_build(wfa_plan *p, const int16_t *dir_dev)
{
    const long long n = p->ncell;
    const int nb = ceil_div(n, 256);
    int rc;
    if ((rc = p->receiver.ensure(n * 4))) return rc;
    if ((rc = p->inmask.ensure(n))) return rc;
    if ((rc = p->indeg.ensure(n * 4))) return rc;
    if ((rc = p->level.ensure(n * 4))) return rc;
    if ((rc = p->contam.ensure(n))) return rc;
    DevBuf cnt, f0, f1, state;
    auto cleanup = [&]() { cnt.release(); f0.release(); f1.release(); state.release(); };
#define PCK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { gr2m_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); cleanup(); return e_ == cudaErrorMemoryAllocation ? GR2M_ERR_NOMEM : GR2M_ERR_CUDA; } } while (0)
#define PRC(call) do { int r_ = (call); if (r_) { cleanup(); return r_; } } while (0)
    PRC(cnt.ensure(n * 4));
    PRC(f0.ensure(n * 4));
    PRC(f1.ensure(n * 4));
    PRC(state.ensure(4 * sizeof(int)));
    cudaStream_t st = p->stream;
    gr2m_count_launch(), k_topology<<<nb, 256, 0, st>>>(dir_dev, p->nrow, p->ncol, p->receiver.as<int32_t>(), p->inmask.as<uint8_t>(),
                                   p->indeg.as<int32_t>(), cnt.as<int32_t>(), p->contam.as<uint8_t>(),
                                   p->level.as<int32_t>());
    PCK(cudaGetLastError());
    int *d_state = state.as<int>();   // [0] level [1] size [2] which [3] scratch counter
    PCK(cudaMemsetAsync(d_state, 0, 4 * sizeof(int), st));
    gr2m_count_launch(), k_frontier_init<<<nb, 256, 0, st>>>(cnt.as<int32_t>(), n, f0.as<int32_t>(), d_state + 1);
    PCK(cudaGetLastError());
    int h[4] = {0, 0, 0, 0};
    PCK(cudaMemcpyAsync(h, d_state, sizeof(h), cudaMemcpyDeviceToHost, st));
    PCK(cudaStreamSynchronize(st));
    long long reached = 0;
    while (h[1] > 0) {
        if (h[1] > THIN_LIMIT) {
            reached += h[1];
            int32_t *cur = h[2] ? f1.as<int32_t>() : f0.as<int32_t>(), *nxt = h[2] ? f0.as<int32_t>() : f1.as<int32_t>();
            PCK(cudaMemsetAsync(d_state + 3, 0, sizeof(int), st));
            gr2m_count_launch(), k_peel<<<ceil_div(h[1], 256), 256, 0, st>>>(cur, h[1], h[0], p->ncol, p->receiver.as<int32_t>(),
                                                         p->inmask.as<uint8_t>(), cnt.as<int32_t>(),
                                                         p->contam.as<uint8_t>(), p->level.as<int32_t>(), nxt, d_state + 3);
            PCK(cudaGetLastError());
            int nn = 0;
            PCK(cudaMemcpyAsync(&nn, d_state + 3, sizeof(int), cudaMemcpyDeviceToHost, st));
            PCK(cudaStreamSynchronize(st));
            h[0] += 1; h[1] = nn; h[2] ^= 1;
        } else {
            PCK(cudaMemcpyAsync(d_state, h, 3 * sizeof(int), cudaMemcpyHostToDevice, st));
            gr2m_count_launch(), k_peel_thin<<<1, 1024, 0, st>>>(f0.as<int32_t>(), f1.as<int32_t>(), d_state, THIN_LIMIT, p->ncol,
                                            p->receiver.as<int32_t>(), p->inmask.as<uint8_t>(), cnt.as<int32_t>(),
                                            p->contam.as<uint8_t>(), p->level.as<int32_t>());
            PCK(cudaGetLastError());
            PCK(cudaMemcpyAsync(h, d_state, 3 * sizeof(int), cudaMemcpyDeviceToHost, st));
            PCK(cudaStreamSynchronize(st));
        }
    }
    p->nlevels = h[0];
    cleanup();
    // order = stable radix sort of cell ids by level (cells never reached sort last and are cut off)
    DevBuf keys, keys2, vals, temp;
    auto cleanup2 = [&]() { keys.release(); keys2.release(); vals.release(); temp.release(); };
#undef PCK
#undef PRC
#define PCK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { gr2m_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); cleanup2(); return e_ == cudaErrorMemoryAllocation ? GR2M_ERR_NOMEM : GR2M_ERR_CUDA; } } while (0)
#define PRC(call) do { int r_ = (call); if (r_) { cleanup2(); return r_; } } while (0)
    PRC(keys.ensure(n * 4));
    PRC(keys2.ensure(n * 4));
    PRC(vals.ensure(n * 4));
    PRC(p->order.ensure(n * 4));
    PRC(p->omask.ensure(n));
    PRC(p->level_off.ensure(((size_t)p->nlevels + 2) * 4));
    gr2m_count_launch(), k_sort_keys<<<nb, 256, 0, st>>>(p->level.as<int32_t>(), n, p->nlevels, keys.as<uint32_t>(), vals.as<int32_t>(),
                                    p->contam.as<uint8_t>());
    PCK(cudaGetLastError());
    int bits = 1;
    while ((1ll << bits) <= p->nlevels) ++bits;
    size_t tbytes = 0;
    PCK(cub::DeviceRadixSort::SortPairs(nullptr, tbytes, keys.as<uint32_t>(), keys2.as<uint32_t>(), vals.as<int32_t>(),
                                        p->order.as<int32_t>(), n, 0, bits, st));
    PRC(temp.ensure(tbytes ? tbytes : 16));
    PCK(cub::DeviceRadixSort::SortPairs(temp.p, tbytes, keys.as<uint32_t>(), keys2.as<uint32_t>(), vals.as<int32_t>(),
                                        p->order.as<int32_t>(), n, 0, bits, st));
    PCK(cudaMemsetAsync(p->level_off.p, 0, ((size_t)p->nlevels + 2) * 4, st));
    gr2m_count_launch(), k_level_offsets<<<nb, 256, 0, st>>>(keys2.as<uint32_t>(), n, p->nlevels, p->level_off.as<int32_t>(),
                                        p->order.as<int32_t>(), p->inmask.as<uint8_t>(), p->omask.as<uint8_t>());
    PCK(cudaGetLastError());
    p->h_level_off.assign((size_t)p->nlevels + 1, 0);
    PCK(cudaMemcpyAsync(p->h_level_off.data(), p->level_off.p, ((size_t)p->nlevels + 1) * 4, cudaMemcpyDeviceToHost, st));
    PCK(cudaStreamSynchronize(st));
    cleanup2();
#undef PCK
#undef PRC
    p->norder = p->nlevels ? p->h_level_off[p->nlevels] : 0;
    // tail = suffix of levels that are all at most TAIL_WIDTH wide
    int tl = p->nlevels;
    while (tl > 0 && p->h_level_off[tl] - p->h_level_off[tl - 1] <= TAIL_WIDTH) --tl;
    if (tl < 1) tl = 1;                 // level 0 has nothing to add and is never swept
    if (tl > p->nlevels) tl = p->nlevels;
    p->tail_level = tl;
    if ((rc = p->tailmask.ensure(n))) return rc;
    if ((rc = p->done.ensure(n * 4))) return rc;
    if ((rc = p->ticket.ensure(sizeof(int)))) return rc;
    CK(cudaMemsetAsync(p->done.p, 0, n * 4, st));
    CK(cudaMemsetAsync(p->tailmask.p, 0, n, st));
    int tb = tl < p->nlevels ? p->h_level_off[tl] : p->norder, te = p->norder;
    if (te > tb) {
        gr2m_count_launch(), k_tailmask<<<ceil_div(te - tb, 256), 256, 0, st>>>(p->order.as<int32_t>(), tb, te, p->ncol, tl,
                                                           p->inmask.as<uint8_t>(), p->level.as<int32_t>(),
                                                           p->tailmask.as<uint8_t>());
        CK(cudaGetLastError());
    }
    CK(cudaStreamSynchronize(st));
    return plan_build_components(p);
}

// Regroup the upper forest (levels >= comp_level) by outlet basin.  Falls back to the level-wide
// sweep (comp_ok = false) when the forest does not split into enough comparable trees.
static int plan_build_components(wfa_plan *p)
{
    cudaStream_t st = p->stream;
    const std::vector<int32_t> &off = p->h_level_off;
    int comp_width = COMP_WIDTH;
    if (const char *ev = getenv("WFA_COMP_WIDTH")) comp_width = atoi(ev);   // development knob
    int lc = p->nlevels;
    while (lc > 1 && off[lc] - off[lc - 1] <= comp_width) --lc;
    p->comp_level = lc;
    p->comp_ok = false;
    if (lc >= p->nlevels) return GR2M_OK;
    const int ub = off[lc], nu = p->norder - ub;
    if (nu <= 0) return GR2M_OK;
    p->nupper = nu;
    const long long n = p->ncell;
    const int nb = ceil_div(nu, 256);
    int rc;
    DevBuf root, keys, keys2, vals, temp, lstart, cbeg, cnt;
    auto cleanup = [&]() { root.release(); keys.release(); keys2.release(); vals.release(); temp.release();
                           lstart.release(); cbeg.release(); cnt.release(); };
#define QCK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { gr2m_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); cleanup(); return e_ == cudaErrorMemoryAllocation ? GR2M_ERR_NOMEM : GR2M_ERR_CUDA; } } while (0)
#define QRC(call) do { int r_ = (call); if (r_) { cleanup(); return r_; } } while (0)
    QRC(root.ensure(n * 4));
    QRC(keys.ensure((size_t)nu * 4));
    QRC(keys2.ensure((size_t)nu * 4));
    QRC(vals.ensure((size_t)nu * 4));
    QRC(lstart.ensure(nu));
    QRC(cbeg.ensure((size_t)nu * 4));
    QRC(cnt.ensure(sizeof(int)));
    QRC(p->corder.ensure((size_t)nu * 4));
    QRC(p->cmask.ensure(nu));
    QRC(p->cnext.ensure((size_t)nu * 4));
    const int32_t *order = p->order.as<int32_t>();
    gr2m_count_launch(), k_root_init<<<nb, 256, 0, st>>>(order, ub, nu, p->receiver.as<int32_t>(), root.as<int32_t>());
    int depth = p->nlevels - lc + 1, jumps = 1;
    while ((1 << jumps) < depth) ++jumps;
    for (int it = 0; it < jumps + 1; ++it) gr2m_count_launch(), k_root_jump<<<nb, 256, 0, st>>>(order, ub, nu, root.as<int32_t>());
    gr2m_count_launch(), k_root_keys<<<nb, 256, 0, st>>>(order, ub, nu, root.as<int32_t>(), keys.as<uint32_t>(), vals.as<int32_t>());
    QCK(cudaGetLastError());
    int bits = 1;
    while ((1ll << bits) < n) ++bits;
    size_t tbytes = 0;
    QCK(cub::DeviceRadixSort::SortPairs(nullptr, tbytes, keys.as<uint32_t>(), keys2.as<uint32_t>(), vals.as<int32_t>(),
                                        p->corder.as<int32_t>(), nu, 0, bits, st));
    QRC(temp.ensure(tbytes ? tbytes : 16));
    QCK(cub::DeviceRadixSort::SortPairs(temp.p, tbytes, keys.as<uint32_t>(), keys2.as<uint32_t>(), vals.as<int32_t>(),
                                        p->corder.as<int32_t>(), nu, 0, bits, st));
    QCK(cudaMemsetAsync(cnt.p, 0, sizeof(int), st));
    gr2m_count_launch(), k_comp_flags<<<nb, 256, 0, st>>>(p->corder.as<int32_t>(), keys2.as<uint32_t>(), nu, p->level.as<int32_t>(),
                                     p->inmask.as<uint8_t>(), p->cmask.as<uint8_t>(), lstart.as<uint8_t>(),
                                     cbeg.as<int32_t>(), cnt.as<int>());
    gr2m_count_launch(), k_comp_next<<<nb, 256, 0, st>>>(lstart.as<uint8_t>(), nu, p->cnext.as<int32_t>());
    QCK(cudaGetLastError());
    int ncomp = 0;
    QCK(cudaMemcpyAsync(&ncomp, cnt.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    // CSR of upstream cells: counts -> exclusive scan -> fill
    QRC(p->cup_off.ensure(((size_t)nu + 1) * 4));
    gr2m_count_launch(), k_comp_upcount<<<nb, 256, 0, st>>>(p->cmask.as<uint8_t>(), nu, keys.as<int32_t>());
    QCK(cudaMemsetAsync(p->cup_off.p, 0, 4, st));
    {
        size_t sbytes = 0;
        QCK(cub::DeviceScan::InclusiveSum(nullptr, sbytes, keys.as<int32_t>(), p->cup_off.as<int32_t>() + 1, nu, st));
        QRC(temp.ensure(sbytes ? sbytes : 16));
        QCK(cub::DeviceScan::InclusiveSum(temp.p, sbytes, keys.as<int32_t>(), p->cup_off.as<int32_t>() + 1, nu, st));
    }
    int nedges = 0;
    QCK(cudaMemcpyAsync(&nedges, p->cup_off.as<int32_t>() + nu, sizeof(int), cudaMemcpyDeviceToHost, st));
    QCK(cudaStreamSynchronize(st));
    QRC(p->cup_idx.ensure(std::max<size_t>(16, (size_t)nedges * 4)));
    gr2m_count_launch(), k_comp_upfill<<<nb, 256, 0, st>>>(p->corder.as<int32_t>(), p->cmask.as<uint8_t>(), nu, p->ncol,
                                      p->level.as<int32_t>(), p->cup_off.as<int32_t>(), p->cup_idx.as<uint32_t>());
    QCK(cudaGetLastError());
    QCK(cudaStreamSynchronize(st));
    std::vector<int32_t> beg(ncomp);
    if (ncomp) QCK(cudaMemcpy(beg.data(), cbeg.p, (size_t)ncomp * 4, cudaMemcpyDeviceToHost));
    std::sort(beg.begin(), beg.end());
    std::vector<int2> comps(ncomp);
    int largest = 0;
    for (int k = 0; k < ncomp; ++k) {
        comps[k].x = beg[k];
        comps[k].y = k + 1 < ncomp ? beg[k + 1] : nu;
        largest = std::max(largest, comps[k].y - comps[k].x);
    }
    std::stable_sort(comps.begin(), comps.end(), [](const int2 &a, const int2 &b) { return a.y - a.x > b.y - b.x; });
    p->ncomp = ncomp;
    if (ncomp > 0) {
        rc = p->comps.ensure((size_t)ncomp * sizeof(int2));
        if (rc) { cleanup(); return rc; }
        QCK(cudaMemcpy(p->comps.p, comps.data(), (size_t)ncomp * sizeof(int2), cudaMemcpyHostToDevice));
    }
    // worth it only if the forest splits into many trees none of which dominates
    p->comp_ok = ncomp >= 32 && (long long)largest * 8 <= nu;
    cleanup();
#undef QCK
#undef QRC
    int rc_runs = plan_build_runs(p);
    if (rc_runs) return rc_runs;
    return plan_build_segments(p);
}


// Regroup the upper network (levels >= comp_level) by river reach; see k_reach_sweep.
static int plan_build_runs(wfa_plan *p)
{
    p->run_ok = false;
    const std::vector<int32_t> &off = p->h_level_off;
    const int lc = p->comp_level;
    if (lc >= p->nlevels) return GR2M_OK;
    const int ub = off[lc], nu = p->norder - ub;
    if (nu <= 0 || nu >= (1 << 27)) return GR2M_OK;     // sort key bits; stream rows (<= 8 per cell) stay below 2^30
    cudaStream_t st = p->stream;
    const long long n = p->ncell;
    const int nb = ceil_div(nu, 256);
    const int32_t *uorder = p->order.as<int32_t>() + ub;
    DevBuf upos, pred, npu, h0, h1, d0, d1, rl, flag, keys, keys2, vals, sj, temp, isstart, rank, cnt, runrl;
    auto cleanup = [&]() {
        DevBuf *t[] = {&upos, &pred, &npu, &h0, &h1, &d0, &d1, &rl, &flag, &keys, &keys2, &vals, &sj, &temp, &isstart,
                       &rank, &cnt, &runrl};
        FreeBatch fb;
        for (DevBuf *d : t) d->release();
    };
#define RCK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { gr2m_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); cleanup(); return e_ == cudaErrorMemoryAllocation ? GR2M_ERR_NOMEM : GR2M_ERR_CUDA; } } while (0)
#define RRC(call) do { int r_ = (call); if (r_) { cleanup(); return r_; } } while (0)
    RRC(upos.ensure(n * 4));
    RRC(pred.ensure((size_t)nu * 4));
    RRC(npu.ensure(nu));
    RRC(h0.ensure((size_t)nu * 4));
    RRC(h1.ensure((size_t)nu * 4));
    RRC(d0.ensure((size_t)nu * 4));
    RRC(d1.ensure((size_t)nu * 4));
    RRC(rl.ensure((size_t)nu * 4));
    RRC(flag.ensure(sizeof(int)));
    gr2m_count_launch(), k_run_upos<<<nb, 256, 0, st>>>(uorder, nu, upos.as<int32_t>());
    gr2m_count_launch(), k_run_pred<<<nb, 256, 0, st>>>(uorder, nu, p->ncol, lc, p->inmask.as<uint8_t>(), p->level.as<int32_t>(),
                                   upos.as<int32_t>(), pred.as<int32_t>(), npu.as<uint8_t>());
    gr2m_count_launch(), k_run_rank_init<<<nb, 256, 0, st>>>(pred.as<int32_t>(), nu, h0.as<int32_t>(), d0.as<int32_t>());
    RCK(cudaGetLastError());
    int32_t *hs = h0.as<int32_t>(), *hd = h1.as<int32_t>(), *ds = d0.as<int32_t>(), *dd = d1.as<int32_t>();
    int changed = 1;
    for (int it = 0; changed && it < 40; ++it) {
        RCK(cudaMemsetAsync(flag.p, 0, sizeof(int), st));
        gr2m_count_launch(), k_run_rank_step<<<nb, 256, 0, st>>>(hs, ds, nu, hd, dd, flag.as<int>());
        RCK(cudaGetLastError());
        RCK(cudaMemcpyAsync(&changed, flag.p, sizeof(int), cudaMemcpyDeviceToHost, st));
        RCK(cudaStreamSynchronize(st));
        std::swap(hs, hd);
        std::swap(ds, dd);
    }
    // reach levels
    RCK(cudaMemsetAsync(rl.p, 0, (size_t)nu * 4, st));
    changed = 1;
    int nrl_iter = 0;
    for (; changed && nrl_iter < 100000; ++nrl_iter) {
        RCK(cudaMemsetAsync(flag.p, 0, sizeof(int), st));
        gr2m_count_launch(), k_run_level_step<<<nb, 256, 0, st>>>(uorder, nu, p->ncol, lc, p->inmask.as<uint8_t>(), p->level.as<int32_t>(),
                                             upos.as<int32_t>(), hs, npu.as<uint8_t>(), rl.as<int32_t>(), flag.as<int>());
        RCK(cudaGetLastError());
        RCK(cudaMemcpyAsync(&changed, flag.p, sizeof(int), cudaMemcpyDeviceToHost, st));
        RCK(cudaStreamSynchronize(st));
    }
    if (nrl_iter >= 65000) { cleanup(); return GR2M_OK; }     // does not fit the sort key: keep the other sweeps
    // sort by (reach level, reach head, position in reach)
    RRC(keys.ensure((size_t)nu * 8));
    RRC(keys2.ensure((size_t)nu * 8));
    RRC(vals.ensure((size_t)nu * 4));
    RRC(sj.ensure((size_t)nu * 4));
    RCK(cudaMemsetAsync(flag.p, 0, sizeof(int), st));
    gr2m_count_launch(), k_run_keys<<<nb, 256, 0, st>>>(hs, ds, rl.as<int32_t>(), nu, keys.as<unsigned long long>(), vals.as<int32_t>(),
                                   flag.as<int>());
    RCK(cudaGetLastError());
    RCK(cudaMemcpyAsync(&changed, flag.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    RCK(cudaStreamSynchronize(st));
    if (changed) { cleanup(); return GR2M_OK; }               // keep the per-basin sweep
    size_t tbytes = 0;
    RCK(cub::DeviceRadixSort::SortPairs(nullptr, tbytes, keys.as<unsigned long long>(), keys2.as<unsigned long long>(),
                                        vals.as<int32_t>(), sj.as<int32_t>(), nu, 0, 64, st));
    RRC(temp.ensure(tbytes ? tbytes : 16));
    RCK(cub::DeviceRadixSort::SortPairs(temp.p, tbytes, keys.as<unsigned long long>(), keys2.as<unsigned long long>(),
                                        vals.as<int32_t>(), sj.as<int32_t>(), nu, 0, 64, st));
    RRC(p->rcell.ensure((size_t)nu * 4));
    RRC(p->rup_off.ensure(((size_t)nu + 1) * 4));
    RRC(isstart.ensure((size_t)nu * 4));
    RRC(rank.ensure((size_t)nu * 4));
    RRC(cnt.ensure((size_t)nu * 4));
    gr2m_count_launch(), k_run_gather<<<nb, 256, 0, st>>>(sj.as<int32_t>(), nu, uorder, ds, p->inmask.as<uint8_t>(), p->rcell.as<int32_t>(),
                                     isstart.as<int32_t>(), cnt.as<int32_t>());
    RCK(cudaGetLastError());
    size_t sbytes = 0;
    RCK(cub::DeviceScan::InclusiveSum(nullptr, sbytes, cnt.as<int32_t>(), p->rup_off.as<int32_t>() + 1, nu, st));
    if (sbytes > temp.cap) RRC(temp.ensure(sbytes));
    RCK(cudaMemsetAsync(p->rup_off.p, 0, 4, st));
    RCK(cub::DeviceScan::InclusiveSum(temp.p, sbytes, cnt.as<int32_t>(), p->rup_off.as<int32_t>() + 1, nu, st));
    RCK(cub::DeviceScan::InclusiveSum(nullptr, sbytes, isstart.as<int32_t>(), rank.as<int32_t>(), nu, st));
    if (sbytes > temp.cap) RRC(temp.ensure(sbytes));
    RCK(cub::DeviceScan::InclusiveSum(temp.p, sbytes, isstart.as<int32_t>(), rank.as<int32_t>(), nu, st));
    int nedges = 0, nr = 0;
    RCK(cudaMemcpyAsync(&nedges, p->rup_off.as<int32_t>() + nu, 4, cudaMemcpyDeviceToHost, st));
    RCK(cudaMemcpyAsync(&nr, rank.as<int32_t>() + (nu - 1), 4, cudaMemcpyDeviceToHost, st));
    RCK(cudaStreamSynchronize(st));
    if (nr <= 0) { cleanup(); return GR2M_OK; }
    RRC(p->rup_idx.ensure(std::max<size_t>(16, (size_t)nedges * 4)));
    RRC(p->run_start.ensure(((size_t)nr + 1) * 4));
    RRC(runrl.ensure((size_t)nr * 4));
    gr2m_count_launch(), k_run_starts<<<nb, 256, 0, st>>>(isstart.as<int32_t>(), rank.as<int32_t>(), nu, p->run_start.as<int32_t>());
    RCK(cudaMemcpyAsync(p->run_start.as<int32_t>() + nr, &nu, 4, cudaMemcpyHostToDevice, st));
    gr2m_count_launch(), k_run_upfill<<<nb, 256, 0, st>>>(sj.as<int32_t>(), nu, p->ncol, p->rcell.as<int32_t>(), p->inmask.as<uint8_t>(),
                                     pred.as<int32_t>(), uorder, p->rup_off.as<int32_t>(), p->rup_idx.as<uint32_t>());
    // the reach stream: rows per cell, their offsets, and which cell completes at which row
    RRC(p->row_off.ensure(((size_t)nu + 1) * 4));
    gr2m_count_launch(), k_run_rows<<<nb, 256, 0, st>>>(isstart.as<int32_t>(), p->rup_off.as<int32_t>(), p->rup_idx.as<uint32_t>(), nu,
                                   cnt.as<int32_t>());
    RCK(cudaGetLastError());
    RCK(cub::DeviceScan::InclusiveSum(nullptr, sbytes, cnt.as<int32_t>(), p->row_off.as<int32_t>() + 1, nu, st));
    if (sbytes > temp.cap) RRC(temp.ensure(sbytes));
    RCK(cudaMemsetAsync(p->row_off.p, 0, 4, st));
    RCK(cub::DeviceScan::InclusiveSum(temp.p, sbytes, cnt.as<int32_t>(), p->row_off.as<int32_t>() + 1, nu, st));
    int nrows = 0;
    RCK(cudaMemcpyAsync(&nrows, p->row_off.as<int32_t>() + nu, 4, cudaMemcpyDeviceToHost, st));
    RCK(cudaStreamSynchronize(st));
    RRC(p->rowcell.ensure(((size_t)nrows + 2 * RS_ROWS) * 4));        // the sweep copies whole RS_ROWS blocks
    RCK(cudaMemsetAsync(p->rowcell.p, 0xff, ((size_t)nrows + 2 * RS_ROWS) * 4, st));
    gr2m_count_launch(), k_run_rowcell<<<nb, 256, 0, st>>>(p->row_off.as<int32_t>(), p->rcell.as<int32_t>(), nu, p->rowcell.as<int32_t>());
    RRC(p->rinfo.ensure((size_t)nr * RINFO_VEC * 16));
    RRC(p->rdone.ensure((size_t)nr * 4));
    RCK(cudaMemsetAsync(p->rdone.p, 0, (size_t)nr * 4, st));
    int32_t *inv = vals.as<int32_t>();            // the sort's input values are no longer needed
    gr2m_count_launch(), k_run_inv<<<nb, 256, 0, st>>>(sj.as<int32_t>(), nu, inv);
    gr2m_count_launch(), k_run_pack<<<ceil_div(nr, 256), 256, 0, st>>>(p->run_start.as<int32_t>(), nr, p->rcell.as<int32_t>(),
                                                 p->rup_off.as<int32_t>(), p->rup_idx.as<uint32_t>(),
                                                 p->row_off.as<int32_t>(), p->level.as<int32_t>(), lc,
                                                 upos.as<int32_t>(), inv, rank.as<int32_t>(), p->rinfo.as<int4>());
    p->nrows = nrows;
    gr2m_count_launch(), k_run_levels<<<ceil_div(nr, 256), 256, 0, st>>>(p->run_start.as<int32_t>(), nr, sj.as<int32_t>(), rl.as<int32_t>(),
                                                   runrl.as<int32_t>());
    RCK(cudaGetLastError());
    std::vector<int32_t> hrl(nr);
    RCK(cudaMemcpyAsync(hrl.data(), runrl.p, (size_t)nr * 4, cudaMemcpyDeviceToHost, st));
    RCK(cudaStreamSynchronize(st));
    p->h_rlv_off.clear();
    for (int r = 0; r < nr; ++r)
        if (r == 0 || hrl[r] != hrl[r - 1]) p->h_rlv_off.push_back(r);      // reaches are sorted by level
    p->h_rlv_off.push_back(nr);
    {   // latency chain of the level-synchronous reach sweep: longest reach of every level, summed
        std::vector<int32_t> hs(nr + 1);
        RCK(cudaMemcpy(hs.data(), p->run_start.p, ((size_t)nr + 1) * 4, cudaMemcpyDeviceToHost));
        long long crit = 0;
        for (size_t l = 0; l + 1 < p->h_rlv_off.size(); ++l) {
            int mx = 0;
            for (int r = p->h_rlv_off[l]; r < p->h_rlv_off[l + 1]; ++r) mx = std::max(mx, hs[r + 1] - hs[r]);
            crit += mx;
        }
        p->run_critical = (int)std::min<long long>(crit, INT32_MAX);
    }
    p->nruns = nr;
    p->run_ok = true;
    cleanup();
#undef RCK
#undef RRC
    return GR2M_OK;
}


__global__ void k_seg_offsets(const int32_t *__restrict__ level_off, int begin, int nlev, const int32_t *__restrict__ slot,
                              int32_t *__restrict__ seg_off)
{
    int l = blockIdx.x * blockDim.x + threadIdx.x;      // seg_off[l] = heads before level l, l = 1 .. nlev
    if (l < 1 || l > nlev) return;
    const int pos = level_off[l] - begin;
    seg_off[l] = pos > 0 ? slot[pos - 1] : 0;
}

// Chain segments of the levels [1, comp_level) (see k_accum_segments).  Optional: on any failure the level sweep stays.
static int plan_build_segments(wfa_plan *p)
{
    p->seg_ok = false;
    p->nseg = p->nseg_cells = 0;
    static const bool off_env = getenv("WFA_NO_SEGMENTS") != nullptr;
    const int L = p->comp_level;
    if (off_env || !(p->run_ok || p->comp_ok) || L <= 2) return GR2M_OK;
    cudaStream_t st = p->stream;
    const std::vector<int32_t> &off = p->h_level_off;
    const int begin = off[1], end = off[L], n = end - begin;
    if (n <= 0) return GR2M_OK;
    DevBuf slot, temp, segoff, cnt;
    auto cleanup = [&]() { FreeBatch fb; for (DevBuf *b : {&slot, &temp, &segoff, &cnt}) b->release(); };
#define SCK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { gr2m_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); cleanup(); return e_ == cudaErrorMemoryAllocation ? GR2M_ERR_NOMEM : GR2M_ERR_CUDA; } } while (0)
#define SRC(call) do { int r_ = (call); if (r_) { cleanup(); return r_; } } while (0)
    SRC(slot.ensure((size_t)n * 4)); SRC(segoff.ensure(((size_t)L + 2) * 4)); SRC(cnt.ensure(8));
    gr2m_count_launch(), k_seg_flag<<<ceil_div(n, 256), 256, 0, st>>>(p->order.as<int32_t>(), begin, end, L, p->ncol, p->inmask.as<uint8_t>(),
                                                                     p->level.as<int32_t>(), slot.as<int32_t>());
    SCK(cudaGetLastError());
    size_t tbytes = 0;        // head flags -> their inclusive scan, in place
    SCK(cub::DeviceScan::InclusiveSum(nullptr, tbytes, slot.as<int32_t>(), slot.as<int32_t>(), n, st));
    SRC(temp.ensure(tbytes));
    SCK(cub::DeviceScan::InclusiveSum(temp.p, tbytes, slot.as<int32_t>(), slot.as<int32_t>(), n, st));
    int32_t nseg = 0;
    SCK(cudaMemcpyAsync(&nseg, slot.as<int32_t>() + (n - 1), 4, cudaMemcpyDeviceToHost, st));
    SCK(cudaStreamSynchronize(st));
    if (nseg <= 0) { cleanup(); return GR2M_OK; }
    SRC(p->segs.ensure((size_t)nseg * sizeof(SegDesc)));
    SCK(cudaMemsetAsync(cnt.p, 0, 8, st));
    gr2m_count_launch(), k_seg_build<<<ceil_div(n, 256), 256, 0, st>>>(p->order.as<int32_t>(), begin, end, L, p->ncol, p->inmask.as<uint8_t>(),
                                                                      p->level.as<int32_t>(), p->receiver.as<int32_t>(), slot.as<int32_t>(), p->segs.as<SegDesc>(),
                                                                      cnt.as<unsigned long long>());
    SCK(cudaGetLastError());
    gr2m_count_launch(), k_seg_offsets<<<ceil_div(L + 1, 256), 256, 0, st>>>(p->level_off.as<int32_t>(), begin, L, slot.as<int32_t>(),
                                                                        segoff.as<int32_t>());
    SCK(cudaGetLastError());
    p->h_seg_off.assign((size_t)L + 1, 0);
    unsigned long long ncells = 0;
    SCK(cudaMemcpyAsync(p->h_seg_off.data() + 1, segoff.as<int32_t>() + 1, (size_t)L * 4, cudaMemcpyDeviceToHost, st));
    SCK(cudaMemcpyAsync(&ncells, cnt.p, 8, cudaMemcpyDeviceToHost, st));
    SCK(cudaStreamSynchronize(st));
    cleanup();
#undef SCK
#undef SRC
    if ((long long)ncells != (long long)n || p->h_seg_off[L] != nseg) {      // every cell of the levels in exactly one segment
        p->segs.release();
        return GR2M_OK;
    }
    p->seg_levels = L;
    p->nseg = nseg;
    p->nseg_cells = (long long)ncells;
    p->seg_ok = true;
    return GR2M_OK;
}

static int plan_create_common(const int16_t *flowdir, bool on_device, int nrow, int ncol, int flags, wfa_plan **out)
{
    REQUIRE(out != nullptr, "out is NULL");
    *out = nullptr;
    REQUIRE(flowdir != nullptr, "flowdir is NULL");
    REQUIRE(nrow > 0 && ncol > 0, "nrow and ncol must be positive");
    REQUIRE((long long)nrow * ncol < 0x7fffffffll, "grid too large for 32-bit cell ids");
    int rc = gr2m_check_device();
    if (rc) return rc;
    CK(cudaSetDevice(gr2m_current_device()));
    wfa_plan *p = new (std::nothrow) wfa_plan();
    if (!p) { gr2m_set_error("host allocation failed"); return GR2M_ERR_NOMEM; }
    p->device = gr2m_current_device();
    p->nrow = nrow; p->ncol = ncol; p->ncell = (long long)nrow * ncol; p->flags = flags;
    cudaError_t e = cudaStreamCreateWithFlags(&p->own_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreate(&p->ev0);
    if (e == cudaSuccess) e = cudaEventCreate(&p->ev1);
    if (e == cudaSuccess) e = cudaEventCreate(&p->evm);
    if (e != cudaSuccess) {
        gr2m_set_error("wfa_plan_create: %s", cudaGetErrorString(e));
        plan_free(p);
        return GR2M_ERR_CUDA;
    }
    p->stream = p->own_stream;
    DevBuf dir;
    const int16_t *dir_dev = flowdir;
    if (!on_device) {
        rc = dir.ensure(p->ncell * sizeof(int16_t));
        if (!rc) {
            e = cudaMemcpyAsync(dir.p, flowdir, p->ncell * sizeof(int16_t), cudaMemcpyHostToDevice, p->stream);
            if (e != cudaSuccess) { gr2m_set_error("wfa_plan_create: %s", cudaGetErrorString(e)); rc = GR2M_ERR_CUDA; }
        }
        dir_dev = dir.as<int16_t>();
    }
    if (!rc) rc = plan_build(p, dir_dev);
    dir.release();
    if (rc) { plan_free(p); return rc; }
    *out = p;
    return GR2M_OK;
}

extern "C" int wfa_plan_create(const int16_t *flowdir, int nrow, int ncol, int flags, wfa_plan **out)
{
    return plan_create_common(flowdir, false, nrow, ncol, flags, out);
}

extern "C" int wfa_plan_create_dev(const int16_t *flowdir_dev, int nrow, int ncol, int flags, wfa_plan **out)
{
    return plan_create_common(flowdir_dev, true, nrow, ncol, flags, out);
}

extern "C" int wfa_plan_destroy(wfa_plan *p)
{
    plan_free(p);
    return GR2M_OK;
}

extern "C" int wfa_plan_set_stream(wfa_plan *p, void *cuda_stream)
{
    REQUIRE(p != nullptr, "plan is NULL");
    p->stream = cuda_stream ? (cudaStream_t)cuda_stream : p->own_stream;
    return GR2M_OK;
}

extern "C" int wfa_plan_info(const wfa_plan *p, int *nlevels, int *norder, int64_t *ncell)
{
    REQUIRE(p != nullptr, "plan is NULL");
    if (nlevels) *nlevels = p->nlevels;
    if (norder) *norder = p->norder;
    if (ncell) *ncell = p->ncell;
    return GR2M_OK;
}

extern "C" int wfa_plan_basin_info(const wfa_plan *p, int *cut_level, int *nbasins, int *nupper, int *enabled)
{
    REQUIRE(p != nullptr, "plan is NULL");
    if (cut_level) *cut_level = p->comp_level;
    if (nbasins) *nbasins = p->ncomp;
    if (nupper) *nupper = p->nupper;
    if (enabled) *enabled = p->comp_ok ? 1 : 0;
    return GR2M_OK;
}

extern "C" int wfa_plan_reach_info(const wfa_plan *p, int *nreaches, int *nreach_levels, int *critical_cells,
                                   int *enabled)
{
    REQUIRE(p != nullptr, "plan is NULL");
    if (nreaches) *nreaches = p->run_ok ? p->nruns : 0;
    if (nreach_levels) *nreach_levels = p->run_ok ? (int)p->h_rlv_off.size() - 1 : 0;
    if (critical_cells) *critical_cells = p->run_ok ? p->run_critical : 0;
    if (enabled) *enabled = p->run_ok ? 1 : 0;
    return GR2M_OK;
}

extern "C" int wfa_plan_segment_info(const wfa_plan *p, long long *nsegments, long long *ncells, int *levels, int *enabled)
{
    REQUIRE(p != nullptr, "plan is NULL");
    if (nsegments) *nsegments = p->seg_ok ? p->nseg : 0;
    if (ncells) *ncells = p->seg_ok ? p->nseg_cells : 0;
    if (levels) *levels = p->seg_ok ? p->seg_levels : 0;
    if (enabled) *enabled = p->seg_ok ? 1 : 0;
    return GR2M_OK;
}

extern "C" int wfa_plan_export(const wfa_plan *pc, int32_t *receiver, uint8_t *inmask, int32_t *indeg, int32_t *level,
                               int32_t *order, int32_t *level_off, uint8_t *contam)
{
    REQUIRE(pc != nullptr, "plan is NULL");
    wfa_plan *p = const_cast<wfa_plan *>(pc);
    CK(cudaSetDevice(p->device));
    const long long n = p->ncell;
    if (receiver) CK(cudaMemcpy(receiver, p->receiver.p, n * 4, cudaMemcpyDeviceToHost));
    if (inmask) CK(cudaMemcpy(inmask, p->inmask.p, n, cudaMemcpyDeviceToHost));
    if (indeg) CK(cudaMemcpy(indeg, p->indeg.p, n * 4, cudaMemcpyDeviceToHost));
    if (level) CK(cudaMemcpy(level, p->level.p, n * 4, cudaMemcpyDeviceToHost));
    if (order && p->norder) CK(cudaMemcpy(order, p->order.p, (size_t)p->norder * 4, cudaMemcpyDeviceToHost));
    if (level_off) memcpy(level_off, p->h_level_off.data(), ((size_t)p->nlevels + 1) * 4);
    if (contam) CK(cudaMemcpy(contam, p->contam.p, n, cudaMemcpyDeviceToHost));
    return GR2M_OK;
}

extern "C" int wfa_plan_last_phase_ms(wfa_plan *p, float *wide_ms, float *tail_ms)
{
    REQUIRE(p && wide_ms && tail_ms, "NULL argument");
    REQUIRE(p->timed, "no timed call yet");
    CK(cudaSetDevice(p->device));
    CK(cudaEventSynchronize(p->ev1));
    CK(cudaEventElapsedTime(wide_ms, p->ev0, p->evm));
    CK(cudaEventElapsedTime(tail_ms, p->evm, p->ev1));
    return GR2M_OK;
}

extern "C" int wfa_plan_last_kernel_ms(wfa_plan *p, float *ms)
{
    REQUIRE(p && ms, "NULL argument");
    REQUIRE(p->timed, "no timed call yet");
    CK(cudaSetDevice(p->device));
    CK(cudaEventSynchronize(p->ev1));
    CK(cudaEventElapsedTime(ms, p->ev0, p->ev1));
    return GR2M_OK;
}

template <typename T, int G>
static int sweep_g(wfa_plan *p, int nmonth, int ld, T *A, int flags)
{
    cudaStream_t st = p->stream;
    const int32_t *order = p->order.as<int32_t>();
    const uint8_t *omask = p->omask.as<uint8_t>();   // in-flow masks in sweep order
    const std::vector<int32_t> &off = p->h_level_off;
    const int gpb = 256 / G;   // lane groups per 256-thread block
    const bool dataflow = (flags & WFA_DATAFLOW) != 0;
    const bool by_reach = p->run_ok && !dataflow && !(flags & (WFA_LEVEL_SYNC | WFA_BY_BASIN));
    const bool by_basin = p->comp_ok && !dataflow && !(flags & WFA_LEVEL_SYNC) && !by_reach;
    // the thin tail starts where every later level fits a few passes of the single resident CTA
    int l0 = p->tail_level;
    if (by_basin || by_reach) {
        l0 = p->comp_level;
    } else if (!dataflow) {
        const int width = 4 * (1024 / G);
        l0 = p->nlevels;
        while (l0 > 1 && off[l0] - off[l0 - 1] <= width) --l0;
    }
    // reach sweep: the stream buffer
    const int nu = by_reach ? p->norder - off[p->comp_level] : 0;
    const int lds = (nmonth + 3) & ~3;            // stream rows are multiples of 16 bytes for the bulk copies
    T *S = nullptr;
    if (by_reach && l0 < p->nlevels) {
        int rc_ = p->rstream.ensure(((size_t)p->nrows + 2 * RS_ROWS) * lds * sizeof(T));
        if (rc_) return rc_;
        S = p->rstream.as<T>();
    }
    // level 0 has no upstream cells: nothing to add
    {
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = (flags & WFA_NO_OVERLAP) ? 0 : 1;
        cudaLaunchConfig_t cfg = {};
        cfg.blockDim = dim3(256);
        cfg.stream = st;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        const bool by_seg = p->seg_ok && (by_reach || by_basin) && l0 == p->seg_levels && !(flags & WFA_NO_SEGMENTS);
        for (int l = 1; l < l0; ++l) {
            if (by_seg) {      // chain segments headed at this level (absorbed cells of later levels ride along)
                const int b = p->h_seg_off[l], e = l + 1 < l0 ? p->h_seg_off[l + 1] : (int)p->nseg;
                if (e <= b) continue;
                cfg.gridDim = dim3(ceil_div(e - b, gpb));
                gr2m_count_launch();
                CK(cudaLaunchKernelEx(&cfg, k_accum_segments<T, G>, (const SegDesc *)p->segs.as<SegDesc>(), b, e, p->ncol, A, ld, nmonth));
                continue;
            }
            int b = off[l], e = off[l + 1];
            cfg.gridDim = dim3(ceil_div(ceil_div(e - b, 2), gpb));
            gr2m_count_launch();
            CK(cudaLaunchKernelEx(&cfg, k_accum_level<T, G>, order, b, e, p->ncol, omask, A, ld, nmonth));
        }
    }
    CK(cudaGetLastError());
    CK(cudaEventRecord(p->evm, st));
    if (l0 < p->nlevels) {
        if (by_reach) {
            gr2m_count_launch(), k_reach_gather<T, G><<<ceil_div(nu, gpb), 256, 0, st>>>(p->rcell.as<int32_t>(), p->rup_off.as<int32_t>(),
                                                                    p->rup_idx.as<uint32_t>(), p->row_off.as<int32_t>(), nu,
                                                                    A, S, ld, lds, nmonth);
            if (flags & WFA_NO_OVERLAP) {         // one launch per reach level
                for (size_t l = 0; l + 1 < p->h_rlv_off.size(); ++l) {
                    const int r0 = p->h_rlv_off[l], r1 = p->h_rlv_off[l + 1];
                    gr2m_count_launch(), k_reach_sweep<T><<<ceil_div(r1 - r0, RS_WARPS), 32 * RS_WARPS, 0, st>>>(
                        p->rinfo.as<int4>(), p->rowcell.as<int32_t>(), r0, r1, S, lds, A, ld, nmonth, nullptr, nullptr, 0u);
                }
            } else {                              // all levels in one launch, heads wait for the reaches that feed them
                p->epoch += 1;
                CK(cudaMemsetAsync(p->ticket.p, 0, sizeof(int), st));
                gr2m_count_launch(), k_reach_sweep<T><<<ceil_div(p->nruns, RS_WARPS), 32 * RS_WARPS, 0, st>>>(
                    p->rinfo.as<int4>(), p->rowcell.as<int32_t>(), 0, p->nruns, S, lds, A, ld, nmonth, p->ticket.as<int>(),
                    p->rdone.as<unsigned>(), p->epoch);
            }
        } else if (by_basin) {
            CK(cudaMemsetAsync(p->ticket.p, 0, sizeof(int), st));
            int blocks = std::min(p->ncomp, 148 * 7);
            if (G >= 16 && nmonth <= G && !(flags & WFA_NO_FORWARD))
                gr2m_count_launch(), k_accum_components_fw<T, G><<<blocks, 288, 0, st>>>(p->corder.as<int32_t>(), p->cup_off.as<int32_t>(),
                                                                    p->cup_idx.as<uint32_t>(), p->cnext.as<int32_t>(),
                                                                    p->comps.as<int2>(), p->ncomp, p->ticket.as<int>(),
                                                                    A, ld, nmonth);
            else
                gr2m_count_launch(), k_accum_components<T, G><<<blocks, 288, 0, st>>>(p->corder.as<int32_t>(), p->cup_off.as<int32_t>(),
                                                                 p->cup_idx.as<uint32_t>(), p->cnext.as<int32_t>(),
                                                                 p->comps.as<int2>(), p->ncomp, p->ticket.as<int>(), A,
                                                                 ld, nmonth);
        } else if (!dataflow) {
            gr2m_count_launch(), k_accum_thin<T, G><<<1, 1024, 0, st>>>(order, p->level_off.as<int32_t>(), l0, p->nlevels, p->ncol, omask, A,
                                                   ld, nmonth);
        } else {
            int b = off[l0], e = p->norder;
            p->epoch += 1;
            CK(cudaMemsetAsync(p->ticket.p, 0, sizeof(int), st));
            int groups = ceil_div(e - b, 4);
            int blocks = std::min(ceil_div(groups, gpb), 148 * 4);
            gr2m_count_launch(), k_accum_dataflow<T, G><<<blocks, 256, 0, st>>>(order, b, e, p->ncol, omask, p->tailmask.as<uint8_t>(), A, ld,
                                                           nmonth, p->done.as<unsigned>(), p->epoch, p->ticket.as<int>());
        }
        CK(cudaGetLastError());
    }
    return GR2M_OK;
}

template <typename T>
static int sweep(wfa_plan *p, int nmonth, int ld, T *A, int flags)
{
    static const int gmin = getenv("WFA_G32_FROM") ? atoi(getenv("WFA_G32_FROM")) : 32;   // experiment: months from which 32 lanes own a cell
    if (nmonth >= gmin) return sweep_g<T, 32>(p, nmonth, ld, A, flags);
    if (nmonth >= 16) return sweep_g<T, 16>(p, nmonth, ld, A, flags);
    if (nmonth >= 8) return sweep_g<T, 8>(p, nmonth, ld, A, flags);
    return sweep_g<T, 4>(p, nmonth, ld, A, flags);
}

extern "C" int wfa_accumulate_dev(wfa_plan *p, int nmonth, int ld, void *WA_dev, int flags)
{
    REQUIRE(p && WA_dev, "NULL plan or buffer");
    REQUIRE(nmonth > 0 && ld >= nmonth, "need 0 < nmonth <= ld");
    CK(cudaSetDevice(p->device));
    CK(cudaEventRecord(p->ev0, p->stream));
    int rc = (flags & WFA_F32) ? sweep<float>(p, nmonth, ld, (float *)WA_dev, flags)
                               : sweep<double>(p, nmonth, ld, (double *)WA_dev, flags);
    if (rc) return rc;
    CK(cudaEventRecord(p->ev1, p->stream));
    p->timed = true;
    return GR2M_OK;
}

extern "C" int wfa_mask_dev(wfa_plan *p, int nmonth, int ld, void *WA_dev, int flags)
{
    REQUIRE(p && WA_dev, "NULL plan or buffer");
    REQUIRE(nmonth > 0 && ld >= nmonth, "need 0 < nmonth <= ld");
    CK(cudaSetDevice(p->device));
    const int use_contam = (flags & WFA_NO_CONTAMINATION) ? 0 : 1;
    long long total = p->ncell * nmonth;
    if (flags & WFA_F32)
        gr2m_count_launch(), k_mask<float><<<ceil_div(total, 256), 256, 0, p->stream>>>(p->level.as<int32_t>(), p->contam.as<uint8_t>(),
                                                                  use_contam, p->ncell, (float *)WA_dev, ld, nmonth);
    else
        gr2m_count_launch(), k_mask<double><<<ceil_div(total, 256), 256, 0, p->stream>>>(p->level.as<int32_t>(), p->contam.as<uint8_t>(),
                                                                   use_contam, p->ncell, (double *)WA_dev, ld, nmonth);
    CK(cudaGetLastError());
    return GR2M_OK;
}

extern "C" int wfa_accumulate(wfa_plan *p, int nmonth, const double *W, int flags, double *A)
{
    REQUIRE(p && W && A, "NULL argument");
    REQUIRE(nmonth > 0, "nmonth must be positive");
    CK(cudaSetDevice(p->device));
    const long long n = p->ncell;
    const int ld = ceil_div(nmonth, 4) * 4;
    const size_t esz = (flags & WFA_F32) ? 4 : 8;
    int rc;
    if ((rc = p->work.ensure((size_t)n * nmonth * 8))) return rc;
    if ((rc = p->work2.ensure((size_t)n * ld * esz))) return rc;
    cudaStream_t st = p->stream;
    CK(cudaMemcpyAsync(p->work.p, W, (size_t)n * nmonth * 8, cudaMemcpyHostToDevice, st));
    dim3 grid(ceil_div(n, 32), ceil_div(ld, 32)), blk(32, 8);
    if (flags & WFA_F32) gr2m_count_launch(), k_planes_to_cells<float><<<grid, blk, 0, st>>>(p->work.as<double>(), n, nmonth, ld, p->work2.as<float>());
    else gr2m_count_launch(), k_planes_to_cells<double><<<grid, blk, 0, st>>>(p->work.as<double>(), n, nmonth, ld, p->work2.as<double>());
    CK(cudaGetLastError());
    if ((rc = wfa_accumulate_dev(p, nmonth, ld, p->work2.p, flags))) return rc;
    if ((rc = wfa_mask_dev(p, nmonth, ld, p->work2.p, flags))) return rc;
    if (flags & WFA_F32) gr2m_count_launch(), k_cells_to_planes<float><<<grid, blk, 0, st>>>(p->work2.as<float>(), n, nmonth, ld, p->work.as<double>());
    else gr2m_count_launch(), k_cells_to_planes<double><<<grid, blk, 0, st>>>(p->work2.as<double>(), n, nmonth, ld, p->work.as<double>());
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(A, p->work.p, (size_t)n * nmonth * 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return GR2M_OK;
}

extern "C" int wfa_route(wfa_plan *p, int ntime, int nsub, const double *QS, const int32_t *src_cell,
                         const int32_t *poly_off, const int32_t *poly_cells, int flags, double *QR)
{
    REQUIRE(p && QS && src_cell && poly_off && poly_cells && QR, "NULL argument");
    REQUIRE(ntime > 0 && nsub > 0, "ntime and nsub must be positive");
    const long long n = p->ncell;
    for (int s = 0; s < nsub; ++s) {
        REQUIRE(src_cell[s] >= 0 && src_cell[s] < n, "src_cell out of range (point on surface outside the DEM)");
        REQUIRE(poly_off[s + 1] >= poly_off[s], "poly_off must be non-decreasing");
    }
    REQUIRE(poly_off[0] == 0, "poly_off[0] must be 0");
    const int npc = poly_off[nsub];
    for (int q = 0; q < npc; ++q) REQUIRE(poly_cells[q] >= 0 && poly_cells[q] < n, "poly_cells out of range");
    CK(cudaSetDevice(p->device));
    if (!(flags & WFA_ROUTE_DENSE) && n < (1ll << 30)) {
        // weights sit at nsub cells only: build the month-invariant operator and apply it to all months
        wfa_router *r = nullptr;
        int rc = wfa_router_create(p, nsub, src_cell, poly_off, poly_cells, flags, &r);
        if (rc) return rc;
        rc = wfa_router_apply(r, ntime, QS, QR);
        wfa_router_destroy(r);
        return rc;
    }
    // duplicates: the last assignment wins, as R's `[<-` (Routing:109-113)
    std::vector<int32_t> src(src_cell, src_cell + nsub);
    {
        std::vector<int> idx(nsub);
        for (int s = 0; s < nsub; ++s) idx[s] = s;
        std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return src_cell[a] < src_cell[b]; });
        for (int i = 0; i + 1 < nsub; ++i)
            if (src_cell[idx[i]] == src_cell[idx[i + 1]]) src[idx[i]] = -1;
    }
    // months are swept in chunks so the dense buffer stays within ~16 GB whatever the grid size;
    // chunks of <= 32 months also take the shared-memory-forwarding per-basin kernel
    const size_t esz = (flags & WFA_F32) ? 4 : 8;
    long long budget = 16ll << 30;
    if (const char *ev = getenv("WFA_ROUTE_BUDGET_MB")) budget = std::max(1ll, atoll(ev)) << 20;   // test knob
    long long fit = budget / ((long long)n * (long long)esz);
    int chunk = fit >= ntime ? ntime : (int)std::max<long long>(32, fit / 32 * 32);
    if (chunk > ntime) chunk = ntime;
    const int ld = ceil_div(chunk, 4) * 4;
    int rc;
    if ((rc = p->work2.ensure((size_t)n * ld * esz))) return rc;
    if ((rc = p->qs.ensure((size_t)nsub * ntime * 8))) return rc;
    if ((rc = p->qr.ensure((size_t)nsub * ntime * 8))) return rc;
    if ((rc = p->src.ensure((size_t)nsub * 4))) return rc;
    if ((rc = p->poff.ensure(((size_t)nsub + 1) * 4))) return rc;
    if ((rc = p->pcells.ensure(std::max<size_t>(4, (size_t)npc * 4)))) return rc;
    cudaStream_t st = p->stream;
    CK(cudaMemcpyAsync(p->qs.p, QS, (size_t)nsub * ntime * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(p->src.p, src.data(), (size_t)nsub * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(p->poff.p, poly_off, ((size_t)nsub + 1) * 4, cudaMemcpyHostToDevice, st));
    if (npc) CK(cudaMemcpyAsync(p->pcells.p, poly_cells, (size_t)npc * 4, cudaMemcpyHostToDevice, st));
    const int use_contam = (flags & WFA_NO_CONTAMINATION) ? 0 : 1;
    for (int t0 = 0; t0 < ntime; t0 += chunk) {
        const int nt = std::min(chunk, ntime - t0);
        const int nb = ceil_div((long long)nsub * nt, 256);
        dim3 pg(nsub, ceil_div(nt, 32));
        CK(cudaMemsetAsync(p->work2.p, 0, (size_t)n * ld * esz, st));
        if (flags & WFA_F32) {
            gr2m_count_launch(), k_scatter<float><<<nb, 256, 0, st>>>(p->qs.as<double>(), p->src.as<int32_t>(), nsub, ntime, t0, nt,
                                                 p->work2.as<float>(), ld);
            CK(cudaGetLastError());
            if ((rc = wfa_accumulate_dev(p, nt, ld, p->work2.p, flags))) return rc;
            gr2m_count_launch(), k_polygon_max<float><<<pg, 256, 0, st>>>(p->poff.as<int32_t>(), p->pcells.as<int32_t>(), p->level.as<int32_t>(),
                                                     p->contam.as<uint8_t>(), use_contam, p->work2.as<float>(), ld, ntime,
                                                     t0, nt, p->qr.as<double>());
        } else {
            gr2m_count_launch(), k_scatter<double><<<nb, 256, 0, st>>>(p->qs.as<double>(), p->src.as<int32_t>(), nsub, ntime, t0, nt,
                                                  p->work2.as<double>(), ld);
            CK(cudaGetLastError());
            if ((rc = wfa_accumulate_dev(p, nt, ld, p->work2.p, flags))) return rc;
            gr2m_count_launch(), k_polygon_max<double><<<pg, 256, 0, st>>>(p->poff.as<int32_t>(), p->pcells.as<int32_t>(), p->level.as<int32_t>(),
                                                      p->contam.as<uint8_t>(), use_contam, p->work2.as<double>(), ld,
                                                      ntime, t0, nt, p->qr.as<double>());
        }
        CK(cudaGetLastError());
    }
    CK(cudaMemcpyAsync(QR, p->qr.p, (size_t)nsub * ntime * 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return GR2M_OK;
}
