// gr2m.cu -- GR2M recursion + objective kernels and their C ABI (sm_100a).
//
// Data layout in HBM (owned by gr2m_basin):
//   F      tiled month-major forcing  F[tile][t][var][lane]  doubles, tile = 32 consecutive cells,
//          var 0 = P, 1 = E, t < ntime_pad (ntime rounded up to 16 months, zero padded).  One stage of
//          8 months of one tile is ONE contiguous 4 KiB block, fetched by a single 1-D bulk
//          async copy (TMA engine, UBLKCP) into shared memory and shared by all sets of the block;
//          inside a stage the 32 lanes of a warp read 32 consecutive doubles (coalesced, conflict-free).
//   area, region   per cell;  den[t] = 86.4*nDays[t], conv[t] = 1/den[t]   (Run:229)
// Thread mapping (north_star): one thread per (cell, parameter set); lanes of a warp = 32 cells of a
// tile, warps of a block = SPB parameter sets; S and R live in registers for all months.
#include "gr2m_math.cuh"

#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <algorithm>
#include <atomic>
#include <future>
#include <mutex>
#include <new>
#include <vector>

// ---------------------------------------------------------------------------------------------
// error plumbing / device selection
// ---------------------------------------------------------------------------------------------
// Both per host thread, like the CUDA runtime's own current device: a thread that never called
// gr2m_set_device() works on device 0 and reads its own last error.
static thread_local char g_err[512] = "";
static thread_local int g_device = 0;
static thread_local int g_free_batch = 0;   // > 0: a FreeBatch on this thread has already synchronised the device

void gr2m_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char *gr2m_last_error(void) { return g_err; }

static std::atomic<long long> g_launches{0};
void gr2m_count_launch(void) { g_launches.fetch_add(1, std::memory_order_relaxed); }
extern "C" long long gr2m_kernel_launches(void) { return g_launches.load(std::memory_order_relaxed); }
extern "C" int gr2m_version(void) { return 1; }

extern "C" int gr2m_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

// ---- cache of freed device blocks ------------------------------------------------------------------
namespace {
struct CachedBlock { void *p; size_t bytes; int dev; };
std::mutex g_cache_mu;
std::vector<CachedBlock> g_cache;
size_t g_cache_bytes = 0;
constexpr size_t CACHE_MIN_BLOCK = 256;             // every cudaFree synchronises and can stall for tens of ms
constexpr size_t CACHE_MAX_BLOCK = 4ull << 30;      // the dense routing buffers are never kept
constexpr size_t CACHE_MAX_TOTAL = 8ull << 30;
bool cache_enabled()
{
    static const bool on = getenv("GR2M_NO_CACHE") == nullptr;
    return on;
}
void cache_trim_locked(int dev)     // dev < 0: every device
{
    int cur = 0;
    cudaGetDevice(&cur);
    for (size_t i = 0; i < g_cache.size();) {
        if (dev < 0 || g_cache[i].dev == dev) {
            cudaSetDevice(g_cache[i].dev);
            cudaFree(g_cache[i].p);
            g_cache_bytes -= g_cache[i].bytes;
            g_cache[i] = g_cache.back();
            g_cache.pop_back();
        } else {
            ++i;
        }
    }
    cudaSetDevice(cur);
}
}  // namespace

// ---- guard-band debug mode (GR2M_GUARD=1) --------------------------------------------------------------
// compute-sanitizer is not available on the GPU pool this library is developed on.  As a stand-in for its memcheck
// and initcheck, every handle buffer is allocated with GUARD_BYTES of a known pattern on either side and its payload
// filled with 0xFF (NaN doubles, -1 integers); the bands are checked when the buffer is given back.  An out-of-bounds
// write of any kernel shows up in gr2m_guard_violations(), a read of uninitialised memory as NaN / garbage in the
// parity tests (tests/test_gpu_guard.py runs every kernel family once in this mode).
namespace {
constexpr size_t GUARD_BYTES = 512;
constexpr unsigned char GUARD_PATTERN = 0xA5;
std::atomic<long long> g_guard_violations{0}, g_guard_checked{0};
bool guard_enabled()
{
    static const bool on = getenv("GR2M_GUARD") != nullptr && atoi(getenv("GR2M_GUARD")) != 0;
    return on;
}
}  // namespace

extern "C" long long gr2m_guard_violations(void) { return g_guard_violations.load(); }
extern "C" long long gr2m_guard_checked(void) { return g_guard_checked.load(); }

int gr2m_dev_alloc(void **p, size_t bytes, size_t *got)
{
    *p = nullptr; *got = 0;
    int dev = 0;
    cudaGetDevice(&dev);
    if (guard_enabled()) {
        unsigned char *raw = nullptr;
        const size_t pay = (bytes + 15) & ~(size_t)15;
        cudaError_t e = cudaMalloc(&raw, pay + 2 * GUARD_BYTES);
        if (e == cudaSuccess) e = cudaMemset(raw, GUARD_PATTERN, GUARD_BYTES);
        if (e == cudaSuccess) e = cudaMemset(raw + GUARD_BYTES, 0xFF, pay);
        if (e == cudaSuccess) e = cudaMemset(raw + GUARD_BYTES + pay, GUARD_PATTERN, GUARD_BYTES);
        if (e == cudaSuccess) e = cudaDeviceSynchronize();   // the fills run on the null stream, handles work on non-blocking ones
        if (e != cudaSuccess) {
            gr2m_set_error("guarded cudaMalloc(%zu bytes) -> %s", bytes, cudaGetErrorString(e));
            cudaGetLastError();
            return e == cudaErrorMemoryAllocation ? GR2M_ERR_NOMEM : GR2M_ERR_CUDA;
        }
        *p = raw + GUARD_BYTES;
        *got = bytes;          // exact capacity: the first byte past the request is already guard (after 16-byte rounding)
        return GR2M_OK;
    }
    if (bytes >= CACHE_MIN_BLOCK && cache_enabled()) {
        std::lock_guard<std::mutex> lk(g_cache_mu);
        int best = -1;
        for (size_t i = 0; i < g_cache.size(); ++i) {
            const CachedBlock &b = g_cache[i];
            if (b.dev == dev && b.bytes >= bytes && b.bytes - bytes <= bytes / 4 &&
                (best < 0 || b.bytes < g_cache[best].bytes))
                best = (int)i;
        }
        if (best >= 0) {
            *p = g_cache[best].p; *got = g_cache[best].bytes;
            g_cache_bytes -= g_cache[best].bytes;
            g_cache[best] = g_cache.back();
            g_cache.pop_back();
            return GR2M_OK;
        }
    }
    cudaError_t e = cudaMalloc(p, bytes);
    if (e == cudaErrorMemoryAllocation) {           // give the cached blocks back and try once more
        cudaGetLastError();
        {
            std::lock_guard<std::mutex> lk(g_cache_mu);
            cache_trim_locked(dev);
        }
        e = cudaMalloc(p, bytes);
    }
    if (e != cudaSuccess) {
        *p = nullptr;
        gr2m_set_error("cudaMalloc(%zu bytes) -> %s", bytes, cudaGetErrorString(e));
        cudaGetLastError();
        return e == cudaErrorMemoryAllocation ? GR2M_ERR_NOMEM : GR2M_ERR_CUDA;
    }
    *got = bytes;
    return GR2M_OK;
}

void gr2m_dev_free(void *p, size_t bytes)
{
    if (!p) return;
    if (guard_enabled()) {
        const size_t pay = (bytes + 15) & ~(size_t)15;
        unsigned char *raw = (unsigned char *)p - GUARD_BYTES, host[2 * GUARD_BYTES];
        cudaDeviceSynchronize();
        bool ok = cudaMemcpy(host, raw, GUARD_BYTES, cudaMemcpyDeviceToHost) == cudaSuccess &&
                  cudaMemcpy(host + GUARD_BYTES, raw + GUARD_BYTES + pay, GUARD_BYTES, cudaMemcpyDeviceToHost) == cudaSuccess;
        for (size_t i = 0; ok && i < 2 * GUARD_BYTES; ++i) ok = host[i] == GUARD_PATTERN;
        g_guard_checked.fetch_add(1);
        if (!ok) {
            g_guard_violations.fetch_add(1);
            fprintf(stderr, "[gr2m guard] guard band of a %zu-byte device buffer was overwritten (or could not be read)\n", bytes);
            cudaGetLastError();
        }
        cudaFree(raw);
        return;
    }
    static const bool dbg = getenv("GR2M_CACHE_DEBUG") != nullptr;
    if (dbg) fprintf(stderr, "[gr2m cache] free %zu bytes, cached %zu bytes in %zu blocks\n", bytes, g_cache_bytes, g_cache.size());
    if (bytes >= CACHE_MIN_BLOCK && bytes <= CACHE_MAX_BLOCK && cache_enabled()) {
        // like cudaFree, nothing may still be using the block when it changes hands (a handle that releases
        // many buffers synchronises once, see FreeBatch)
        if (g_free_batch > 0 || cudaDeviceSynchronize() == cudaSuccess) {
            std::lock_guard<std::mutex> lk(g_cache_mu);
            if (g_cache_bytes + bytes <= CACHE_MAX_TOTAL && g_cache.size() < 512) {
                int dev = 0;
                cudaGetDevice(&dev);
                g_cache.push_back({p, bytes, dev});
                g_cache_bytes += bytes;
                return;
            }
        }
    }
    cudaFree(p);
}

FreeBatch::FreeBatch() : ok(cudaDeviceSynchronize() == cudaSuccess) { if (ok) ++g_free_batch; }
FreeBatch::~FreeBatch() { if (ok) --g_free_batch; }

extern "C" int gr2m_release_cached_memory(void)
{
    std::lock_guard<std::mutex> lk(g_cache_mu);
    cache_trim_locked(-1);
    return GR2M_OK;
}

// ---- large copies to / from pageable host memory ---------------------------------------------------
// R's matrices are ordinary (pageable) memory.  A plain cudaMemcpy on them is staged by the driver through one
// buffer with one host thread (~4-10 GB/s); here the staging is explicit: a ring of pinned buffers, the device
// copy of one chunk overlapping the host memcpy of the others on helper threads.
namespace {
constexpr size_t STG_CHUNK = 4u << 20;
constexpr int STG_NBUF = 16;
constexpr size_t STG_MIN_BYTES = 32u << 20;     // below this a plain copy is as good
std::mutex g_stg_mu;
void *g_stg_buf[STG_NBUF] = {};
cudaEvent_t g_stg_ev[STG_NBUF];
bool stg_init_locked()
{
    if (g_stg_buf[0]) return true;
    for (int i = 0; i < STG_NBUF; ++i) {
        if (cudaHostAlloc(&g_stg_buf[i], STG_CHUNK, cudaHostAllocPortable) != cudaSuccess ||
            cudaEventCreateWithFlags(&g_stg_ev[i], cudaEventDisableTiming) != cudaSuccess) {
            cudaGetLastError();
            for (int j = 0; j <= i; ++j) if (g_stg_buf[j]) { cudaFreeHost(g_stg_buf[j]); g_stg_buf[j] = nullptr; }
            return false;
        }
    }
    return true;
}
bool host_is_pinned(const void *p)
{
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeManaged;
}
}  // namespace

// device -> host; returns after the data is in `dst` when staged, otherwise asynchronously on `st` like cudaMemcpyAsync
int gr2m_copy_d2h(void *dst, const void *src_dev, size_t bytes, cudaStream_t st)
{
    static const bool off = getenv("GR2M_NO_STAGING") != nullptr;
    if (bytes < STG_MIN_BYTES || off || host_is_pinned(dst)) {
        CK(cudaMemcpyAsync(dst, src_dev, bytes, cudaMemcpyDeviceToHost, st));
        return GR2M_OK;
    }
    std::lock_guard<std::mutex> lk(g_stg_mu);
    if (!stg_init_locked()) {
        CK(cudaMemcpyAsync(dst, src_dev, bytes, cudaMemcpyDeviceToHost, st));
        return GR2M_OK;
    }
    std::future<void> fut[STG_NBUF];
    const size_t nch = (bytes + STG_CHUNK - 1) / STG_CHUNK;
    cudaError_t err = cudaSuccess;
    for (size_t i = 0; i < nch && err == cudaSuccess; ++i) {
        const int bi = (int)(i % STG_NBUF);
        if (fut[bi].valid()) fut[bi].get();               // the host copy out of this buffer is done
        const size_t o = i * STG_CHUNK, n = bytes - o < STG_CHUNK ? bytes - o : STG_CHUNK;
        err = cudaMemcpyAsync(g_stg_buf[bi], (const char *)src_dev + o, n, cudaMemcpyDeviceToHost, st);
        if (err == cudaSuccess) err = cudaEventRecord(g_stg_ev[bi], st);
        if (err != cudaSuccess) break;
        void *pin = g_stg_buf[bi];
        cudaEvent_t ev = g_stg_ev[bi];
        char *d = (char *)dst + o;
        fut[bi] = std::async(std::launch::async, [pin, ev, d, n]() { cudaEventSynchronize(ev); memcpy(d, pin, n); });
    }
    for (auto &f : fut) if (f.valid()) f.get();
    if (err != cudaSuccess) {
        gr2m_set_error("staged device-to-host copy failed: %s", cudaGetErrorString(err));
        cudaGetLastError();
        return GR2M_ERR_CUDA;
    }
    return GR2M_OK;
}

// host -> device; the host memory may be reused when this returns if staged, otherwise as for cudaMemcpyAsync
int gr2m_copy_h2d(void *dst_dev, const void *src, size_t bytes, cudaStream_t st)
{
    static const bool off = getenv("GR2M_NO_STAGING") != nullptr;
    if (bytes < STG_MIN_BYTES || off || host_is_pinned(src)) {
        CK(cudaMemcpyAsync(dst_dev, src, bytes, cudaMemcpyHostToDevice, st));
        return GR2M_OK;
    }
    std::lock_guard<std::mutex> lk(g_stg_mu);
    if (!stg_init_locked()) {
        CK(cudaMemcpyAsync(dst_dev, src, bytes, cudaMemcpyHostToDevice, st));
        return GR2M_OK;
    }
    const size_t nch = (bytes + STG_CHUNK - 1) / STG_CHUNK;
    std::future<void> fut[STG_NBUF];
    bool used[STG_NBUF] = {};
    auto fill = [&](size_t i) {     // helper thread: wait until the previous device copy out of the buffer is done, then fill it
        const int bi = (int)(i % STG_NBUF);
        const size_t o = i * STG_CHUNK, n = bytes - o < STG_CHUNK ? bytes - o : STG_CHUNK;
        void *pin = g_stg_buf[bi];
        cudaEvent_t ev = g_stg_ev[bi];
        const bool wait = used[bi];
        const char *sp = (const char *)src + o;
        fut[bi] = std::async(std::launch::async, [pin, ev, sp, n, wait]() { if (wait) cudaEventSynchronize(ev); memcpy(pin, sp, n); });
    };
    cudaError_t err = cudaSuccess;
    for (size_t i = 0; i < nch && i < (size_t)STG_NBUF - 1; ++i) fill(i);
    for (size_t i = 0; i < nch && err == cudaSuccess; ++i) {
        const int bi = (int)(i % STG_NBUF);
        fut[bi].get();
        const size_t o = i * STG_CHUNK, n = bytes - o < STG_CHUNK ? bytes - o : STG_CHUNK;
        err = cudaMemcpyAsync((char *)dst_dev + o, g_stg_buf[bi], n, cudaMemcpyHostToDevice, st);
        if (err == cudaSuccess) err = cudaEventRecord(g_stg_ev[bi], st);
        used[bi] = true;
        if (i + STG_NBUF - 1 < nch) fill(i + STG_NBUF - 1);
    }
    for (auto &f : fut) if (f.valid()) f.get();
    if (err == cudaSuccess) err = cudaStreamSynchronize(st);      // the ring may be reused by the next call
    if (err != cudaSuccess) {
        gr2m_set_error("staged host-to-device copy failed: %s", cudaGetErrorString(err));
        cudaGetLastError();
        return GR2M_ERR_CUDA;
    }
    return GR2M_OK;
}

int gr2m_check_device(void)
{
    if (gr2m_device_count() <= 0) {
        gr2m_set_error("no CUDA device available (libgr2m_b200 has no CPU path)");
        return GR2M_ERR_NODEVICE;
    }
    return GR2M_OK;
}

int gr2m_current_device(void) { return g_device; }

extern "C" int gr2m_set_device(int device)
{
    int n = gr2m_device_count();
    if (n <= 0) return gr2m_check_device();
    REQUIRE(device >= 0 && device < n, "device index out of range");
    g_device = device;
    CK(cudaSetDevice(device));
    return GR2M_OK;
}

// ---------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------
constexpr int TM = 16;    // the forcing table is padded to a multiple of this many months
constexpr int SPB = 8;    // parameter sets (warps) per block

// raw [ncell][ntime] (R column-major) -> tiled month-major F[tile][t][var][32]
__global__ void k_tile_forcing(const double *__restrict__ P, const double *__restrict__ E, int ncell, int ntime,
                               int ntime_pad, double *__restrict__ F, int *__restrict__ bad)
{
    __shared__ double sp[32][33], se[32][33];
    int tile = blockIdx.x, t0 = blockIdx.y * 32;
    int lx = threadIdx.x, ly = threadIdx.y;   // 32 x 8
    for (int j = ly; j < 32; j += 8) {
        int c = tile * 32 + j, t = t0 + lx;
        bool ok = c < ncell && t < ntime;
        double pv = ok ? P[(size_t)c * ntime + t] : 0.0, ev = ok ? E[(size_t)c * ntime + t] : 0.0;
        sp[j][lx] = pv;
        se[j][lx] = ev;
        if (!(fabs(pv) < 1e5 && fabs(ev) < 1e5)) *bad = 1;   // mm per month; also catches NaN / Inf
    }
    __syncthreads();
    for (int j = ly; j < 32; j += 8) {
        int t = t0 + j;
        if (t < ntime_pad) {
            size_t base = ((size_t)tile * ntime_pad + t) * 64;
            F[base + lx] = sp[lx][j];
            F[base + 32 + lx] = se[lx][j];
        }
    }
}

__device__ __forceinline__ CellPar load_par(const double *__restrict__ par, int nreg, int reg, double area,
                                            bool forcing_ok)
{
    CellPar p;
    p.X1 = par[reg];                         // RunModel_GR2M.R floors X1, X2 at 1e-2 (NaN stays NaN)
    p.X2 = par[nreg + reg];
    if (p.X1 < 0.01) p.X1 = 0.01;
    if (p.X2 < 0.01) p.X2 = 0.01;
    p.fp = par[2 * nreg + reg];
    p.fe = par[3 * nreg + reg];
    p.iX1 = 1.0 / p.X1;
    p.i2X1 = 2.0 / p.X1;
    p.fp10 = p.fp * 10.0;
    p.fe10 = p.fe * 10.0;
    p.c01 = 0.1 * p.i2X1;
    p.g = p.c01 * 0x1.71547652b82fep+11;     // * 2048/ln2: exp arguments in table-index units
    p.cP = 0.1 * p.iX1;
    p.kappa = 60.0 * p.iX1;
    p.area = area;
    // fast rounding needs |10 f x| < 1e9 (see r_round1); forcing magnitudes are bounded at upload (forcing_ok)
    p.thr = (forcing_ok && fabs(p.fp) < 1e3 && fabs(p.fe) < 1e3) ? GR2M_TIE_THR : -1.0;   // => |10 f x| < 1e9
    return p;
}

// 8 values per lane -> sums over the 32 lanes; lanes with (lane & 3) == 0 return the total of value
// index ((lane>>4)&1)*4 + ((lane>>3)&1)*2 + ((lane>>2)&1).  9 adds + 9 shuffles for 8 reductions,
// fixed association order (deterministic).
__device__ __forceinline__ double warp_reduce8(const double (&v)[8], int lane)
{
    double w[4], u[2];
    bool hi = lane & 16;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        double send = hi ? v[i] : v[i + 4], keep = hi ? v[i + 4] : v[i];
        w[i] = keep + shfl_xor_f64(send, 16);
    }
    hi = lane & 8;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        double send = hi ? w[i] : w[i + 2], keep = hi ? w[i + 2] : w[i];
        u[i] = keep + shfl_xor_f64(send, 8);
    }
    hi = lane & 4;
    double send = hi ? u[0] : u[1], keep = hi ? u[1] : u[0];
    double s = keep + shfl_xor_f64(send, 4);
    s += shfl_xor_f64(s, 2);
    s += shfl_xor_f64(s, 1);
    return s;
}

struct CalibArgs {
    const double *F, *area, *params, *tconv;   // tconv: conv[t] (fast) or den[t] (literal)
    const int *region;
    double *partial;                           // [csplit][nsets][ntime]
    const double *frange;                      // per cell max P, max E (k_cell_range); table-mode kernel only
    int ncell, ntiles, ntime, ntime_pad, nreg, nsets, nsg, tiles_per_split, f32exp, forcing_ok;
    int group_tiles;                           // tiles per canonical partial sum (k_gr2m_calib4)
};

// Calibration-mode recursion: every (cell, set) is simulated for all months; only the outlet sums
// sum_c QS[t, c] per (set, month) leave the SM (Optim:186-196).  Block = SPB sets x (a range of cell
// tiles processed one after another); forcing stages arrive by bulk async copy and are reused by
// all SPB sets; per-set outlet accumulators live in shared memory across the tile loop.
//  * 8-month stages (4 KiB), 4 in flight  -> 53 KiB shared memory per CTA, 4 CTAs (32 warps) per SM;
//  * no block barrier in the month loop: every warp counts itself out of a stage and the LAST warp
//    to leave a stage re-arms its mbarrier and issues the bulk copy that refills it, so fast warps
//    run up to 3 stages ahead of slow ones instead of meeting them every stage.
constexpr int TM2 = 8;
#ifndef GR2M_NST2
#define GR2M_NST2 4
#endif
constexpr int NST2 = GR2M_NST2;
constexpr int STAGE2_DBL = TM2 * 64;

// 2^(j/2048) table in shared memory (16 KiB) -> 67 KiB per CTA, 3 CTAs (24 warps) per SM; the kernel is bound
// by instruction dispatch, not by occupancy (4 CTAs with the 256-entry table were no faster)
constexpr int CALIB_TAB = 2048, CALIB_CTAS = 3;

template <bool FAST, bool F32EXP>
__global__ void __launch_bounds__(SPB * 32, CALIB_CTAS) k_gr2m_calib2(CalibArgs a)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *stage = reinterpret_cast<double *>(smem_raw);   // NST2 * STAGE2_DBL
    double *qacc = stage + NST2 * STAGE2_DBL;               // SPB * ntime_pad
    double *tconv = qacc + SPB * a.ntime_pad;               // ntime_pad
    double *tab = tconv + a.ntime_pad;                      // CALIB_TAB: 2^(j/2048)
    uint64_t *full = reinterpret_cast<uint64_t *>(tab + CALIB_TAB);   // NST2
    int *left = reinterpret_cast<int *>(full + NST2);          // NST2: warps that have left the stage

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int sg = blockIdx.x % a.nsg, cs = blockIdx.x / a.nsg;
    const int set = sg * SPB + warp;
    const bool set_ok = set < a.nsets;
    const int tile0 = cs * a.tiles_per_split;
    const int tile1 = min(a.ntiles, tile0 + a.tiles_per_split);
    const int nchunk = a.ntime_pad / TM2;
    const int total = (tile1 - tile0) * nchunk;

    for (int i = tid; i < SPB * a.ntime_pad; i += blockDim.x) qacc[i] = 0.0;
    for (int i = tid; i < a.ntime_pad; i += blockDim.x) tconv[i] = i < a.ntime ? a.tconv[i] : 1.0;
    for (int i = tid; i < CALIB_TAB; i += blockDim.x) tab[i] = g_exp2_tab2048[i];
    if (tid < NST2) left[tid] = 0;
    if (tid == 0) {
        for (int i = 0; i < NST2; ++i) mbar_init(&full[i], 1);
        mbar_fence_init();
    }
    __syncthreads();

    const double *Fbase = a.F + (size_t)tile0 * a.ntime_pad * 64;
    if (tid == 0) {
        for (int g = 0; g < NST2 && g < total; ++g) {
            mbar_arrive_expect_tx(&full[g], STAGE2_DBL * 8);
            bulk_g2s(stage + g * STAGE2_DBL, Fbase + (size_t)g * STAGE2_DBL, STAGE2_DBL * 8, &full[g]);
        }
    }

    const double *par = a.params + (size_t)(set_ok ? set : 0) * 4 * a.nreg;
    double *myacc = qacc + warp * a.ntime_pad;
    const int ridx = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
    int g = 0;
    for (int tile = tile0; tile < tile1; ++tile) {
        const int cell = tile * 32 + lane;
        const bool cell_ok = cell < a.ncell;
        CellPar p = load_par(par, a.nreg, cell_ok ? a.region[cell] : 0, cell_ok ? a.area[cell] : 0.0, a.forcing_ok);
        // FAST carries both stores divided by X1 (step_fast4); airGR's default start is 0.3 X1, 0.5 X2
        double S = FAST ? 0.3 : 0.3 * p.X1, R = FAST ? 0.5 * p.X2 * p.iX1 : 0.5 * p.X2;
        const double aX1 = p.area * p.X1;
        for (int ch = 0; ch < nchunk; ++ch, ++g) {
            const int st = g % NST2;
            mbar_wait(&full[st], (g / NST2) & 1);
            if (set_ok) {
                const double *sp = stage + st * STAGE2_DBL + lane;
                const double *tc = tconv + ch * TM2;
                double v[8];
#pragma unroll
                for (int m = 0; m < 8; ++m) {
                    double Pr = sp[m * 64], Er = sp[m * 64 + 32];
                    if (FAST) {
                        v[m] = aX1 * step_fast4<F32EXP>(p, Pr, Er, S, R, tab);   // x conv[t] after the reduction
                    } else {
                        StepOut o = step_literal(p, Pr, Er, S, R, F32EXP);
                        v[m] = (p.area * o.Q) / tc[m];
                    }
                }
                double s = warp_reduce8(v, lane);
                if ((lane & 3) == 0) myacc[ch * TM2 + ridx] += FAST ? s * tc[ridx] : s;
            }
            __syncwarp();
            if (lane == 0) {
                // this warp's reads of the stage are complete (their values were consumed above); the fences
                // order them before the count and the count before the refill (release / acquire on left[st])
                __threadfence_block();
                if (atomicAdd(&left[st], 1) == SPB - 1) {
                    __threadfence_block();
                    left[st] = 0;
                    const int gn = g + NST2;
                    if (gn < total) {
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                        mbar_arrive_expect_tx(&full[st], STAGE2_DBL * 8);
                        bulk_g2s(stage + st * STAGE2_DBL, Fbase + (size_t)gn * STAGE2_DBL, STAGE2_DBL * 8, &full[st]);
                    }
                }
            }
        }
    }
    if (set_ok)
        for (int t = lane; t < a.ntime; t += 32)
            a.partial[((size_t)cs * a.nsets + set) * a.ntime + t] = myacc[t];
}

#include "gr2m_calib3.cuh"

struct SimArgs {
    const double *F, *area, *params, *den, *S0, *R0;
    const int *region;
    double *out[5];     // PR, AE, SM, RU, QS, each [ncell][ntime] = R's column-major [ntime x ncell]; NULL = not wanted
    double *S_end, *R_end;
    int ncell, ncell_pad, ntime, ntime_pad, nreg, f32exp, forcing_ok;
    // many parameter sets in one launch (blockIdx.y = set): parameters `param_stride` doubles apart, a set's series
    // `set_col` columns apart in rows of `pitch` doubles; round_qs: QS rounded like Run:231-235 (R's round(, 1))
    long long pitch;
    int param_stride, set_col, round_qs;
};

// Simulation-mode recursion: one parameter vector, every series written (Run:223-229) -- straight into R's
// layout.  A thread owns a subbasin for all months; results of SIM_MB months are parked in a per-warp shared-memory
// tile [series][cell][month] and written out as runs of SIM_MB consecutive months per subbasin (128-byte segments),
// so no transposed copy of the 5 x [ntime x ncell] outputs is ever made.
constexpr int SIM_MB = 16, SIM_WARPS = 4, SIM_LD = SIM_MB + 1;
constexpr int SIM_SMEM_MAX = (32 + SIM_WARPS * 5 * 32 * SIM_LD) * (int)sizeof(double);     // all five series wanted

template <bool FAST>
__global__ void __launch_bounds__(SIM_WARPS * 32) k_gr2m_sim(SimArgs a)
{
    extern __shared__ double sim_smem[];
    double *tab = sim_smem;                                                      // 32: 2^(j/32)
    // only the wanted series get a tile (the routed objective wants QS alone: 17 KiB per block instead of 87, three
    // times the resident warps); slot[v] = position of series v among the wanted ones
    int slot[5], nser = 0;
#pragma unroll
    for (int v = 0; v < 5; ++v) slot[v] = a.out[v] ? nser++ : -1;
    double *tile = sim_smem + 32 + (threadIdx.x >> 5) * (nser * 32 * SIM_LD);    // [nser][32][SIM_LD]
    if (threadIdx.x < 32) tab[threadIdx.x] = c_exp2_tab[threadIdx.x];
    __syncthreads();
    const int cell = blockIdx.x * blockDim.x + threadIdx.x;      // ncell_pad is a multiple of 32: whole warps leave together
    if (cell >= a.ncell_pad) return;
    const bool ok = cell < a.ncell;
    const int tl = cell >> 5, lane = cell & 31;
    CellPar p = load_par(a.params + (size_t)blockIdx.y * a.param_stride, a.nreg, ok ? a.region[cell] : 0, ok ? a.area[cell] : 0.0,
                         a.forcing_ok);
    double S = (a.S0 && ok) ? a.S0[cell] : 0.3 * p.X1;
    double R = (a.R0 && ok) ? a.R0[cell] : 0.5 * p.X2;
    const size_t col0 = (size_t)blockIdx.y * a.set_col;
    const double *__restrict__ f = a.F + (size_t)tl * a.ntime_pad * 64 + lane;
    const double *__restrict__ den = a.den;
    // the recursion is one dependent chain per thread: keep the forcing of the next U months in flight so a
    // month costs its arithmetic, not a memory round trip (ntime_pad rows exist in F; den is guarded)
    constexpr int U = 4;
    double Pn[U], En[U], Dn[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        Pn[u] = __ldg(f + (size_t)u * 64); En[u] = __ldg(f + (size_t)u * 64 + 32);
        Dn[u] = u < a.ntime ? __ldg(den + u) : 1.0;
    }
    const int cell0 = cell - lane;
    for (int tb = 0; tb < a.ntime; tb += SIM_MB) {
#pragma unroll 1
        for (int t0 = tb; t0 < tb + SIM_MB && t0 < a.ntime; t0 += U) {
            double Pc_[U], Ec_[U], Dc_[U];
#pragma unroll
            for (int u = 0; u < U; ++u) { Pc_[u] = Pn[u]; Ec_[u] = En[u]; Dc_[u] = Dn[u]; }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int tn = t0 + U + u;
                if (tn < a.ntime) {
                    Pn[u] = __ldg(f + (size_t)tn * 64); En[u] = __ldg(f + (size_t)tn * 64 + 32); Dn[u] = __ldg(den + tn);
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int t = t0 + u;
                if (t < a.ntime) {
                    StepOut o = FAST ? step_fast(p, Pc_[u], Ec_[u], S, R, a.f32exp, tab)
                                     : step_literal(p, Pc_[u], Ec_[u], S, R, a.f32exp);
                    double *w = tile + lane * SIM_LD + (t - tb);
                    if (slot[0] >= 0) w[slot[0] * 32 * SIM_LD] = o.Pc;
                    if (slot[1] >= 0) w[slot[1] * 32 * SIM_LD] = o.AE;
                    if (slot[2] >= 0) w[slot[2] * 32 * SIM_LD] = S;
                    if (slot[3] >= 0) w[slot[3] * 32 * SIM_LD] = o.Q;
                    if (slot[4] >= 0) {
                        const double qs = (p.area * o.Q) / Dc_[u];
                        w[slot[4] * 32 * SIM_LD] = a.round_qs ? r_round1(qs, fabs(qs) < 1e8 ? GR2M_TIE_THR : -1.0) : qs;
                    }
                }
            }
        }
        __syncwarp();
        // half a warp writes one subbasin's SIM_MB months (consecutive in R's layout); two subbasins per pass
        const int nm = min(SIM_MB, a.ntime - tb);
        const int mo = lane & (SIM_MB - 1), sub = lane / SIM_MB;
#pragma unroll
        for (int v = 0; v < 5; ++v) {
            double *__restrict__ dst = a.out[v];
            if (!dst) continue;
            const double *src = tile + slot[v] * 32 * SIM_LD;
#pragma unroll 4
            for (int r = sub; r < 32; r += 32 / SIM_MB) {
                const int c = cell0 + r;
                if (c < a.ncell && mo < nm) __stcs(dst + (size_t)c * a.pitch + col0 + tb + mo, src[r * SIM_LD + mo]);
            }
        }
        __syncwarp();
    }
    if (ok && gridDim.y == 1) {
        if (a.S_end) a.S_end[cell] = S;
        if (a.R_end) a.R_end[cell] = R;
    }
}

// gauge row of the routed discharge -> [set][month], rounded like Routing:131 (R's round(, 1)) unless told otherwise
__global__ void k_gauge_row(const double *__restrict__ row, long long n, int round1, double *__restrict__ out)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double v = row[i];
    out[i] = round1 ? r_round1(v, fabs(v) < 1e8 ? GR2M_TIE_THR : -1.0) : v;
}

struct ObjArgs {
    const double *partial, *qobs;   // partial [csplit][nsets][ntime]
    double *qt, *of, *crit;         // qt [nsets][ntime] (scratch or output), of [nsets], crit [nsets][3]
    int csplit, nsets, ntime, warmup, which, single_cell, unrounded, has_obs;
};

// Objective (Optim:186-216): one warp per parameter set.  sink = round(sum over cell splits, 2);
// warm-up and NA pairs dropped; hydroGOF KGE("2009") / NSE / rmse with two-pass moments.
__global__ void __launch_bounds__(128) k_objective(ObjArgs a)
{
    const int set = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (set >= a.nsets) return;
    double *qt = a.qt + (size_t)set * a.ntime;
    for (int t = lane; t < a.ntime; t += 32) {
        double s = 0.0;
        for (int c = 0; c < a.csplit; ++c) s += a.partial[((size_t)c * a.nsets + set) * a.ntime + t];
        qt[t] = (a.single_cell || a.unrounded) ? s : r_round_dig(s, 2);
    }
    __syncwarp();
    if (!a.has_obs) return;
    // pass 1: n, means (R's mean(): sum/n then one refinement pass)
    double n = 0.0, ss = 0.0, so = 0.0;
    for (int t = a.warmup + lane; t < a.ntime; t += 32) {
        double s = qt[t], o = a.qobs[t];
        if (!isnan(s) && !isnan(o)) { n += 1.0; ss += s; so += o; }
    }
    n = warp_sum_f64(n); ss = warp_sum_f64(ss); so = warp_sum_f64(so);
    double ms = ss / n, mo = so / n, cs = 0.0, co = 0.0;
    for (int t = a.warmup + lane; t < a.ntime; t += 32) {
        double s = qt[t], o = a.qobs[t];
        if (!isnan(s) && !isnan(o)) { cs += s - ms; co += o - mo; }
    }
    ms += warp_sum_f64(cs) / n;
    mo += warp_sum_f64(co) / n;
    // pass 2: centred second moments and squared errors
    double sse = 0.0, den = 0.0, vs = 0.0, vo = 0.0, cv = 0.0;
    for (int t = a.warmup + lane; t < a.ntime; t += 32) {
        double s = qt[t], o = a.qobs[t];
        if (!isnan(s) && !isnan(o)) {
            double d = o - s, x = s - ms, y = o - mo;
            sse += d * d; den += y * y; vs += x * x; vo += y * y; cv += x * y;
        }
    }
    sse = warp_sum_f64(sse); den = warp_sum_f64(den); vs = warp_sum_f64(vs); vo = warp_sum_f64(vo);
    cv = warp_sum_f64(cv);
    // rmse = sqrt(mean(d^2)) with R's two-pass mean
    double mse = sse / n, cm = 0.0;
    for (int t = a.warmup + lane; t < a.ntime; t += 32) {
        double s = qt[t], o = a.qobs[t];
        if (!isnan(s) && !isnan(o)) { double d = s - o; cm += d * d - mse; }
    }
    mse += warp_sum_f64(cm) / n;
    if (lane == 0) {
        double nse = NAN, kge = NAN, rmse = NAN;
        if (n > 0.0) {
            if (den != 0.0) nse = 1.0 - sse / den;
            rmse = sqrt(mse);
            if (n > 1.0) {
                double sd_s = sqrt(vs / (n - 1.0)), sd_o = sqrt(vo / (n - 1.0));
                // cor() clamps to [-1, 1] and is NA when either standard deviation is 0; the comparisons keep a
                // NaN (fmin/fmax would return the other operand and turn it into r = -1)
                double r = cv / (sqrt(vs) * sqrt(vo));
                if (r > 1.0) r = 1.0;
                if (r < -1.0) r = -1.0;
                double alpha = sd_s / sd_o, beta = ms / mo;
                if (mo != 0.0 || sd_o != 0.0) {
                    double x = r - 1.0, y = alpha - 1.0, z = beta - 1.0;
                    kge = 1.0 - sqrt(x * x + y * y + z * z);
                }
            }
        }
        if (a.crit) { a.crit[3 * set] = kge; a.crit[3 * set + 1] = nse; a.crit[3 * set + 2] = rmse; }
        if (a.of) {
            double v = a.which == GR2M_OF_KGE ? 1.0 - r_round_dig(kge, 3)
                     : a.which == GR2M_OF_NSE ? 1.0 - r_round_dig(nse, 3) : r_round_dig(rmse, 3);
            a.of[set] = v;
        }
    }
}

// test hook: element-wise device math (op 0 r_round1, 1 r_round1_exact, 2 exp_tab, 3 div_fast,
// 4 rcbrt_fast, 5 r_round_dig(x, (int)y))
__global__ void k_debug_math(int op, int n, const double *x, const double *y, double *out)
{
    __shared__ double tab[32], tab256[256];
    __shared__ __align__(16) unsigned char rcb_raw[2 * 4096];       // the seed table wants a 4 KiB-aligned shared address
    float2 *rcb = reinterpret_cast<float2 *>(rcb_raw + ((0u - smem_u32(rcb_raw)) & 4095u));
    if (threadIdx.x < 32) tab[threadIdx.x] = c_exp2_tab[threadIdx.x];
    tab256[threadIdx.x & 255] = c_exp2_tab256[threadIdx.x & 255];
    for (int j = threadIdx.x; j < 512; j += blockDim.x) rcb[j] = g_rcbrt_lin[j];
    __syncthreads();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double r;
    switch (op) {
        case 0: r = r_round1(x[i], fabs(x[i]) < 1e14 ? GR2M_TIE_THR : -1.0); break;
        case 1: r = r_round1_exact(x[i]); break;
        case 2: r = exp_tab(x[i], tab); break;
        case 3: r = div_fast(x[i], y[i]); break;
        case 4: r = rcbrt_fast(x[i]); break;
        case 6: r = div_fast1(x[i], y[i]); break;
        case 7: r = div_fast3(x[i], y[i]); break;
        case 8: r = exp_tab256(x[i], tab256); break;
        case 9: r = r_round1_cvt(x[i], fabs(x[i]) < 1e14 ? GR2M_TIE_THR : -1.0); break;
        case 10: r = exp_tab2048(x[i], g_exp2_tab2048); break;
        case 11: r = 2.0 * exp_idx2048_half(clamp_u26(x[i] * 0x1.71547652b82fep+11), g_exp2_tab2048); break;
        case 12: r = rcbrt_fast_poly(x[i]); break;
        case 13: r = div_fast3q(x[i], y[i]); break;
        case 14: r = rcbrt_fast_tab(x[i], smem_u32(rcb)); break;
        default: r = r_round_dig(x[i], (int)y[i]); break;
    }
    out[i] = r;
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------

struct gr2m_basin {
    int device = 0, sm_count = 0, ntime = 0, ntime_pad = 0, ncell = 0, ncell_pad = 0, ntiles = 0, nreg = 0;
    cudaStream_t stream = nullptr, own_stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool timed = false;
    int forcing_ok = 0;   // every |P|, |E| < 1e5 and finite (checked on the device at upload)
    DevBuf F, frange, area, region, den, conv, params, partial, qt, of, crit, qobs, sim_out, s0, r0, send, rend;
    DevBuf region_area, region_params;  // several regions, cells mostly grouped by region: areas masked per region [nreg][ncell]
                                        // (table-mode recursion, one pass per region) and the per-region parameters of a batch
    DevBuf gauge_area;                  // routed objective, fused path: areas masked per maximal candidate of the last gauge
    long long gauge_router = 0;         // ... keyed on (router uid, gauge polygon)
    int gauge_poly = -1, gauge_nk = 0;
    double *qobs_host_copy = nullptr;   // last uploaded qobs (to skip re-uploads)
    int qobs_n = 0;
};

static void basin_free(gr2m_basin *b)
{
    if (!b) return;
    cudaSetDevice(b->device);
    FreeBatch fb;
    DevBuf *bufs[] = {&b->F, &b->frange, &b->area, &b->region, &b->den, &b->conv, &b->params, &b->partial, &b->qt, &b->of,
                      &b->crit, &b->qobs, &b->sim_out, &b->s0, &b->r0, &b->send, &b->rend, &b->gauge_area, &b->region_area,
                      &b->region_params};
    for (DevBuf *d : bufs) d->release();
    if (b->ev0) cudaEventDestroy(b->ev0);
    if (b->ev1) cudaEventDestroy(b->ev1);
    if (b->own_stream) cudaStreamDestroy(b->own_stream);
    free(b->qobs_host_copy);
    delete b;
}

// on_device: P and E are device pointers (same layout), written on `src_stream` (NULL = already complete)
static int basin_create_common(int ntime, int ncell, const double *P, const double *E, const double *area,
                               const int *ndays, const int *region_id, int nreg, bool on_device,
                               cudaStream_t src_stream, gr2m_basin **out)
{
    REQUIRE(out != nullptr, "out is NULL");
    *out = nullptr;
    REQUIRE(ntime > 0 && ncell > 0 && nreg > 0, "ntime, ncell and nreg must be positive");
    REQUIRE(P && E && area && ndays && region_id, "NULL input array");
    for (int c = 0; c < ncell; ++c) REQUIRE(region_id[c] >= 0 && region_id[c] < nreg, "region_id out of range");
    for (int t = 0; t < ntime; ++t) REQUIRE(ndays[t] > 0, "ndays must be positive");
    int rc = gr2m_check_device();
    if (rc) return rc;
    CK(cudaSetDevice(g_device));
    gr2m_basin *b = new (std::nothrow) gr2m_basin();
    if (!b) { gr2m_set_error("host allocation failed"); return GR2M_ERR_NOMEM; }
    b->device = g_device;
    b->ntime = ntime; b->ncell = ncell; b->nreg = nreg;
    b->ntime_pad = ceil_div(ntime, TM) * TM;
    b->ntiles = ceil_div(ncell, 32);
    b->ncell_pad = b->ntiles * 32;
#define BCK(call) do { int rc_ = (call); if (rc_) { basin_free(b); return rc_; } } while (0)
#define BCU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { gr2m_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); basin_free(b); return e_ == cudaErrorMemoryAllocation ? GR2M_ERR_NOMEM : GR2M_ERR_CUDA; } } while (0)
    BCU(cudaStreamCreateWithFlags(&b->own_stream, cudaStreamNonBlocking));
    b->stream = b->own_stream;
    BCU(cudaEventCreate(&b->ev0));
    BCU(cudaEventCreate(&b->ev1));
    size_t nraw = (size_t)ncell * ntime * sizeof(double);
    BCK(b->F.ensure((size_t)b->ntiles * b->ntime_pad * 64 * sizeof(double)));
    BCK(b->frange.ensure((size_t)ncell * 2 * sizeof(double)));
    BCK(b->area.ensure(ncell * sizeof(double)));
    BCK(b->region.ensure(ncell * sizeof(int)));
    BCK(b->den.ensure(ntime * sizeof(double)));
    BCK(b->conv.ensure(ntime * sizeof(double) + 16));   // + a device flag used while tiling
    DevBuf rawP, rawE;
    if (on_device && src_stream) b->stream = src_stream;      // tile in order after the producer of P and E
    int r1 = on_device ? 0 : rawP.ensure(nraw), r2 = on_device ? 0 : rawE.ensure(nraw);
    if (r1 || r2) { rawP.release(); rawE.release(); basin_free(b); return r1 ? r1 : r2; }
    const double *Pd = on_device ? P : rawP.as<double>(), *Ed = on_device ? E : rawE.as<double>();
    double *den = (double *)malloc(2 * ntime * sizeof(double));
    if (!den) { rawP.release(); rawE.release(); basin_free(b); gr2m_set_error("host allocation failed"); return GR2M_ERR_NOMEM; }
    for (int t = 0; t < ntime; ++t) { den[t] = 86.4 * (double)ndays[t]; den[ntime + t] = 1.0 / den[t]; }
    cudaError_t e = cudaSuccess;
    auto step = [&](cudaError_t x) { if (e == cudaSuccess) e = x; };
    if (!on_device) {
        if (gr2m_copy_h2d(rawP.p, P, nraw, b->stream) || gr2m_copy_h2d(rawE.p, E, nraw, b->stream)) step(cudaErrorUnknown);
    }
    step(cudaMemcpyAsync(b->area.p, area, ncell * sizeof(double), cudaMemcpyHostToDevice, b->stream));
    step(cudaMemcpyAsync(b->region.p, region_id, ncell * sizeof(int), cudaMemcpyHostToDevice, b->stream));
    // Several regions: the table-mode recursion needs one X1 per warp, i.e. one region per pair of cell tiles (64
    // consecutive subbasins).  It is run once per region with the other regions' areas set to zero (pairs without a
    // cell of the region are skipped, k_gr2m_calib4<MASKED>) -- worth it when the subbasins are mostly grouped by region,
    // so that the passes together touch little more than every pair once; otherwise the general kernel serves.
    std::vector<double> rarea;
    if (nreg > 1 && nreg <= 64 && e == cudaSuccess) {
        const int npairs = ceil_div(b->ntiles, 2);
        long long touched = 0;
        std::vector<unsigned long long> seen(npairs, 0ull);
        for (int c = 0; c < ncell; ++c) seen[c >> 6] |= 1ull << region_id[c];
        for (int p = 0; p < npairs; ++p) touched += __builtin_popcountll(seen[p]);
        if (touched * 4 <= (long long)npairs * 5 + 4 * nreg) {
            rarea.assign((size_t)nreg * ncell, 0.0);
            for (int c = 0; c < ncell; ++c) rarea[(size_t)region_id[c] * ncell + c] = area[c];
            int rr = b->region_area.ensure(rarea.size() * sizeof(double));
            if (rr) { free(den); rawP.release(); rawE.release(); basin_free(b); return rr; }
            step(cudaMemcpyAsync(b->region_area.p, rarea.data(), rarea.size() * sizeof(double), cudaMemcpyHostToDevice, b->stream));
        }
    }
    step(cudaMemcpyAsync(b->den.p, den, ntime * sizeof(double), cudaMemcpyHostToDevice, b->stream));
    step(cudaMemcpyAsync(b->conv.p, den + ntime, ntime * sizeof(double), cudaMemcpyHostToDevice, b->stream));
    if (e == cudaSuccess) {
        dim3 grid(b->ntiles, ceil_div(b->ntime_pad, 32)), blk(32, 8);
        int *bad = reinterpret_cast<int *>(b->conv.as<char>() + (size_t)ntime * sizeof(double));
        step(cudaMemsetAsync(bad, 0, sizeof(int), b->stream));
        gr2m_count_launch(), k_tile_forcing<<<grid, blk, 0, b->stream>>>(Pd, Ed, ncell, ntime, b->ntime_pad, b->F.as<double>(), bad);
        step(cudaGetLastError());
        gr2m_count_launch(), k_cell_range<<<ceil_div(ncell, 128), 128, 0, b->stream>>>(b->F.as<double>(), ncell, ntime, b->ntime_pad,
                                                                  b->frange.as<double>());
        step(cudaGetLastError());
        int hbad = 1;
        step(cudaMemcpyAsync(&hbad, bad, sizeof(int), cudaMemcpyDeviceToHost, b->stream));
        step(cudaStreamSynchronize(b->stream));
        b->forcing_ok = hbad ? 0 : 1;
    }
    step(cudaStreamSynchronize(b->stream));
    b->stream = b->own_stream;
    free(den);
    rawP.release(); rawE.release();
    if (e != cudaSuccess) {
        gr2m_set_error("gr2m_basin_create: upload failed: %s", cudaGetErrorString(e));
        basin_free(b);
        return GR2M_ERR_CUDA;
    }
    *out = b;
    return GR2M_OK;
}

extern "C" int gr2m_basin_create(int ntime, int ncell, const double *P, const double *E, const double *area,
                                 const int *ndays, const int *region_id, int nreg, gr2m_basin **out)
{
    return basin_create_common(ntime, ncell, P, E, area, ndays, region_id, nreg, false, nullptr, out);
}

extern "C" int gr2m_basin_create_dev(int ntime, int ncell, const double *P_dev, const double *E_dev, const double *area,
                                     const int *ndays, const int *region_id, int nreg, void *cuda_stream,
                                     gr2m_basin **out)
{
    return basin_create_common(ntime, ncell, P_dev, E_dev, area, ndays, region_id, nreg, true, (cudaStream_t)cuda_stream,
                               out);
}

extern "C" int gr2m_basin_destroy(gr2m_basin *b)
{
    basin_free(b);
    return GR2M_OK;
}

extern "C" int gr2m_basin_set_stream(gr2m_basin *b, void *cuda_stream)
{
    REQUIRE(b != nullptr, "handle is NULL");
    b->stream = cuda_stream ? (cudaStream_t)cuda_stream : b->own_stream;
    return GR2M_OK;
}

extern "C" int gr2m_basin_last_kernel_ms(gr2m_basin *b, float *ms)
{
    REQUIRE(b && ms, "NULL argument");
    REQUIRE(b->timed, "no timed call yet");
    CK(cudaSetDevice(b->device));
    CK(cudaEventSynchronize(b->ev1));
    CK(cudaEventElapsedTime(ms, b->ev0, b->ev1));
    return GR2M_OK;
}

extern "C" int gr2m_run(gr2m_basin *b, const double *params, const double *S0, const double *R0, int flags,
                        double *PR, double *AE, double *SM, double *RU, double *QS, double *S_end, double *R_end)
{
    REQUIRE(b && params, "NULL handle or params");
    REQUIRE((S0 == nullptr) == (R0 == nullptr), "S0 and R0 must both be given or both be NULL");
    CK(cudaSetDevice(b->device));
    const size_t plane = (size_t)b->ntime * b->ncell;
    double *outs[5] = {PR, AE, SM, RU, QS};
    int nout = 0, slot[5];
    for (int v = 0; v < 5; ++v) slot[v] = outs[v] ? nout++ : -1;
    int rc;
    if ((rc = b->params.ensure(4 * b->nreg * sizeof(double)))) return rc;
    if (nout && (rc = b->sim_out.ensure(nout * plane * sizeof(double)))) return rc;
    if ((rc = b->send.ensure(b->ncell * sizeof(double)))) return rc;
    if ((rc = b->rend.ensure(b->ncell * sizeof(double)))) return rc;
    CK(cudaMemcpyAsync(b->params.p, params, 4 * b->nreg * sizeof(double), cudaMemcpyHostToDevice, b->stream));
    if (S0) {
        if ((rc = b->s0.ensure(b->ncell * sizeof(double)))) return rc;
        if ((rc = b->r0.ensure(b->ncell * sizeof(double)))) return rc;
        CK(cudaMemcpyAsync(b->s0.p, S0, b->ncell * sizeof(double), cudaMemcpyHostToDevice, b->stream));
        CK(cudaMemcpyAsync(b->r0.p, R0, b->ncell * sizeof(double), cudaMemcpyHostToDevice, b->stream));
    }
    SimArgs a;
    a.F = b->F.as<double>(); a.area = b->area.as<double>(); a.params = b->params.as<double>();
    a.den = b->den.as<double>(); a.S0 = S0 ? b->s0.as<double>() : nullptr; a.R0 = R0 ? b->r0.as<double>() : nullptr;
    a.region = b->region.as<int>();
    for (int v = 0; v < 5; ++v) a.out[v] = outs[v] ? b->sim_out.as<double>() + slot[v] * plane : nullptr;
    a.S_end = b->send.as<double>(); a.R_end = b->rend.as<double>();
    a.ncell = b->ncell; a.ncell_pad = b->ncell_pad; a.ntime = b->ntime; a.ntime_pad = b->ntime_pad;
    a.nreg = b->nreg; a.f32exp = (flags & GR2M_PERC_F32_EXPONENT) ? 1 : 0; a.forcing_ok = b->forcing_ok;
    a.pitch = b->ntime; a.param_stride = 0; a.set_col = 0; a.round_qs = 0;
    const size_t smem = (32 + SIM_WARPS * (nout ? nout : 1) * 32 * SIM_LD) * sizeof(double);   // a tile per wanted series
    auto kern = (flags & GR2M_MATH_LITERAL) ? k_gr2m_sim<false> : k_gr2m_sim<true>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SIM_SMEM_MAX));
    CK(cudaEventRecord(b->ev0, b->stream));
    gr2m_count_launch(), kern<<<ceil_div(b->ncell_pad, SIM_WARPS * 32), SIM_WARPS * 32, smem, b->stream>>>(a);
    CK(cudaGetLastError());
    CK(cudaEventRecord(b->ev1, b->stream));
    b->timed = true;
    // the series are already in the caller's layout: straight (staged when large and pageable) copies
    for (int v = 0; v < 5; ++v)
        if (outs[v] && (rc = gr2m_copy_d2h(outs[v], a.out[v], plane * sizeof(double), b->stream))) return rc;
    if (S_end) CK(cudaMemcpyAsync(S_end, b->send.p, b->ncell * sizeof(double), cudaMemcpyDeviceToHost, b->stream));
    if (R_end) CK(cudaMemcpyAsync(R_end, b->rend.p, b->ncell * sizeof(double), cudaMemcpyDeviceToHost, b->stream));
    CK(cudaStreamSynchronize(b->stream));
    return GR2M_OK;
}

extern "C" int gr2m_run_once(int ntime, int ncell, const double *P, const double *E, const double *area,
                             const int *ndays, const int *region_id, int nreg, const double *params,
                             const double *S0, const double *R0, int flags, double *PR, double *AE, double *SM,
                             double *RU, double *QS, double *S_end, double *R_end)
{
    gr2m_basin *b = nullptr;
    int rc = gr2m_basin_create(ntime, ncell, P, E, area, ndays, region_id, nreg, &b);
    if (rc) return rc;
    rc = gr2m_run(b, params, S0, R0, flags, PR, AE, SM, RU, QS, S_end, R_end);
    gr2m_basin_destroy(b);
    return rc;
}

// qobs -> device (cached: re-uploaded only when the host values change)
static int upload_qobs(gr2m_basin *b, const double *qobs)
{
    if (!qobs) return GR2M_OK;
    int rc;
    bool same = b->qobs_host_copy && b->qobs_n == b->ntime && memcmp(b->qobs_host_copy, qobs, b->ntime * sizeof(double)) == 0;
    if (!same) {
        if ((rc = b->qobs.ensure(b->ntime * sizeof(double)))) return rc;
        free(b->qobs_host_copy);
        b->qobs_host_copy = (double *)malloc(b->ntime * sizeof(double));
        if (!b->qobs_host_copy) { gr2m_set_error("host allocation failed"); return GR2M_ERR_NOMEM; }
        memcpy(b->qobs_host_copy, qobs, b->ntime * sizeof(double));
        b->qobs_n = b->ntime;
        CK(cudaMemcpyAsync(b->qobs.p, b->qobs_host_copy, b->ntime * sizeof(double), cudaMemcpyHostToDevice, b->stream));
    }
    return GR2M_OK;
}

// params [nsets][X1 x nreg | X2 x nreg | Fpp x nreg | Fpe x nreg] (Run:85-88) -> out [nreg][nsets][X1, X2, Fpp, Fpe]
__global__ void k_region_params(const double *__restrict__ params, int nsets, int nreg, double *__restrict__ out)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)nreg * nsets * 4) return;
    const int j = (int)(i & 3), set = (int)((i >> 2) % nsets), r = (int)((i >> 2) / nsets);
    out[i] = params[(size_t)set * 4 * nreg + (size_t)j * nreg + r];
}

static int launch_objective(gr2m_basin *b, const double *params_dev, int nsets, const double *qobs, int warmup,
                            int which, int flags, double *of_dev, double *crit_dev, double *qt_dev)
{
    int rc;
    if (!b->sm_count) CK(cudaDeviceGetAttribute(&b->sm_count, cudaDevAttrMultiProcessorCount, b->device));
    const bool has_obs = qobs != nullptr;
    if ((rc = upload_qobs(b, qobs))) return rc;
    const bool fast = !(flags & GR2M_MATH_LITERAL);
    // one region: per-set power tables (gr2m_calib3.cuh); several regions: a lane's X1 depends on its cell
    static const bool no_tab = getenv("GR2M_CALIB_GENERIC") != nullptr;
    const bool tabmulti = fast && b->nreg > 1 && !no_tab && b->region_area.p != nullptr;     // one table-mode pass per region
    const bool tabmode = fast && !no_tab && (b->nreg == 1 || tabmulti);
    const int nsg = ceil_div(nsets, SPB);
    int csplit, tiles_per_split, group_tiles = 0;
    if (tabmode) {
        // canonical groups of tiles (a function of the number of cells only): one partial sum per group, set and month,
        // so a set's outlet sums do not depend on the batch it is evaluated in nor on how many ranks share the sets
        group_tiles = 4;                 // a multiple of the cells per thread of every k_gr2m_calib4 variant
        while (group_tiles * 64 < b->ntiles) group_tiles *= 2;       // at most 64 groups (128 measured 2-5 % slower on C4)
        const int ngroups = ceil_div(b->ntiles, group_tiles);
        int gps = (int)((long long)ngroups * nsg / ((long long)b->sm_count * 3 * 16));   // >= ~16 waves of blocks when possible
        if (const char *ev = getenv("GR2M_CALIB_GPS")) gps = atoi(ev);      // experiment switch: canonical groups per block
        if (gps < 1) gps = 1;
        tiles_per_split = gps * group_tiles;
        csplit = ngroups;
    } else {
        // general kernel (several regions, literal arithmetic): one block per set group and canonical group of tiles, the
        // same grouping rule, so these sums do not depend on the batch size either
        group_tiles = 4;
        while (group_tiles * 64 < b->ntiles) group_tiles *= 2;       // at most 64 groups (128 measured 2-5 % slower on C4)
        tiles_per_split = group_tiles;
        csplit = ceil_div(b->ntiles, tiles_per_split);
    }
    if ((rc = b->partial.ensure((size_t)csplit * nsets * b->ntime * sizeof(double)))) return rc;
    if (!qt_dev) {
        if ((rc = b->qt.ensure((size_t)nsets * b->ntime * sizeof(double)))) return rc;
        qt_dev = b->qt.as<double>();
    }
    CalibArgs a;
    a.F = b->F.as<double>(); a.area = b->area.as<double>(); a.params = params_dev;
    a.tconv = fast ? b->conv.as<double>() : b->den.as<double>();
    a.region = b->region.as<int>(); a.partial = b->partial.as<double>();
    a.ncell = b->ncell; a.ntiles = b->ntiles; a.ntime = b->ntime; a.ntime_pad = b->ntime_pad; a.nreg = b->nreg;
    a.nsets = nsets; a.nsg = nsg; a.tiles_per_split = tiles_per_split; a.group_tiles = group_tiles;
    a.f32exp = (flags & GR2M_PERC_F32_EXPONENT) ? 1 : 0;
    a.forcing_ok = b->forcing_ok;
    a.frange = b->frange.as<double>();
    if (tabmode) {
        static const int ctas = getenv("GR2M_CALIB_CTAS") ? atoi(getenv("GR2M_CALIB_CTAS")) : 3;   // experiment switches
        static const int cpt = getenv("GR2M_CALIB_CPT") ? atoi(getenv("GR2M_CALIB_CPT")) : 2;
        static const int magic = getenv("GR2M_CALIB_MAGIC") ? atoi(getenv("GR2M_CALIB_MAGIC")) : 3;
        const int c = tabmulti ? 2 : cpt == 1 ? 1 : cpt == 4 ? 4 : 2;
        const size_t smem = (size_t)(RCB_DBL + CALIB4_NST * c * STAGE2_DBL + SPB * PT_WARP_DBL + 2 * c * SPB * 32) * sizeof(double) +
                            CALIB4_NST * 12 + CALIB4_ALIGN_PAD;
        void (*kern)(CalibArgs) = k_gr2m_calib4<false, 3, 2, 2>;
        if (a.f32exp) kern = k_gr2m_calib4<true, 3, 2, 2>;
        else if (c == 1) kern = ctas == 4 ? k_gr2m_calib4<false, 4, 1, 2> : k_gr2m_calib4<false, 3, 1, 2>;
        else if (c == 4) kern = k_gr2m_calib4<false, 2, 4, 3>;
        else if (ctas == 2) kern = magic == 2 ? k_gr2m_calib4<false, 2, 2, 2> : k_gr2m_calib4<false, 2, 2, 0>;
        else if (magic == 3) kern = k_gr2m_calib4<false, 3, 2, 3>;
        else if (magic == 1) kern = k_gr2m_calib4<false, 3, 2, 1>;
        else if (magic == 0) kern = k_gr2m_calib4<false, 3, 2, 0>;
        const int nblk_split = ceil_div(b->ntiles, tiles_per_split);
        if (tabmulti) {
            kern = a.f32exp ? k_gr2m_calib4<true, 3, 2, 3, false, true> : k_gr2m_calib4<false, 3, 2, 3, false, true>;
            if ((rc = b->region_params.ensure((size_t)b->nreg * nsets * 4 * sizeof(double)))) return rc;
        }
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CK(cudaEventRecord(b->ev0, b->stream));
        CK(cudaMemsetAsync(b->partial.p, 0, (size_t)csplit * nsets * b->ntime * sizeof(double), b->stream));
        if (tabmulti) {
            const long long np4 = (long long)b->nreg * nsets * 4;
            gr2m_count_launch(), k_region_params<<<ceil_div(np4, 256), 256, 0, b->stream>>>(params_dev, nsets, b->nreg,
                                                                                       b->region_params.as<double>());
            a.nreg = 1;
            for (int r = 0; r < b->nreg; ++r) {       // passes add into the same partial words one after another: deterministic
                a.params = b->region_params.as<double>() + (size_t)r * nsets * 4;
                a.area = b->region_area.as<double>() + (size_t)r * b->ncell;
                gr2m_count_launch(), kern<<<nsg * nblk_split, SPB * 32, smem, b->stream>>>(a);
            }
        } else {
            gr2m_count_launch(), kern<<<nsg * nblk_split, SPB * 32, smem, b->stream>>>(a);
        }
        CK(cudaGetLastError());
        CK(cudaEventRecord(b->ev1, b->stream));
    } else {
        const size_t smem = (size_t)(NST2 * STAGE2_DBL + SPB * b->ntime_pad + b->ntime_pad + CALIB_TAB) * sizeof(double) + NST2 * 12;
        REQUIRE(smem <= 200 * 1024, "ntime too large for the shared-memory outlet accumulators");
        auto kern = fast ? (a.f32exp ? k_gr2m_calib2<true, true> : k_gr2m_calib2<true, false>)
                         : (a.f32exp ? k_gr2m_calib2<false, true> : k_gr2m_calib2<false, false>);
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CK(cudaEventRecord(b->ev0, b->stream));
        gr2m_count_launch(), kern<<<nsg * csplit, SPB * 32, smem, b->stream>>>(a);
        CK(cudaGetLastError());
        CK(cudaEventRecord(b->ev1, b->stream));
    }
    ObjArgs o;
    o.partial = b->partial.as<double>(); o.qobs = has_obs ? b->qobs.as<double>() : nullptr;
    o.qt = qt_dev; o.of = of_dev; o.crit = crit_dev;
    o.csplit = csplit; o.nsets = nsets; o.ntime = b->ntime; o.warmup = warmup < 0 ? 0 : warmup; o.which = which;
    o.single_cell = b->ncell == 1; o.unrounded = (flags & GR2M_SINK_UNROUNDED) ? 1 : 0; o.has_obs = has_obs;
    gr2m_count_launch(), k_objective<<<ceil_div((long long)nsets * 32, 128), 128, 0, b->stream>>>(o);
    CK(cudaGetLastError());
    b->timed = true;
    return GR2M_OK;
}

extern "C" int gr2m_objective_batch_dev(gr2m_basin *b, const double *params_dev, int nsets, const double *qobs,
                                        int warmup, int which, int flags, double *out_of_dev,
                                        double *out_crit_dev, double *out_qt_dev)
{
    REQUIRE(b && params_dev, "NULL handle or params");
    REQUIRE(nsets > 0, "nsets must be positive");
    REQUIRE(which >= 0 && which <= 2, "which must be GR2M_OF_KGE, _NSE or _RMSE");
    REQUIRE(warmup < b->ntime, "warmup must be smaller than ntime");
    CK(cudaSetDevice(b->device));
    return launch_objective(b, params_dev, nsets, qobs, warmup, which, flags, out_of_dev, out_crit_dev, out_qt_dev);
}

extern "C" int gr2m_objective_batch(gr2m_basin *b, const double *params, int nsets, const double *qobs, int warmup,
                                    int which, int flags, double *out_of, double *out_crit, double *out_qt)
{
    REQUIRE(b && params, "NULL handle or params");
    REQUIRE(nsets > 0, "nsets must be positive");
    REQUIRE(which >= 0 && which <= 2, "which must be GR2M_OF_KGE, _NSE or _RMSE");
    REQUIRE(warmup < b->ntime, "warmup must be smaller than ntime");
    REQUIRE(qobs || out_qt, "nothing to compute: qobs and out_qt are both NULL");
    CK(cudaSetDevice(b->device));
    int rc;
    const size_t np = (size_t)nsets * 4 * b->nreg;
    if ((rc = b->params.ensure(np * sizeof(double)))) return rc;
    if ((rc = b->of.ensure(nsets * sizeof(double)))) return rc;
    if ((rc = b->crit.ensure((size_t)nsets * 3 * sizeof(double)))) return rc;
    if ((rc = b->qt.ensure((size_t)nsets * b->ntime * sizeof(double)))) return rc;
    CK(cudaMemcpyAsync(b->params.p, params, np * sizeof(double), cudaMemcpyHostToDevice, b->stream));
    rc = launch_objective(b, b->params.as<double>(), nsets, qobs, warmup, which, flags, b->of.as<double>(),
                          b->crit.as<double>(), b->qt.as<double>());
    if (rc) return rc;
    if (out_of && qobs) CK(cudaMemcpyAsync(out_of, b->of.p, nsets * sizeof(double), cudaMemcpyDeviceToHost, b->stream));
    if (out_crit && qobs)
        CK(cudaMemcpyAsync(out_crit, b->crit.p, (size_t)nsets * 3 * sizeof(double), cudaMemcpyDeviceToHost, b->stream));
    if (out_qt)
        CK(cudaMemcpyAsync(out_qt, b->qt.p, (size_t)nsets * b->ntime * sizeof(double), cudaMemcpyDeviceToHost,
                           b->stream));
    CK(cudaStreamSynchronize(b->stream));
    return GR2M_OK;
}

// ---- routed objective at a gauge, fused (no QS / QR planes) ---------------------------------------------------------
// For a fixed gauge polygon the routing is a linear functional of QS: QR_gauge = max over the polygon's maximal candidate
// nodes of the sum of (rounded) QS over the node's upstream subbasins (wfa_router_gauge_owner, router.cu).  So the table-
// mode recursion kernel computes it directly: one launch per maximal candidate with the areas of the other subbasins set
// to zero and, unless GR2M_SINK_UNROUNDED, every subbasin's discharge rounded to 0.1 before the sum (k_gr2m_calib4<RQS>).
__global__ void k_gauge_combine(const double *__restrict__ partial, int nk, int csplit, int nsets, int ntime, int round1,
                                double *__restrict__ out)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;      // set * ntime + month
    if (i >= (long long)nsets * ntime) return;
    double best = 0.0;
    for (int k = 0; k < nk; ++k) {
        const double *pk = partial + (size_t)k * csplit * nsets * ntime;
        double s = 0.0;
        for (int c = 0; c < csplit; ++c) s += pk[(size_t)c * nsets * ntime + i];
        best = (k == 0 || s > best || isnan(s)) ? s : best;
    }
    out[i] = round1 ? r_round1(best, fabs(best) < 1e8 ? GR2M_TIE_THR : -1.0) : best;
}

// returns GR2M_OK with *done = 0 when the fused path does not apply (several regions, literal arithmetic, no or too many
// maximal candidates): the caller then materialises QS and applies the operator
static int routed_fused(gr2m_basin *b, wfa_router *r, int nsets, int gauge, int flags, double *series_dev, int *done)
{
    *done = 0;
    const bool off = getenv("GR2M_ROUTED_MATERIALISE") != nullptr;      // cross-check switch, read at every call
    if (off || b->nreg != 1 || (flags & GR2M_MATH_LITERAL) || getenv("GR2M_CALIB_GENERIC")) return GR2M_OK;
    int rc, nk = 0;
    if (b->gauge_router == wfa_router_uid(r) && b->gauge_poly == gauge) {
        nk = b->gauge_nk;                                         // same router and gauge as the last call: masks are in place
    } else {
        std::vector<int32_t> owner(b->ncell);
        if ((rc = wfa_router_gauge_owner(r, gauge, &nk, owner.data()))) return rc;
        b->gauge_router = 0;
        if (nk >= 1 && nk <= 4) {
            std::vector<double> area_h(b->ncell), masked((size_t)nk * b->ncell);
            CK(cudaMemcpyAsync(area_h.data(), b->area.p, b->ncell * sizeof(double), cudaMemcpyDeviceToHost, b->stream));
            CK(cudaStreamSynchronize(b->stream));
            for (int k = 0; k < nk; ++k)
                for (int c = 0; c < b->ncell; ++c) masked[(size_t)k * b->ncell + c] = owner[c] == k ? area_h[c] : 0.0;
            if ((rc = b->gauge_area.ensure(masked.size() * sizeof(double)))) return rc;
            CK(cudaMemcpyAsync(b->gauge_area.p, masked.data(), masked.size() * sizeof(double), cudaMemcpyHostToDevice, b->stream));
            CK(cudaStreamSynchronize(b->stream));                 // `masked` leaves scope
        }
        b->gauge_router = wfa_router_uid(r); b->gauge_poly = gauge; b->gauge_nk = nk;
    }
    if (nk < 1 || nk > 4) return GR2M_OK;
    const bool unr = (flags & GR2M_SINK_UNROUNDED) != 0;
    int group_tiles = 4;
    while (group_tiles * 64 < b->ntiles) group_tiles *= 2;
    const int ngroups = ceil_div(b->ntiles, group_tiles), nsg = ceil_div(nsets, SPB);
    int gps = (int)((long long)ngroups * nsg / ((long long)b->sm_count * 3 * 16));
    if (gps < 1) gps = 1;
    const size_t per_k = (size_t)ngroups * nsets * b->ntime;
    DevBuf part;
    if ((rc = part.ensure(nk * per_k * sizeof(double)))) { part.release(); return rc; }
    cudaError_t e = cudaMemsetAsync(part.p, 0, nk * per_k * sizeof(double), b->stream);
    CalibArgs a;
    a.F = b->F.as<double>(); a.params = b->params.as<double>(); a.tconv = b->conv.as<double>();
    a.region = b->region.as<int>(); a.frange = b->frange.as<double>();
    a.ncell = b->ncell; a.ntiles = b->ntiles; a.ntime = b->ntime; a.ntime_pad = b->ntime_pad; a.nreg = 1;
    a.nsets = nsets; a.nsg = nsg; a.tiles_per_split = gps * group_tiles; a.group_tiles = group_tiles;
    a.f32exp = (flags & GR2M_PERC_F32_EXPONENT) ? 1 : 0; a.forcing_ok = b->forcing_ok;
    const size_t smem = (size_t)(RCB_DBL + CALIB4_NST * 2 * STAGE2_DBL + SPB * PT_WARP_DBL + 2 * 2 * SPB * 32 + b->ntime_pad + 2) * sizeof(double) +
                        CALIB4_NST * 12 + CALIB4_ALIGN_PAD;
    void (*kern)(CalibArgs) = unr ? (a.f32exp ? k_gr2m_calib4<true, 3, 2, 3, false, true> : k_gr2m_calib4<false, 3, 2, 3, false, true>)
                                  : (a.f32exp ? k_gr2m_calib4<true, 3, 2, 3, true, true> : k_gr2m_calib4<false, 3, 2, 3, true, true>);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int nblk_split = ceil_div(b->ntiles, a.tiles_per_split);
    for (int k = 0; k < nk && e == cudaSuccess; ++k) {
        a.area = b->gauge_area.as<double>() + (size_t)k * b->ncell;
        a.partial = part.as<double>() + (size_t)k * per_k;
        gr2m_count_launch(), kern<<<nsg * nblk_split, SPB * 32, smem, b->stream>>>(a);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) {
        const long long n = (long long)nsets * b->ntime;
        gr2m_count_launch(), k_gauge_combine<<<ceil_div(n, 256), 256, 0, b->stream>>>(part.as<double>(), nk, ngroups, nsets, b->ntime,
                                                                                 unr ? 0 : 1, series_dev);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(b->stream);
    part.release();
    if (e != cudaSuccess) { gr2m_set_error("gr2m_objective_routed_batch (fused): %s", cudaGetErrorString(e)); cudaGetLastError(); return GR2M_ERR_CUDA; }
    *done = 1;
    return GR2M_OK;
}

// Routed objective: the reference's Run -> Routing -> hydroGOF chain for many parameter sets at once.  Every set is
// simulated for every subbasin (QS rounded to 0.1 like Run_GR2MSemiDistr returns it, Run:229-235), the sets x months
// become the columns of ONE application of the month-invariant routing operator (wfa_router, Routing:100-123), the
// routed discharge at the gauge polygon (rounded to 0.1, Routing:131) is scored against qobs (Optim:199-212).
// GR2M_SINK_UNROUNDED skips both roundings.  Sets are processed in chunks of <= ~1 GB of QS.
extern "C" int gr2m_objective_routed_batch(gr2m_basin *b, wfa_router *r, const double *params, int nsets, const double *qobs,
                                           int gauge, int warmup, int which, int flags, double *out_of, double *out_crit,
                                           double *out_qt)
{
    REQUIRE(b && r && params, "NULL handle, router or params");
    REQUIRE(nsets > 0, "nsets must be positive");
    REQUIRE(which >= 0 && which <= 2, "which must be GR2M_OF_KGE, _NSE or _RMSE");
    REQUIRE(warmup < b->ntime, "warmup must be smaller than ntime");
    REQUIRE(qobs || out_qt, "nothing to compute: qobs and out_qt are both NULL");
    REQUIRE(wfa_router_nsub(r) == b->ncell, "the router was built for a different number of subbasins");
    REQUIRE(!wfa_router_is_f32(r), "the routed objective needs an FP64 router");
    REQUIRE(gauge >= 0 && gauge < b->ncell, "gauge polygon index out of range");
    CK(cudaSetDevice(b->device));
    if (!b->sm_count) CK(cudaDeviceGetAttribute(&b->sm_count, cudaDevAttrMultiProcessorCount, b->device));
    int rc;
    const int nt = b->ntime;
    const size_t np = (size_t)nsets * 4 * b->nreg;
    if ((rc = b->params.ensure(np * sizeof(double)))) return rc;
    if ((rc = b->of.ensure(nsets * sizeof(double)))) return rc;
    if ((rc = b->crit.ensure((size_t)nsets * 3 * sizeof(double)))) return rc;
    if ((rc = b->qt.ensure((size_t)nsets * nt * sizeof(double)))) return rc;
    if ((rc = b->partial.ensure((size_t)nsets * nt * sizeof(double)))) return rc;      // gauge series [set][month]
    CK(cudaMemcpyAsync(b->params.p, params, np * sizeof(double), cudaMemcpyHostToDevice, b->stream));
    if ((rc = upload_qobs(b, qobs))) return rc;
    CK(cudaEventRecord(b->ev0, b->stream));
    int fused = 0;
    if ((rc = routed_fused(b, r, nsets, gauge, flags, b->partial.as<double>(), &fused))) return rc;
    long long per_set = (long long)b->ncell * nt * 8;
    int chunk = (int)std::min<long long>(nsets, std::max<long long>(1, (1ll << 30) / per_set));
    DevBuf qs, qr;
    if (!fused && ((rc = qs.ensure((size_t)chunk * per_set)) || (rc = qr.ensure((size_t)chunk * per_set)))) { qs.release(); qr.release(); return rc; }
    const bool unr = (flags & GR2M_SINK_UNROUNDED) != 0;
    const size_t smem = (32 + SIM_WARPS * 1 * 32 * SIM_LD) * sizeof(double);           // QS is the only series wanted
    auto kern = (flags & GR2M_MATH_LITERAL) ? k_gr2m_sim<false> : k_gr2m_sim<true>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SIM_SMEM_MAX);
    for (int s0 = 0; s0 < nsets && e == cudaSuccess && !rc && !fused; s0 += chunk) {
        const int cs = std::min(chunk, nsets - s0);
        SimArgs a;
        a.F = b->F.as<double>(); a.area = b->area.as<double>(); a.params = b->params.as<double>() + (size_t)s0 * 4 * b->nreg;
        a.den = b->den.as<double>(); a.S0 = nullptr; a.R0 = nullptr; a.region = b->region.as<int>();
        for (int v = 0; v < 4; ++v) a.out[v] = nullptr;
        a.out[4] = qs.as<double>();
        a.S_end = nullptr; a.R_end = nullptr;
        a.ncell = b->ncell; a.ncell_pad = b->ncell_pad; a.ntime = nt; a.ntime_pad = b->ntime_pad; a.nreg = b->nreg;
        a.f32exp = (flags & GR2M_PERC_F32_EXPONENT) ? 1 : 0; a.forcing_ok = b->forcing_ok;
        a.pitch = (long long)cs * nt; a.param_stride = 4 * b->nreg; a.set_col = nt; a.round_qs = unr ? 0 : 1;
        gr2m_count_launch(), kern<<<dim3(ceil_div(b->ncell_pad, SIM_WARPS * 32), cs), SIM_WARPS * 32, smem, b->stream>>>(a);
        e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaStreamSynchronize(b->stream);          // the router works on its plan's stream
        if (e != cudaSuccess) break;
        if ((rc = wfa_router_apply_dev(r, cs * nt, qs.as<double>(), qr.as<double>()))) break;
        if ((rc = wfa_router_wait(r))) break;
        const long long n = (long long)cs * nt;
        gr2m_count_launch(), k_gauge_row<<<ceil_div(n, 256), 256, 0, b->stream>>>(qr.as<double>() + (size_t)gauge * cs * nt, n, unr ? 0 : 1,
                                                                             b->partial.as<double>() + (size_t)s0 * nt);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess && !rc) e = cudaStreamSynchronize(b->stream);
    qs.release(); qr.release();
    if (rc) return rc;
    if (e != cudaSuccess) { gr2m_set_error("gr2m_objective_routed_batch: %s", cudaGetErrorString(e)); cudaGetLastError(); return GR2M_ERR_CUDA; }
    ObjArgs o;
    o.partial = b->partial.as<double>(); o.qobs = qobs ? b->qobs.as<double>() : nullptr;
    o.qt = b->qt.as<double>(); o.of = b->of.as<double>(); o.crit = b->crit.as<double>();
    o.csplit = 1; o.nsets = nsets; o.ntime = nt; o.warmup = warmup < 0 ? 0 : warmup; o.which = which;
    o.single_cell = 0; o.unrounded = 1; o.has_obs = qobs != nullptr;       // the series is already rounded the reference's way
    gr2m_count_launch(), k_objective<<<ceil_div((long long)nsets * 32, 128), 128, 0, b->stream>>>(o);
    CK(cudaGetLastError());
    CK(cudaEventRecord(b->ev1, b->stream));
    b->timed = true;
    if (out_of && qobs) CK(cudaMemcpyAsync(out_of, b->of.p, nsets * sizeof(double), cudaMemcpyDeviceToHost, b->stream));
    if (out_crit && qobs) CK(cudaMemcpyAsync(out_crit, b->crit.p, (size_t)nsets * 3 * sizeof(double), cudaMemcpyDeviceToHost, b->stream));
    if (out_qt) CK(cudaMemcpyAsync(out_qt, b->qt.p, (size_t)nsets * nt * sizeof(double), cudaMemcpyDeviceToHost, b->stream));
    CK(cudaStreamSynchronize(b->stream));
    return GR2M_OK;
}

// FP64 pipe peak: 8 independent DFMA chains per thread, no memory traffic.  Returns TFLOP/s
// (2 flop per DFMA).  The roofline denominator for the recursion kernel, measured in place.
__global__ void __launch_bounds__(256) k_fp64_peak(int iters, double seed, double *sink)
{
    double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6,
           a7 = a0 + 7;
    const double m = 0.999999, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            a0 = __fma_rn(a0, m, c); a1 = __fma_rn(a1, m, c); a2 = __fma_rn(a2, m, c); a3 = __fma_rn(a3, m, c);
            a4 = __fma_rn(a4, m, c); a5 = __fma_rn(a5, m, c); a6 = __fma_rn(a6, m, c); a7 = __fma_rn(a7, m, c);
        }
    }
    double r = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if (r == 123.456) sink[0] = r;
}

// Same chains, but every DFMA reads three distinct per-thread register pairs (no constant-bank or
// immediate operands): the register-file-limited rate real FP64 code can sustain.
__global__ void __launch_bounds__(256) k_fp64_peak_rrr(int iters, double seed, double *sink)
{
    double a[8], m[8], c[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {   // operands come from memory so the compiler cannot fold them
        a[u] = seed + threadIdx.x + u;
        m[u] = sink[1 + threadIdx.x * 16 + u];
        c[u] = sink[1 + threadIdx.x * 16 + 8 + u];
    }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int u = 0; u < 8; ++u) a[u] = __fma_rn(a[u], m[u], c[u]);
        }
    }
    double r = ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]));
    if (r == 123.456) sink[0] = r;
}

// operand-mix variants of the same chains (development probe, not in the header):
//   mode 0: fma(a, m, const)   two register operands + constant bank
//   mode 1: a + m              DADD, two register operands
//   mode 2: a * m              DMUL, two register operands
//   mode 3: fma(a, a, c)       two distinct register pairs, one repeated
template <int MODE>
__global__ void __launch_bounds__(256) k_fp64_peak_mix(int iters, double seed, double *sink)
{
    double a[8], m[8], c[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        a[u] = seed + threadIdx.x + u;
        m[u] = sink[1 + threadIdx.x * 16 + u];
        c[u] = sink[1 + threadIdx.x * 16 + 8 + u];
    }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                if (MODE == 0) a[u] = __fma_rn(a[u], m[u], 1e-9);
                if (MODE == 1) a[u] = __dadd_rn(a[u], m[u]);
                if (MODE == 2) a[u] = __dmul_rn(a[u], m[u]);
                if (MODE == 3) a[u] = __fma_rn(a[u], a[u], c[u]);
            }
        }
    }
    double r = ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]));
    if (r == 123.456) sink[0] = r;
}

extern "C" int gr2m_fp64_peak_mix(int mode, int iters, double *tinstr)
{
    REQUIRE(iters > 0 && tinstr && mode >= 0 && mode <= 3, "bad arguments");
    int rc = gr2m_check_device();
    if (rc) return rc;
    CK(cudaSetDevice(g_device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, g_device));
    const int blocks = prop.multiProcessorCount * 8;
    double *sink = nullptr;
    CK(cudaMalloc(&sink, (1 + 256 * 16) * sizeof(double)));
    CK(cudaMemset(sink, 0, (1 + 256 * 16) * sizeof(double)));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        CK(cudaEventRecord(e0));
        if (mode == 0) gr2m_count_launch(), k_fp64_peak_mix<0><<<blocks, 256>>>(iters, 1.0, sink);
        if (mode == 1) gr2m_count_launch(), k_fp64_peak_mix<1><<<blocks, 256>>>(iters, 1.0, sink);
        if (mode == 2) gr2m_count_launch(), k_fp64_peak_mix<2><<<blocks, 256>>>(iters, 1.0, sink);
        if (mode == 3) gr2m_count_launch(), k_fp64_peak_mix<3><<<blocks, 256>>>(iters, 1.0, sink);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    CK(cudaGetLastError());
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(sink);
    *tinstr = 64.0 * (double)iters * 256.0 * blocks / (best * 1e-3) / 1e12;    // 1e12 thread-instructions/s
    return GR2M_OK;
}

extern "C" int gr2m_fp64_peak_rrr(int iters, double *tflops)
{
    REQUIRE(iters > 0 && tflops, "bad arguments");
    int rc = gr2m_check_device();
    if (rc) return rc;
    CK(cudaSetDevice(g_device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, g_device));
    const int blocks = prop.multiProcessorCount * 8;
    double *sink = nullptr;
    CK(cudaMalloc(&sink, (1 + 256 * 16) * sizeof(double)));
    CK(cudaMemset(sink, 0, (1 + 256 * 16) * sizeof(double)));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    gr2m_count_launch(), k_fp64_peak_rrr<<<blocks, 256>>>(iters / 4 + 1, 1.0, sink);
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(e0));
        gr2m_count_launch(), k_fp64_peak_rrr<<<blocks, 256>>>(iters, 1.0, sink);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(sink);
    *tflops = 2.0 * 64.0 * (double)iters * 256.0 * blocks / (best * 1e-3) / 1e12;
    return GR2M_OK;
}

extern "C" int gr2m_fp64_peak(int iters, double *tflops, float *ms_out)
{
    REQUIRE(iters > 0 && tflops, "bad arguments");
    int rc = gr2m_check_device();
    if (rc) return rc;
    CK(cudaSetDevice(g_device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, g_device));
    const int blocks = prop.multiProcessorCount * 8;
    double *sink = nullptr;
    CK(cudaMalloc(&sink, sizeof(double)));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    gr2m_count_launch(), k_fp64_peak<<<blocks, 256>>>(iters / 4 + 1, 1.0, sink);   // warm-up
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(e0));
        gr2m_count_launch(), k_fp64_peak<<<blocks, 256>>>(iters, 1.0, sink);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(sink);
    *tflops = 2.0 * 64.0 * (double)iters * 256.0 * blocks / (best * 1e-3) / 1e12;
    if (ms_out) *ms_out = best;
    return GR2M_OK;
}

// test hook (declared in tests only): fill an n-double handle buffer, writing `extra` doubles past its end
__global__ void k_debug_fill(double *p, int n) { int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) p[i] = 1.0; }
extern "C" int gr2m_debug_overrun(int n, int extra)
{
    REQUIRE(n > 0 && extra >= 0, "bad arguments");
    int rc = gr2m_check_device();
    if (rc) return rc;
    CK(cudaSetDevice(g_device));
    DevBuf b;
    if ((rc = b.ensure((size_t)n * sizeof(double)))) return rc;
    gr2m_count_launch(), k_debug_fill<<<ceil_div(n + extra, 256), 256>>>(b.as<double>(), n + extra);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    b.release();
    return GR2M_OK;
}

// test hook (declared in tests only): element-wise device math, see k_debug_math
extern "C" int gr2m_debug_math(int op, int n, const double *x, const double *y, double *out)
{
    REQUIRE(n > 0 && x && out, "bad arguments");
    int rc = gr2m_check_device();
    if (rc) return rc;
    CK(cudaSetDevice(g_device));
    double *dx = nullptr, *dy = nullptr, *dout = nullptr;
    CK(cudaMalloc(&dx, n * sizeof(double)));
    CK(cudaMalloc(&dy, n * sizeof(double)));
    CK(cudaMalloc(&dout, n * sizeof(double)));
    CK(cudaMemcpy(dx, x, n * sizeof(double), cudaMemcpyHostToDevice));
    if (y) CK(cudaMemcpy(dy, y, n * sizeof(double), cudaMemcpyHostToDevice));
    else CK(cudaMemset(dy, 0, n * sizeof(double)));
    gr2m_count_launch(), k_debug_math<<<ceil_div(n, 256), 256>>>(op, n, dx, dy, dout);
    CK(cudaGetLastError());
    CK(cudaMemcpy(out, dout, n * sizeof(double), cudaMemcpyDeviceToHost));
    cudaFree(dx); cudaFree(dy); cudaFree(dout);
    return GR2M_OK;
}
