// common.cuh -- shared host/device helpers for libgr2m_b200.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/gr2m_b200.h"

// ---- error plumbing ---------------------------------------------------------------------------
void gr2m_set_error(const char *fmt, ...);

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            gr2m_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_));  \
            return (e_ == cudaErrorMemoryAllocation) ? GR2M_ERR_NOMEM : GR2M_ERR_CUDA;            \
        }                                                                                          \
    } while (0)

#define REQUIRE(cond, msg)                                                                         \
    do {                                                                                           \
        if (!(cond)) {                                                                             \
            gr2m_set_error("%s: %s", __func__, msg);                                               \
            return GR2M_ERR_INVALID;                                                               \
        }                                                                                          \
    } while (0)

void gr2m_count_launch(void);   // every kernel launch of the library is counted (gr2m_kernel_launches)
int gr2m_check_device(void);   // GR2M_OK, or GR2M_ERR_NODEVICE with the error text set
int gr2m_current_device(void);

struct wfa_router;
int wfa_router_nsub(const wfa_router *r);      // router.cu
int wfa_router_is_f32(const wfa_router *r);
int wfa_router_wait(wfa_router *r);
int wfa_router_gauge_owner(wfa_router *r, int gauge, int *nmax, int32_t *owner_host);   // see router.cu
long long wfa_router_uid(const wfa_router *r);

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// Device memory goes through a small per-process cache of freed blocks (gr2m.cu): handles are created and
// destroyed around every call in the reference's workflow (one Run_/Optim_/Routing_ call = one handle), and
// cudaFree of a few hundred MB costs milliseconds to tenths of a second.  `got` = the block's real size.
int gr2m_dev_alloc(void **p, size_t bytes, size_t *got);
void gr2m_dev_free(void *p, size_t bytes);

// Copies between device memory and (possibly pageable) host memory; large pageable transfers are pipelined through a
// ring of pinned buffers with the host-side memcpy on helper threads (gr2m.cu).
int gr2m_copy_d2h(void *dst_host, const void *src_dev, size_t bytes, cudaStream_t st);
int gr2m_copy_h2d(void *dst_dev, const void *src_host, size_t bytes, cudaStream_t st);

// A handle that gives back many buffers at once synchronises the device once instead of once per buffer:
// while a FreeBatch is alive on the calling thread, gr2m_dev_free() skips its own synchronisation.
struct FreeBatch {
    bool ok;
    FreeBatch();
    ~FreeBatch();
    FreeBatch(const FreeBatch &) = delete;
    FreeBatch &operator=(const FreeBatch &) = delete;
};

// grow-only device buffer owned by a handle (capacity kept across calls)
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes)
    {
        if (bytes <= cap) return GR2M_OK;
        if (p) gr2m_dev_free(p, cap);
        p = nullptr; cap = 0;
        return gr2m_dev_alloc(&p, bytes, &cap);
    }
    void release()
    {
        if (p) gr2m_dev_free(p, cap);
        p = nullptr; cap = 0;
    }
    template <typename T> T *as() { return static_cast<T *>(p); }
};


// ---- warp helpers -----------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ double shfl_xor_f64(double v, int m)
{
    return __shfl_xor_sync(0xffffffffu, v, m);
}

__device__ __forceinline__ double warp_sum_f64(double v)   // fixed butterfly order, all lanes get it
{
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v += shfl_xor_f64(v, m);
    return v;
}

// ---- mbarrier + 1-D bulk async copy (TMA engine; SASS: UBLKCP / SYNCS) -------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// global -> shared, completes `bytes` on the mbarrier; addresses and size multiples of 16
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
#endif  // __CUDACC__
