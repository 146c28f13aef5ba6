// Dependent-chain latencies of the instructions the recursion kernel is made of (sm_100a).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/lat scripts/ubench/lat.cu && gpurun_out/lat
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ double rcp_seed(double d) { double r; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d)); return r; }
template <int OP> __global__ void k(double *out, long long *cyc, double x0, double c0, int n)
{
    double x = x0 + threadIdx.x * 1e-9, c = c0;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            if (OP == 0) x = __fma_rn(x, c, c);
            if (OP == 1) x = __dadd_rn(x, c);
            if (OP == 2) x = __dmul_rn(x, c);
            if (OP == 3) x = rcp_seed(x);
            if (OP == 4) x = rint(x * 1.0000001);                       // DMUL + FRND
            if (OP == 5) x = (double)__double2float_rn(x);              // F2F + F2F
            if (OP == 6) { float f = (float)x; asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(f)); x = (double)f; }   // F2F + LG2 + F2F
            if (OP == 7) x = (double)__double2int_rn(x);                // F2I + I2F
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int OP> void run(const char *name, int warps, double x0, double c0)
{
    double *out; long long *cyc, h;
    cudaMalloc(&out, 1024 * 8 * 64); cudaMalloc(&cyc, 8);
    const int n = 2000;
    k<OP><<<1, 32 * warps>>>(out, cyc, x0, c0, 10);
    k<OP><<<1, 32 * warps>>>(out, cyc, x0, c0, n);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-22s warps/SM %2d  cycles per op (per warp chain) %7.2f\n", name, warps, (double)h / (n * 16.0));
    cudaFree(out); cudaFree(cyc);
}
int main()
{
    for (int w : {1, 4, 8, 16, 24, 32}) {
        if (w == 1) printf("-- single warp = dependent latency; more warps -> throughput limit = cycles * 4 / warps per SMSP\n");
        switch (0) { default:
        run<0>("DFMA", w, 1.0, 0.5); run<1>("DADD", w, 1.0, 1e-9); run<2>("DMUL", w, 1.0, 0.9999999);
        run<3>("MUFU.RCP64H(+mov)", w, 1.5, 0); run<4>("DMUL+FRND", w, 123.4, 0); run<5>("F2F.32+F2F.64", w, 1.5, 0);
        run<6>("F2F+LG2+F2F", w, 3.7, 0); run<7>("F2I+I2F", w, 77.0, 0); }
    }
    return 0;
}
