// Micro-benchmark (GPU box): dependent-issue latency of FP64 FMA, MUFU.RCP64H and of a double shuffle on one warp.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_latency fp64_latency.cu && ./fp64_latency
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double *out, long long *cyc, double a, double b) {
    double x = a + threadIdx.x * 1e-9;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 1024; i++) {
        x = fma(x, b, a); x = fma(x, b, a); x = fma(x, b, a); x = fma(x, b, a);
        x = fma(x, b, a); x = fma(x, b, a); x = fma(x, b, a); x = fma(x, b, a);
    }
    long long t1 = clock64();
    double y = a + threadIdx.x;
#pragma unroll 1
    for (int i = 0; i < 1024; i++) {
        y = __shfl_sync(0xffffffffu, y, (threadIdx.x + 1) & 31); y = __shfl_sync(0xffffffffu, y, (threadIdx.x + 1) & 31);
        y = __shfl_sync(0xffffffffu, y, (threadIdx.x + 1) & 31); y = __shfl_sync(0xffffffffu, y, (threadIdx.x + 1) & 31);
    }
    long long t2 = clock64();
    double z = 1.5 + threadIdx.x;
#pragma unroll 1
    for (int i = 0; i < 1024; i++) {
        double r;
        asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(z));
        z = r + 1.5;
    }
    long long t3 = clock64();
    // four independent FMA chains: issue-limited rate of one warp
    double p = x, q = y, r2 = z, s = a;
    long long t4 = clock64();
#pragma unroll 1
    for (int i = 0; i < 1024; i++) {
        p = fma(p, b, a); q = fma(q, b, a); r2 = fma(r2, b, a); s = fma(s, b, a);
        p = fma(p, b, a); q = fma(q, b, a); r2 = fma(r2, b, a); s = fma(s, b, a);
    }
    long long t5 = clock64();
    out[threadIdx.x] = x + y + z + p + q + r2 + s;
    if (threadIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = t3 - t2; cyc[3] = t5 - t4; }
}
int main() {
    double *o; long long *c, h[4];
    cudaMalloc(&o, 32 * 8); cudaMalloc(&c, 4 * 8);
    k<<<1, 32>>>(o, c, 1e-3, 0.999);
    k<<<1, 32>>>(o, c, 1e-3, 0.999);
    cudaMemcpy(h, c, sizeof(h), cudaMemcpyDeviceToHost);
    printf("dependent DFMA: %.2f cycles each\n", h[0] / 8192.0);
    printf("dependent double shuffle (2 SHFL): %.2f cycles each\n", h[1] / 4096.0);
    printf("dependent rcp.approx.f64 + DADD: %.2f cycles per pair\n", h[2] / 1024.0);
    printf("4 independent DFMA chains, one warp: %.2f cycles per DFMA\n", h[3] / 8192.0);
    return 0;
}
