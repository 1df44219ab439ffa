// round-trip latency of "store a double, other SM polls it with ld.relaxed.gpu" between two CTAs
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ double ldr(const double* p) { double v; asm volatile("ld.relaxed.gpu.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory"); return v; }
__global__ void k(double* a, double* b, long long* cyc, int n, int mode)
{
    if (threadIdx.x != 0) return;
    if (blockIdx.x == 0) {
        long long t0 = clock64();
        for (int i = 1; i <= n; i++) {
            if (mode == 0) a[0] = (double)i; else asm volatile("st.relaxed.gpu.global.f64 [%0], %1;" ::"l"(a), "d"((double)i) : "memory");
            while (ldr(b) != (double)i) {}
        }
        cyc[0] = clock64() - t0;
    } else if (blockIdx.x == gridDim.x - 1) {
        for (int i = 1; i <= n; i++) {
            while (ldr(a) != (double)i) {}
            if (mode == 0) b[0] = (double)i; else asm volatile("st.relaxed.gpu.global.f64 [%0], %1;" ::"l"(b), "d"((double)i) : "memory");
        }
    }
}
int main()
{
    double *a, *b; long long* cyc;
    cudaMalloc(&a, 256); cudaMalloc(&b, 256); cudaMallocManaged(&cyc, 8);
    for (int mode = 0; mode < 2; mode++) for (int grid : {2, 148}) {
        cudaMemset(a, 0, 256); cudaMemset(b, 0, 256);
        k<<<grid, 32>>>(a, b, cyc, 2000, mode); cudaDeviceSynchronize();
        printf("mode %d grid %3d: %.0f cycles per round trip (2 hops)\n", mode, grid, cyc[0] / 2000.0);
    }
    return 0;
}
