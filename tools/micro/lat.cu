// latency probes: dependent FP64 add / mul chains, 64-bit shuffle, LDS, on one warp
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* cyc, double a, double b)
{
    __shared__ double sm[64];
    sm[threadIdx.x] = a; sm[threadIdx.x + 32] = b;
    __syncwarp();
    double x = a + threadIdx.x;
    long long t0 = clock64();
#pragma unroll
    for (int i = 0; i < 256; i++) x = x + b;            // DADD chain
    long long t1 = clock64();
#pragma unroll
    for (int i = 0; i < 256; i++) x = x * b;            // DMUL chain
    long long t2 = clock64();
#pragma unroll
    for (int i = 0; i < 256; i++) x = __shfl_up_sync(0xffffffffu, x, 1);   // 64-bit shuffle chain
    long long t3 = clock64();
#pragma unroll
    for (int i = 0; i < 256; i++) x = x - b * x;        // DMUL + DADD (no fma)
    long long t4 = clock64();
    int idx = threadIdx.x;
#pragma unroll
    for (int i = 0; i < 256; i++) { idx = (int)sm[idx & 63] & 63; }    // LDS + cvt chain
    long long t5 = clock64();
#pragma unroll
    for (int i = 0; i < 256; i++) { double y = __shfl_up_sync(0xffffffffu, x, 1); y = (threadIdx.x & 7) ? y : b; x = x - b * y; }  // one sweep-like step
    long long t6 = clock64();
    out[threadIdx.x] = x + idx;
    if (threadIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = t3 - t2; cyc[3] = t4 - t3; cyc[4] = t5 - t4; cyc[5] = t6 - t5; }
}
int main()
{
    double* out; long long* cyc;
    cudaMalloc(&out, 32 * 8); cudaMallocManaged(&cyc, 6 * 8);
    for (int r = 0; r < 2; r++) { k<<<1, 32>>>(out, cyc, 1.0, 1.0000001); cudaDeviceSynchronize(); }
    const char* n[6] = {"DADD", "DMUL", "SHFL64", "DMUL+DADD", "LDS+cvt", "shfl+sel+mul+add"};
    for (int i = 0; i < 6; i++) printf("%-18s %.1f cycles per link\n", n[i], cyc[i] / 256.0);
    return 0;
}
