// How fast can HBM feed many concurrent sequential streams?  Emulates the operand traffic of the tile-wavefront sweeps:
// one warp per "tile", each step reads NA rows of 32 doubles (one per streamed array) and writes NW rows, WS steps per group,
// loads issued DEPTH groups ahead in registers.  Layouts: 0 = separate arrays [array][tile][step][lane] (what the solver has),
// 1 = interleaved [tile][group][array][WS steps][lane] (one contiguous chunk per group).
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o streams streams.cu && ./streams
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

constexpr int WS = 4;

template <int NA, int NW, int DEPTH, int LAYOUT>
__global__ void __launch_bounds__(32) k_streams(const double* __restrict__ in, double* __restrict__ out, int nTiles, int nSteps, int* ticket,
                                                double* sink)
{
    const int lane = threadIdx.x;
    const int nGroups = nSteps / WS;
    const size_t arrStride = (size_t)nTiles * nSteps * 32;
    double acc = 0;
    for (;;) {
        int T = 0;
        if (lane == 0) T = atomicAdd(ticket, 1);
        T = __shfl_sync(0xffffffffu, T, 0);
        if (T >= nTiles) break;
        auto addr = [&](int a, int grp, int s) -> size_t {
            if (LAYOUT == 0) return (size_t)a * arrStride + ((size_t)T * nSteps + grp * WS + s) * 32 + lane;
            return (((size_t)T * nGroups + grp) * NA + a) * (WS * 32) + s * 32 + lane;
        };
        double buf[DEPTH + 1][NA][WS];
        auto load = [&](int grp, int slot) {
#pragma unroll
            for (int a = 0; a < NA; a++)
#pragma unroll
                for (int s = 0; s < WS; s++) buf[slot][a][s] = __ldg(in + addr(a, grp, s));
        };
#pragma unroll
        for (int d = 0; d < DEPTH; d++) if (d < nGroups) load(d, d);
        for (int g0 = 0; g0 < nGroups; g0 += DEPTH + 1) {
#pragma unroll
            for (int d = 0; d <= DEPTH; d++) {
                const int grp = g0 + d;
                if (grp >= nGroups) break;
                if (grp + DEPTH < nGroups) load(grp + DEPTH, (d + DEPTH) % (DEPTH + 1));
                double v = 0;
#pragma unroll
                for (int a = 0; a < NA; a++)
#pragma unroll
                    for (int s = 0; s < WS; s++) v += buf[d][a][s];
                acc += v;
#pragma unroll
                for (int w = 0; w < NW; w++)
#pragma unroll
                    for (int s = 0; s < WS; s++) out[(size_t)w * arrStride + ((size_t)T * nSteps + grp * WS + s) * 32 + lane] = v;
            }
        }
    }
    if (acc == 123.456) *sink = acc;
}

template <int NA, int NW, int DEPTH, int LAYOUT>
void run(const double* in, double* out, int nTiles, int nSteps, int* ticket, double* sink, int blocksPerSM)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9;
    for (int rep = 0; rep < 4; rep++) {
        cudaMemset(ticket, 0, 4);
        cudaEventRecord(e0);
        k_streams<NA, NW, DEPTH, LAYOUT><<<148 * blocksPerSM, 32>>>(in, out, nTiles, nSteps, ticket, sink);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    const double bytes = (double)nTiles * nSteps * 32 * 8 * (NA + NW);
    printf("NA=%d NW=%d depth=%d layout=%d blocks/SM=%2d : %.3f ms  %.0f GB/s  (%s)\n", NA, NW, DEPTH, LAYOUT, blocksPerSM, best,
           bytes / best * 1e-6, cudaGetErrorString(cudaGetLastError()));
}

int main()
{
    const int nTiles = 3125, nSteps = 1016;
    const size_t arr = (size_t)nTiles * nSteps * 32;
    double *in, *out, *sink; int* ticket;
    cudaMalloc(&in, arr * 6 * 8); cudaMalloc(&out, arr * 3 * 8); cudaMalloc(&sink, 8); cudaMalloc(&ticket, 4);
    cudaMemset(in, 0, arr * 6 * 8); cudaMemset(out, 0, arr * 3 * 8);
    for (int b : {8, 12, 16, 24, 32}) {
        run<5, 2, 1, 0>(in, out, nTiles, nSteps, ticket, sink, b);
        run<5, 2, 1, 1>(in, out, nTiles, nSteps, ticket, sink, b);
        run<5, 2, 2, 0>(in, out, nTiles, nSteps, ticket, sink, b);
        run<5, 2, 2, 1>(in, out, nTiles, nSteps, ticket, sink, b);
        run<5, 0, 1, 0>(in, out, nTiles, nSteps, ticket, sink, b);
        run<5, 0, 2, 1>(in, out, nTiles, nSteps, ticket, sink, b);
    }
    return 0;
}
