// Shared helpers for the sm_100a kernels: error handling, launch counting, deterministic
// block reductions.  All arithmetic is FP64 and compiled with -fmad=false so that a
// product followed by a sum rounds twice, exactly as the reference's x86-64 build does.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <string>
#include <stdexcept>
#include "solve_state.cuh"

namespace b200 {

struct Error : std::runtime_error {
    using std::runtime_error::runtime_error;
};

#define B200_CUDA(expr)                                                                         \
    do {                                                                                        \
        cudaError_t e_ = (expr);                                                                \
        if (e_ != cudaSuccess)                                                                  \
            throw ::b200::Error(std::string(#expr) + ": " + cudaGetErrorString(e_) + " at " +   \
                                __FILE__ + ":" + std::to_string(__LINE__));                     \
    } while (0)

#define B200_CHECK(cond, msg)                                                                   \
    do {                                                                                        \
        if (!(cond)) throw ::b200::Error(std::string(msg) + " (" #cond ") at " + __FILE__ + ":" + \
                                         std::to_string(__LINE__));                             \
    } while (0)

extern int64_t g_launches;     // kernels launched by this library (bench.py "gpu_launches")
inline void count_launch(int n = 1) { g_launches += n; }
#define B200_LAUNCH_CHECK() do { ::b200::count_launch(); B200_CUDA(cudaGetLastError()); } while (0)

// A read-only load that stays where it is written.  The compiler sinks ordinary loads next to (or into the branch of) their
// first use, which turns `load a, b, c, ...; use a; use b; ...` into a chain of load -> wait -> use stages: one memory latency
// per stage instead of one per cell.  A volatile asm is neither sunk nor reordered, so a block of ld_first() calls is issued
// back to back.  Only for data no thread writes during the kernel (ld.global.nc).  Measured: it pays where a thread's loads come
// from different arrays and nothing else hides the latency (the corrector of explicit steps 2.9 -> 1.8 ms, k_faces -8 %).  It
// does NOT pay in the stencil kernels: Amul with all operands -- or one operand per cache line -- loaded first needs 48 instead
// of 32 registers, 5 instead of 8 resident blocks, and went from 0.98 to 1.4-1.8 ms; that kernel lives on its 64 warps per SM.
__device__ __forceinline__ double ld_first(const double* p)
{
    double v;
    asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}

// same, for data the loading thread itself overwrites later in the kernel (plain ld.global)
__device__ __forceinline__ double ld_first_rw(const double* p)
{
    double v;
    asm volatile("ld.global.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}

// ---------------------------------------------------------------------------------------
// Deterministic reductions.  Every reducing kernel runs with a fixed grid; each block
// reduces with a fixed shuffle tree, writes its partial to `partials[slot][block]`, and
// the last block to finish (atomic ticket) sums the partials in a fixed order.  The result
// therefore depends only on (grid, block) geometry, never on scheduling.
// ---------------------------------------------------------------------------------------
constexpr int RED_THREADS = 256;
constexpr int RED_MAX_BLOCKS = 2048;     // partials per slot

struct Sum { __device__ static double id() { return 0.0; } __device__ static double op(double a, double b) { return a + b; } };
struct Max { __device__ static double id() { return -1e300; } __device__ static double op(double a, double b) { return a > b ? a : b; } };
struct Min { __device__ static double id() { return 1e300; } __device__ static double op(double a, double b) { return a < b ? a : b; } };

template <class Op> __device__ inline double warp_reduce(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = Op::op(v, __shfl_down_sync(0xffffffffu, v, o));
    return v;
}

// Reduce `v` over the block (any block size that is a multiple of 32, <= 1024); result valid in thread 0.
template <class Op> __device__ inline double block_reduce(double v, double* smem /* >= 32 doubles */)
{
    const int tid = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
    const int nthreads = blockDim.x * blockDim.y * blockDim.z;
    const int lane = tid & 31, warp = tid >> 5;
    v = warp_reduce<Op>(v);
    __syncthreads();
    if (lane == 0) smem[warp] = v;
    __syncthreads();
    if (warp == 0) {
        const int nw = (nthreads + 31) >> 5;
        v = lane < nw ? smem[lane] : Op::id();
        v = warp_reduce<Op>(v);
    }
    return v;
}

struct RedWork {
    double* partials;      // [nSlots][RED_MAX_BLOCKS]
    unsigned* ticket;      // one counter per reducing kernel in flight (serialised on one stream)
    double* result;        // [nSlots] final values
    // single-rank solves: the Krylov control step that consumes this sum, run by the last block right after it wrote the
    // result (C_NONE: the caller reduces over ranks first, or there is no control step)
    int phase = C_NONE;
    SolveState* state = nullptr;
};

// Called by every thread of every block after computing its private values v[0..N).
// ops[i]: 0 sum, 1 max, 2 min.  Writes result[slot0 + i].
template <int N>
__device__ inline void grid_reduce(double (&v)[N], const int (&ops)[N], RedWork w, int slot0, double* smem)
{
    const int tid = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
    const int nthreads = blockDim.x * blockDim.y * blockDim.z;
    const int bid = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
    const int nblocks = gridDim.x * gridDim.y * gridDim.z;
    __shared__ bool isLast;
#pragma unroll
    for (int i = 0; i < N; i++) {
        double r = ops[i] == 0 ? block_reduce<Sum>(v[i], smem) : ops[i] == 1 ? block_reduce<Max>(v[i], smem) : block_reduce<Min>(v[i], smem);
        if (tid == 0) w.partials[(size_t)(slot0 + i) * RED_MAX_BLOCKS + bid] = r;
    }
    __threadfence();
    if (tid == 0) {
        const unsigned t = atomicAdd(w.ticket, 1u);
        isLast = (t == (unsigned)nblocks - 1u);
    }
    __syncthreads();
    if (!isLast) return;
    __threadfence();
#pragma unroll
    for (int i = 0; i < N; i++) {
        const double* p = w.partials + (size_t)(slot0 + i) * RED_MAX_BLOCKS;
        double acc = ops[i] == 0 ? 0.0 : ops[i] == 1 ? -1e300 : 1e300;
        for (int b = tid; b < nblocks; b += nthreads) {
            const double x = __ldcg(p + b);
            acc = ops[i] == 0 ? acc + x : ops[i] == 1 ? (acc > x ? acc : x) : (acc < x ? acc : x);
        }
        double r = ops[i] == 0 ? block_reduce<Sum>(acc, smem) : ops[i] == 1 ? block_reduce<Max>(acc, smem) : block_reduce<Min>(acc, smem);
        if (tid == 0) w.result[slot0 + i] = r;
    }
    if (tid == 0) {
        *w.ticket = 0u;
        if (w.phase != C_NONE) ctrl_apply(w.phase, w.state, w.result);
    }
}

inline int red_blocks_for(int64_t n, int threads = RED_THREADS)
{
    int64_t b = (n + threads - 1) / threads;
    if (b > RED_MAX_BLOCKS) b = RED_MAX_BLOCKS;
    if (b < 1) b = 1;
    return (int)b;
}

} // namespace b200
