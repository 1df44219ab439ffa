// What does one step of the sweep's compute warp cost?  One warp, operands in shared memory (no memory warp, no waiting):
// variants strip the step down to find where the cycles go.  nvcc -O3 -fmad=false -arch=sm_100a step.cu -o step
#include <cstdio>
#include <cuda_runtime.h>
constexpr int WTJ = 8, HS = 4, NARR = 5, GSZ = 8, NG = 128;
__device__ __forceinline__ bool is_sentinel(double v) { return __double2hiint(v) == -1; }
__device__ __forceinline__ double publishable(double v) { return is_sentinel(v) ? __longlong_as_double(0x7FF8000000000000ll) : v; }
struct Ops { double c0[HS], c1[HS], c2[HS], c3[HS], in[HS], hy[HS], hz[HS]; };
struct Slot { double arr[NARR][GSZ * 32]; double hy[GSZ * 32], hz[GSZ * 32]; };
__device__ __forceinline__ void fetch(const Slot& S, int s0, int lane, Ops& r)
{
#pragma unroll
    for (int s = 0; s < HS; s++) {
        const int o = (s0 + s) * 32 + lane;
        r.hy[s] = S.hy[o]; r.hz[s] = S.hz[o];
        const double rd = S.arr[0][o];
        r.in[s] = S.arr[4][o];
        r.c0[s] = rd * r.in[s];
        r.c1[s] = rd * S.arr[1][o]; r.c2[s] = rd * S.arr[2][o]; r.c3[s] = rd * S.arr[3][o];
    }
}
template <int V> __device__ __forceinline__ void chain(const Ops& r, double* xg, bool edgeInY, bool edgeInZ, double& last)
{
    const unsigned full = 0xffffffffu;
#pragma unroll
    for (int s = 0; s < HS; s++) {
        double vy = __shfl_up_sync(full, last, 1), vz = __shfl_up_sync(full, last, WTJ);
        vy = edgeInY ? r.hy[s] : vy;
        vz = edgeInZ ? r.hz[s] : vz;
        const double m3 = r.c3[s] * vz, m2 = r.c2[s] * vy, m1 = r.c1[s] * last;
        double w = r.c0[s];
        w -= m3; w -= m2; w -= m1;
        if (V <= 1) xg[s * 32] = V == 0 ? publishable(w) : w;
        last = w;
    }
}
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity)
{
    asm volatile("{\n\t.reg .pred P1;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t@P1 bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}"
                 ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// SPIN: 0 = one warp only; 1 = two more warps parked in mbarrier.try_wait until the compute warp is done; 2 = two more warps
// polling global memory with ld.relaxed.gpu until then
template <int V, int SPIN> __global__ void k(double* out, long long* cyc, double seed, double* flag)
{
    extern __shared__ __align__(128) unsigned char raw[];
    __shared__ unsigned long long bar;
    Slot* slots = reinterpret_cast<Slot*>(raw);          // 2 slots
    if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory"); *flag = 0.0; }
    __syncthreads();
    if (threadIdx.x >= 32) {
        if (SPIN == 1) mbar_wait(&bar, 0);
        if (SPIN == 2) { double v = 0; do { asm volatile("ld.relaxed.gpu.global.f64 %0, [%1];" : "=d"(v) : "l"(flag) : "memory"); } while (v == 0.0); }
        return;
    }
    const int lane = threadIdx.x;
    for (int q = lane; q < 2 * (int)(sizeof(Slot) / 8); q += 32) reinterpret_cast<double*>(raw)[q] = seed * (1.0 + 1e-3 * (q % 97));
    __syncwarp();
    const bool edgeInY = (lane & 7) == 0, edgeInZ = lane < 8;
    double last = 0.0;
    Ops A, B;
    double* xg = out + lane;
    fetch(slots[0], 0, lane, A);
    const long long t0 = clock64();
    for (int g = 0; g < NG; g++) {
        const Slot& S = slots[g & 1];
        const Slot& N = slots[(g + 1) & 1];
        if (V == 3) {       // fetch, then chain, as separate phases (k_wave2 style)
            fetch(S, 0, lane, A); chain<0>(A, xg, edgeInY, edgeInZ, last);
            fetch(S, HS, lane, B); chain<0>(B, xg + HS * 32, edgeInY, edgeInZ, last);
        } else {
            fetch(S, HS, lane, B); chain<V>(A, xg, edgeInY, edgeInZ, last);
            fetch(N, 0, lane, A); chain<V>(B, xg + HS * 32, edgeInY, edgeInZ, last);
        }
        xg += GSZ * 32;
    }
    const long long t1 = clock64();
    out[lane] += last;
    if (lane == 0) { cyc[V + 4 * SPIN] = t1 - t0; asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar)) : "memory"); *flag = 1.0; }
}
int main()
{
    double *out, *flag; long long* cyc;
    cudaMalloc(&out, (size_t)(NG + 2) * GSZ * 32 * 8); cudaMalloc(&flag, 8); cudaMallocManaged(&cyc, 16 * 8);
    const int shm = 2 * sizeof(Slot);
#define RUN(V, SPIN) { cudaFuncSetAttribute(k<V, SPIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, shm); k<V, SPIN><<<1, SPIN ? 96 : 32, shm>>>(out, cyc, 1e-3, flag); cudaDeviceSynchronize(); }
    for (int r = 0; r < 2; r++) {
        RUN(0, 0) RUN(1, 0) RUN(2, 0) RUN(3, 0) RUN(0, 1) RUN(3, 1) RUN(0, 2) RUN(3, 2)
    }
    const char* n[4] = {"pipelined fetch + chain + publishable + STG", "pipelined, plain STG", "pipelined, no store", "fetch then chain (two phases)"};
    for (int i = 0; i < 4; i++) printf("%-48s %.1f cycles per step\n", n[i], cyc[i] / double(NG * GSZ));
    printf("with two warps parked in mbarrier.try_wait:   pipelined %.1f, two phases %.1f cycles per step\n", cyc[4] / double(NG * GSZ), cyc[7] / double(NG * GSZ));
    printf("with two warps polling ld.relaxed.gpu:        pipelined %.1f, two phases %.1f cycles per step\n", cyc[8] / double(NG * GSZ), cyc[11] / double(NG * GSZ));
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
