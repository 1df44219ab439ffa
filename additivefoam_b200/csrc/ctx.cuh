// One context per process = one GPU = one subdomain (the reference's one MPI rank).
#pragma once
#include "common.cuh"
#include <nccl.h>

namespace b200 {

#define B200_NCCL(expr)                                                                          \
    do {                                                                                         \
        ncclResult_t r_ = (expr);                                                                \
        if (r_ != ncclSuccess)                                                                   \
            throw ::b200::Error(std::string(#expr) + ": " + ncclGetErrorString(r_) + " at " +    \
                                __FILE__ + ":" + std::to_string(__LINE__));                      \
    } while (0)

struct Ctx {
    int device = 0, rank = 0, nranks = 1;
    int smCount = 148;
    int maxCoopBlocksPerSm = 8;
    cudaStream_t stream = nullptr;
    ncclComm_t comm = nullptr;

    Ctx(int dev, const void* uniqueId, int rank_, int nranks_);
    ~Ctx();
    void sync() { B200_CUDA(cudaStreamSynchronize(stream)); }
    // persistent grid: one wave of 256-thread blocks
    int persistentBlocks(int blocksPerSm = 8) const
    {
        int b = smCount * blocksPerSm;
        return b > RED_MAX_BLOCKS ? RED_MAX_BLOCKS : b;
    }
    void allreduceSum(double* buf, int n);
    void allreduceMax(double* buf, int n);
    void allreduceMin(double* buf, int n);
    // plane / buffer exchange with a neighbour rank (either may be null)
    void sendrecv(const double* sendLo, double* recvLo, int rankLo, const double* sendHi, double* recvHi, int rankHi,
                  size_t count);
};

template <class T> inline T* dev_alloc(size_t n)
{
    T* p = nullptr;
    B200_CUDA(cudaMalloc(&p, (n ? n : 1) * sizeof(T)));
    return p;
}
template <class T> inline void dev_free(T*& p)
{
    if (p) cudaFree(p);
    p = nullptr;
}
template <class T> inline void h2d(T* dst, const T* src, size_t n, cudaStream_t s)
{
    B200_CUDA(cudaMemcpyAsync(dst, src, n * sizeof(T), cudaMemcpyHostToDevice, s));
}
template <class T> inline void d2h(T* dst, const T* src, size_t n, cudaStream_t s)
{
    B200_CUDA(cudaMemcpyAsync(dst, src, n * sizeof(T), cudaMemcpyDeviceToHost, s));
}

} // namespace b200
