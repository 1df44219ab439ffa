#include "ctx.cuh"
#include <cstring>

namespace b200 {

int64_t g_launches = 0;

Ctx::Ctx(int dev, const void* uniqueId, int rank_, int nranks_) : device(dev), rank(rank_), nranks(nranks_)
{
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        throw Error("no CUDA device available: the B200 path has no CPU fallback");
    B200_CUDA(cudaSetDevice(dev));
    cudaDeviceProp prop;
    B200_CUDA(cudaGetDeviceProperties(&prop, dev));
    smCount = prop.multiProcessorCount;
    B200_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    for (auto& e : marks) B200_CUDA(cudaEventCreate(&e));
    if (nranks > 1) {
        B200_CHECK(uniqueId != nullptr, "nccl unique id required for nranks > 1");
        ncclUniqueId id;
        std::memcpy(&id, uniqueId, sizeof(id));
        B200_NCCL(ncclCommInitRank(&comm, nranks, id, rank));
    }
}

const char* prof_name(int cat)
{
    static const char* names[PC_NCAT] = {"amul", "sweep_fwd", "sweep_bwd", "factor", "vector", "normfactor", "ctrl", "props", "dinum",
                                         "faces", "assemble", "corrector", "alpha", "boundary", "qdot", "copy", "comm"};
    return cat >= 0 && cat < PC_NCAT ? names[cat] : "?";
}

cudaEvent_t Ctx::profEvent()
{
    if (!profPool.empty()) { cudaEvent_t e = profPool.back(); profPool.pop_back(); return e; }
    cudaEvent_t e;
    B200_CUDA(cudaEventCreate(&e));
    return e;
}

void Ctx::profRead(double* ms, int64_t* counts)
{
    sync();
    for (auto& r : profRecs) {
        float t = 0;
        if (cudaEventElapsedTime(&t, r.a, r.b) == cudaSuccess) { profMs[r.cat] += t; profCount[r.cat]++; }
        profPool.push_back(r.a); profPool.push_back(r.b);
    }
    profRecs.clear();
    for (int i = 0; i < PC_NCAT; i++) { if (ms) ms[i] = profMs[i]; if (counts) counts[i] = profCount[i]; profMs[i] = 0; profCount[i] = 0; }
}

Ctx::~Ctx()
{
    for (auto& r : profRecs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    for (auto e : profPool) cudaEventDestroy(e);
    for (auto e : marks) if (e) cudaEventDestroy(e);
    if (comm) ncclCommDestroy(comm);
    if (stream) cudaStreamDestroy(stream);
}

void Ctx::allreduceSum(double* buf, int n)
{
    if (nranks > 1) B200_NCCL(ncclAllReduce(buf, buf, n, ncclDouble, ncclSum, comm, stream));
}
void Ctx::allreduceMax(double* buf, int n)
{
    if (nranks > 1) B200_NCCL(ncclAllReduce(buf, buf, n, ncclDouble, ncclMax, comm, stream));
}
void Ctx::allreduceMin(double* buf, int n)
{
    if (nranks > 1) B200_NCCL(ncclAllReduce(buf, buf, n, ncclDouble, ncclMin, comm, stream));
}

void Ctx::sendrecv(const double* sendLo, double* recvLo, int rankLo, const double* sendHi, double* recvHi, int rankHi,
                   size_t count)
{
    if (nranks <= 1) return;
    B200_NCCL(ncclGroupStart());
    if (rankLo >= 0) {
        B200_NCCL(ncclSend(sendLo, count, ncclDouble, rankLo, comm, stream));
        B200_NCCL(ncclRecv(recvLo, count, ncclDouble, rankLo, comm, stream));
    }
    if (rankHi >= 0) {
        B200_NCCL(ncclSend(sendHi, count, ncclDouble, rankHi, comm, stream));
        B200_NCCL(ncclRecv(recvHi, count, ncclDouble, rankHi, comm, stream));
    }
    B200_NCCL(ncclGroupEnd());
}

} // namespace b200
