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
                                         "faces", "assemble", "corrector", "alpha", "boundary", "qdot", "copy", "comm", "funcobj"};
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
    for (int r = 0; r < nranks && peerOK; r++) if (r != rank && peers.p[r]) cudaIpcCloseMemHandle(peers.p[r]);
    if (myMail) cudaFree(myMail);
    if (comm) ncclCommDestroy(comm);
    if (stream) cudaStreamDestroy(stream);
}

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// One warp: lane r pushes this rank's partials into rank r's mailbox (a peer store over NVLink, or a local store for
// r == rank), then waits for rank r's partials in its own mailbox; lanes < n then sum the R partials in rank order.
__global__ void k_peer_allreduce(PeerMail* mine, PeerPtrs peers, int rank, int nranks, unsigned long long seq, double* vals, int n, int op)
{
    const int r = threadIdx.x;
    const int slot = (int)(seq % PEER_RING);
    if (r < nranks) {
        PeerMail* dst = peers.p[r];
        for (int i = 0; i < n; i++) dst->val[slot][rank][i] = vals[i];
        __threadfence_system();
        st_release_sys(&dst->flag[slot][rank], seq);
        while (ld_acquire_sys(&mine->flag[slot][r]) != seq) {}
    }
    __syncwarp();
    if (r < n) {
        double acc = mine->val[slot][0][r];
        for (int q = 1; q < nranks; q++) {
            const double x = mine->val[slot][q][r];
            acc = op == 0 ? acc + x : op == 1 ? (acc > x ? acc : x) : (acc < x ? acc : x);
        }
        vals[r] = acc;
    }
}

void Ctx::peerExport(void* handle64)
{
    if (!myMail) {
        B200_CUDA(cudaMalloc(&myMail, sizeof(PeerMail)));
        B200_CUDA(cudaMemset(myMail, 0, sizeof(PeerMail)));
    }
    cudaIpcMemHandle_t h;
    B200_CUDA(cudaIpcGetMemHandle(&h, myMail));
    static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
    std::memcpy(handle64, &h, sizeof(h));
}

void Ctx::peerImport(const void* handles)
{
    B200_CHECK(myMail != nullptr, "peerExport first");
    B200_CHECK(nranks <= PEER_MAX_RANKS, "too many ranks for the peer mailbox");
    for (int r = 0; r < nranks; r++) {
        if (r == rank) { peers.p[r] = myMail; continue; }
        cudaIpcMemHandle_t h;
        std::memcpy(&h, (const char*)handles + 64 * (size_t)r, sizeof(h));
        void* p = nullptr;
        B200_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        peers.p[r] = (PeerMail*)p;
    }
    peerOK = true;
}

void Ctx::allreduce(double* buf, int n, int op)
{
    if (nranks <= 1) return;
    if (peerOK) {
        B200_CHECK(n <= PEER_VALS, "peer all-reduce handles up to 8 values");
        k_peer_allreduce<<<1, 32, 0, stream>>>(myMail, peers, rank, nranks, ++peerSeq, buf, n, op);
        B200_LAUNCH_CHECK();
        return;
    }
    const ncclRedOp_t o = op == 0 ? ncclSum : op == 1 ? ncclMax : ncclMin;
    B200_NCCL(ncclAllReduce(buf, buf, n, ncclDouble, o, comm, stream));
}
void Ctx::allreduceSum(double* buf, int n) { allreduce(buf, n, 0); }
void Ctx::allreduceMax(double* buf, int n) { allreduce(buf, n, 1); }
void Ctx::allreduceMin(double* buf, int n) { allreduce(buf, n, 2); }

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
