#include "ctx.cuh"
#include <cstring>
#include <algorithm>

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

void Ctx::profDropFrom(size_t firstRec, int firstDeadTag)
{
    size_t w = firstRec;
    for (size_t r = firstRec; r < profRecs.size(); r++) {
        if (profRecs[r].tag >= firstDeadTag) { profPool.push_back(profRecs[r].a); profPool.push_back(profRecs[r].b); }
        else profRecs[w++] = profRecs[r];
    }
    profRecs.resize(w);
}

struct WordBlob { unsigned w[64]; };
__global__ void k_words_to_host(unsigned* __restrict__ host, const unsigned* __restrict__ dev, int n)
{
    for (int i = threadIdx.x; i < n; i += 32) host[i] = dev[i];
    __threadfence_system();
}
__global__ void k_words_to_device(unsigned* __restrict__ dev, WordBlob b, int n)
{
    for (int i = threadIdx.x; i < n; i += 32) dev[i] = b.w[i];
}
void words_to_host(Ctx* ctx, void* hostPinned, const void* dev, size_t bytes)
{
    B200_CHECK(bytes % 4 == 0, "words_to_host: size");
    k_words_to_host<<<1, 32, 0, ctx->stream>>>((unsigned*)hostPinned, (const unsigned*)dev, (int)(bytes / 4));
    B200_CUDA(cudaGetLastError());
}
void words_to_device(Ctx* ctx, void* dev, const void* host, size_t bytes)
{
    B200_CHECK(bytes % 4 == 0 && bytes <= sizeof(WordBlob), "words_to_device: size");
    WordBlob b;
    std::memcpy(b.w, host, bytes);
    k_words_to_device<<<1, 32, 0, ctx->stream>>>((unsigned*)dev, b, (int)(bytes / 4));
    B200_CUDA(cudaGetLastError());
}

Ctx::~Ctx()
{
    for (auto& r : profRecs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    for (auto e : profPool) cudaEventDestroy(e);
    for (auto e : marks) if (e) cudaEventDestroy(e);
    for (int r = 0; r < nranks && peerOK; r++) if (r != rank && peers.p[r]) cudaIpcCloseMemHandle(peers.p[r]);
    if (myMail) cudaFree(myMail);
    if (haloTicket) cudaFree(haloTicket);
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
// `phase` != C_NONE: lane 0 then runs the Krylov control step that consumes the sums (solve_state.cuh) -- `red` is the base of
// the reduction slots, `vals` = red + slot0.
__global__ void k_peer_allreduce(PeerMail* mine, PeerPtrs peers, int rank, int nranks, unsigned long long seq, double* vals, int n, int op,
                                 const double* red, int phase, SolveState* st)
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
    __syncwarp();
    if (r == 0 && phase != C_NONE) ctrl_apply(phase, st, red);
}

void Ctx::peerExport(void* handle64, size_t haloDoublesWanted)
{
    if (!myMail) {
        haloDoubles = (haloDoublesWanted + 15) / 16 * 16;
        const size_t bytes = sizeof(PeerMail) + 4 * haloDoubles * sizeof(double);
        B200_CUDA(cudaMalloc(&myMail, bytes));
        B200_CUDA(cudaMemset(myMail, 0, bytes));
        B200_CUDA(cudaMalloc(&haloTicket, sizeof(unsigned)));
        B200_CUDA(cudaMemset(haloTicket, 0, sizeof(unsigned)));
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
        k_peer_allreduce<<<1, 32, 0, stream>>>(myMail, peers, rank, nranks, ++peerSeq, buf, n, op, nullptr, C_NONE, nullptr);
        B200_LAUNCH_CHECK();
        return;
    }
    const ncclRedOp_t o = op == 0 ? ncclSum : op == 1 ? ncclMax : ncclMin;
    B200_NCCL(ncclAllReduce(buf, buf, n, ncclDouble, o, comm, stream));
}
void Ctx::allreduceCtrl(double* red, int slot0, int n, int phase, SolveState* st)
{
    B200_CHECK(peerOK && n <= PEER_VALS, "allreduceCtrl needs the peer mailbox");
    k_peer_allreduce<<<1, 32, 0, stream>>>(myMail, peers, rank, nranks, ++peerSeq, red + slot0, n, 0, red, phase, st);
    B200_LAUNCH_CHECK();
}
void Ctx::allreduceSum(double* buf, int n) { allreduce(buf, n, 0); }
void Ctx::allreduceMax(double* buf, int n) { allreduce(buf, n, 1); }
void Ctx::allreduceMin(double* buf, int n) { allreduce(buf, n, 2); }

// Plane exchange with the slab neighbours over peer memory, one kernel, no library call:
//   1. every block stores its share of my bottom / top plane into the neighbour's window (peer stores over NVLink),
//   2. the last block to finish (ticket) publishes the sequence number in the neighbour's haloFlag (release, system scope),
//   3. every block waits for the neighbours' sequence number in MY flags and copies their planes from my window into my ghost
//      planes (L1-bypassing loads: the lines were written by the peer).
// No rank waits before it has pushed, so two ranks running this kernel cannot wait for each other in a cycle, whatever part of
// the grid is resident.  Two window slots alternate: a neighbour's push number s+2 follows its wait for my push s+1, which follows
// my pull of s in stream order.
__global__ void __launch_bounds__(256) k_halo_exchange(const double* __restrict__ sendLo, double* __restrict__ recvLo, double* peerLoWin,
                                                       unsigned long long* peerLoFlag, const double* __restrict__ sendHi,
                                                       double* __restrict__ recvHi, double* peerHiWin, unsigned long long* peerHiFlag,
                                                       const double* myWinFromLo, const unsigned long long* myFlagFromLo,
                                                       const double* myWinFromHi, const unsigned long long* myFlagFromHi, size_t count,
                                                       unsigned long long seq, unsigned* ticket)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x, i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (peerLoWin) for (size_t i = i0; i < count; i += stride) peerLoWin[i] = sendLo[i];
    if (peerHiWin) for (size_t i = i0; i < count; i += stride) peerHiWin[i] = sendHi[i];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        if (atomicAdd(ticket, 1u) == gridDim.x - 1) {
            __threadfence_system();
            if (peerLoFlag) st_release_sys(peerLoFlag, seq);
            if (peerHiFlag) st_release_sys(peerHiFlag, seq);
            *ticket = 0u;
        }
        if (myFlagFromLo) while (ld_acquire_sys(myFlagFromLo) != seq) {}
        if (myFlagFromHi) while (ld_acquire_sys(myFlagFromHi) != seq) {}
    }
    __syncthreads();
    if (myWinFromLo) for (size_t i = i0; i < count; i += stride) recvLo[i] = __ldcg(myWinFromLo + i);
    if (myWinFromHi) for (size_t i = i0; i < count; i += stride) recvHi[i] = __ldcg(myWinFromHi + i);
}

void Ctx::sendrecv(const double* sendLo, double* recvLo, int rankLo, const double* sendHi, double* recvHi, int rankHi,
                   size_t count)
{
    if (nranks <= 1) return;
    if (peerHaloOK(count)) {
        const unsigned long long seq = ++haloSeq;
        const int slot = (int)(seq & 1);
        // my bottom plane goes to the rank below, where I am the neighbour ABOVE (side 1); my top plane to side 0 of the rank above
        double* loWin = rankLo >= 0 ? haloWindow(peers.p[rankLo], 1, slot) : nullptr;
        unsigned long long* loFlag = rankLo >= 0 ? &peers.p[rankLo]->haloFlag[1][slot] : nullptr;
        double* hiWin = rankHi >= 0 ? haloWindow(peers.p[rankHi], 0, slot) : nullptr;
        unsigned long long* hiFlag = rankHi >= 0 ? &peers.p[rankHi]->haloFlag[0][slot] : nullptr;
        const int grid = (int)std::max<size_t>(1, std::min<size_t>((count + 1023) / 1024, (size_t)smCount));
        k_halo_exchange<<<grid, 256, 0, stream>>>(sendLo, recvLo, loWin, loFlag, sendHi, recvHi, hiWin, hiFlag,
                                                  rankLo >= 0 ? haloWindow(myMail, 0, slot) : nullptr, rankLo >= 0 ? &myMail->haloFlag[0][slot] : nullptr,
                                                  rankHi >= 0 ? haloWindow(myMail, 1, slot) : nullptr, rankHi >= 0 ? &myMail->haloFlag[1][slot] : nullptr,
                                                  count, seq, haloTicket);
        B200_LAUNCH_CHECK();
        return;
    }
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
