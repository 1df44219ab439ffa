// One context per process = one GPU = one subdomain (the reference's one MPI rank).
#pragma once
#include "common.cuh"
#include <nccl.h>
#include <vector>

namespace b200 {

#define B200_NCCL(expr)                                                                          \
    do {                                                                                         \
        ncclResult_t r_ = (expr);                                                                \
        if (r_ != ncclSuccess)                                                                   \
            throw ::b200::Error(std::string(#expr) + ": " + ncclGetErrorString(r_) + " at " +    \
                                __FILE__ + ":" + std::to_string(__LINE__));                      \
    } while (0)

// kernel groups timed by the built-in CUDA-event profiler (bench.py's roofline numbers)
enum ProfCat { PC_AMUL = 0, PC_SWEEP_FWD, PC_SWEEP_BWD, PC_FACTOR, PC_VECTOR, PC_NORMFACTOR, PC_CTRL, PC_PROPS, PC_DINUM, PC_FACES,
               PC_ASSEMBLE, PC_CORRECTOR, PC_ALPHA, PC_BOUNDARY, PC_QDOT, PC_COPY, PC_COMM, PC_FUNCOBJ, PC_NCAT };
const char* prof_name(int cat);

// Peer-memory mailbox of the small all-reduces (Krylov dot products, residual norms, time-step numbers): every rank stores
// its partial into every peer's mailbox over NVLink and sums the R partials in rank order itself -- one short kernel, no
// library call, bitwise the same result on every rank.
constexpr int PEER_MAX_RANKS = 16, PEER_RING = 4, PEER_VALS = 8;
struct PeerMail {
    double val[PEER_RING][PEER_MAX_RANKS][PEER_VALS];
    unsigned long long flag[PEER_RING][PEER_MAX_RANKS];
    // halo window (follows this struct in the same allocation): haloFlag[side][slot] = sequence number of the plane that the
    // neighbour on that side (0 = rank below, 1 = rank above) has finished storing into slot `slot` of my window
    unsigned long long haloFlag[2][2];
    unsigned long long pad_[12];          // the window starts on a 128-byte boundary
};
static_assert(sizeof(PeerMail) % 128 == 0, "halo window alignment");
struct PeerPtrs { PeerMail* p[PEER_MAX_RANKS]; };

struct Ctx {
    int device = 0, rank = 0, nranks = 1;
    PeerMail* myMail = nullptr;
    PeerPtrs peers{};
    bool peerOK = false;
    unsigned long long peerSeq = 0;
    // Halo planes over NVLink without NCCL: behind the mailbox, in the same IPC allocation, every rank keeps a window of
    // [2 sides][2 slots][haloDoubles]; a neighbour stores its boundary plane straight into it (k_halo_exchange, ctx.cu).
    size_t haloDoubles = 0;
    unsigned long long haloSeq = 0;
    unsigned* haloTicket = nullptr;
    double* haloWindow(PeerMail* base, int side, int slot) const
    {
        return reinterpret_cast<double*>(reinterpret_cast<char*>(base) + sizeof(PeerMail)) + ((size_t)side * 2 + slot) * haloDoubles;
    }
    bool peerHaloOK(size_t count) const { return peerOK && haloDoubles >= count && count > 0; }
    void peerExport(void* handle64, size_t haloDoublesWanted = 0);
    void peerImport(const void* handles);     // nranks x 64 bytes, rank order
    void allreduce(double* buf, int n, int op);   // op: 0 sum, 1 max, 2 min
    void allreduceCtrl(double* red, int slot0, int n, int phase, SolveState* st);   // sum + the control step that consumes it, one kernel
    // profiler: one event pair per instrumented launch group, resolved at read time
    bool profOn = false;
    struct ProfRec { int cat; cudaEvent_t a, b; int tag; };
    int profTag = -1;              // Krylov iteration being enqueued (Ldu::solve); records of iterations that turn out to be past
                                   // convergence -- kernels that return at once -- are dropped again (profDropFrom)
    void profDropFrom(size_t firstRec, int firstDeadTag);
    std::vector<ProfRec> profRecs;
    std::vector<cudaEvent_t> profPool;
    double profMs[PC_NCAT] = {0};
    int64_t profCount[PC_NCAT] = {0};
    cudaEvent_t marks[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t profEvent();
    void profRead(double* ms, int64_t* counts);
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

struct ProfScope {
    Ctx* c; cudaEvent_t b = nullptr;
    ProfScope(Ctx* ctx, int cat) : c(ctx)
    {
        if (!c->profOn) return;
        cudaEvent_t a = c->profEvent(); b = c->profEvent();
        c->profRecs.push_back({cat, a, b, c->profTag});
        cudaEventRecord(a, c->stream);
    }
    ~ProfScope() { if (b) cudaEventRecord(b, c->stream); }
};

// A few words device -> PINNED host (or host-by-value -> device) written by a one-warp kernel instead of a DMA copy: the copy
// engines serve transfers in order, so a tiny cudaMemcpyAsync of the step (solver state, reductions) would queue behind a
// field-sized transfer of the streamed host I/O (b200_case_set_field_async / get_field_async) for tens of milliseconds.
void words_to_host(Ctx* ctx, void* hostPinned, const void* dev, size_t bytes);      // bytes: multiple of 4, any size up to a few KB
void words_to_device(Ctx* ctx, void* dev, const void* host, size_t bytes);          // bytes <= 256 (passed as a kernel argument)

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
