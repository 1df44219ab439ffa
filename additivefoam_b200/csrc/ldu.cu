#include "ldu_kernels.cuh"
#include "wave_kernels.cuh"
#include "wave4_kernels.cuh"
#include "flow_kernels.cuh"
#include <cstdlib>
#include <set>
#include "ctx.cuh"
#include <algorithm>
#include <cstring>

namespace b200 {

// ------------------------------------------------------------------ cooperative launch helper
template <class K> static int coop_max_blocks(Ctx* ctx, K kernel)
{
    int perSm = 0;
    B200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, kernel, BLK, 0));
    B200_CHECK(perSm > 0, "sweep kernel does not fit on an SM");
    return perSm * ctx->smCount;
}
template <class K, class... Args> static void coop_launch(Ctx* ctx, K kernel, int64_t maxWork, Args... args)
{
    int grid = coop_max_blocks(ctx, kernel);
    const int64_t want = (maxWork + BLK - 1) / BLK;
    if (want < grid) grid = (int)std::max<int64_t>(want, 1);
    void* ptrs[] = {(void*)&args...};
    B200_CUDA(cudaLaunchCooperativeKernel((const void*)kernel, dim3(grid), dim3(BLK), ptrs, 0, ctx->stream));
    count_launch();
}

// ------------------------------------------------------------------ construction
Ldu::Ldu(Ctx* c, int nCells_, int nFaces_, const int* lower, const int* upper, int nInterfaces, const int* ifaceNeighbRank,
         const int* ifaceSize, const int* const* ifaceFaceCells)
    : ctx(c), nCells(nCells_), nFaces(nFaces_)
{
    B200_CHECK(nCells >= 0 && nFaces >= 0, "negative sizes");
    hasGeneral = true;
    h_lower.assign(lower, lower + nFaces);
    h_upper.assign(upper, upper + nFaces);
    for (int f = 0; f < nFaces; f++) {
        B200_CHECK(lower[f] >= 0 && upper[f] < nCells && lower[f] < upper[f], "addressing is not upper-triangular");
        if (f > 0) B200_CHECK(lower[f - 1] <= lower[f], "lowerAddr is not sorted (upper-triangular order required)");
    }
    // row table: faces where the cell is the upper address (ascending), then owned faces (ascending)
    std::vector<int> rowStart(nCells + 1, 0), rowMid(nCells, 0);
    std::vector<int> cntL(nCells, 0), cntU(nCells, 0);
    for (int f = 0; f < nFaces; f++) { cntL[upper[f]]++; cntU[lower[f]]++; }
    for (int cI = 0; cI < nCells; cI++) { rowStart[cI + 1] = rowStart[cI] + cntL[cI] + cntU[cI]; rowMid[cI] = rowStart[cI] + cntL[cI]; }
    std::vector<int> rowCol(2 * (size_t)nFaces), rowFace(2 * (size_t)nFaces);
    {
        std::vector<int> posL(rowStart.begin(), rowStart.end() - 1), posU(rowMid);
        for (int f = 0; f < nFaces; f++) {
            int e = posL[upper[f]]++; rowCol[e] = lower[f]; rowFace[e] = f;
            e = posU[lower[f]]++; rowCol[e] = upper[f]; rowFace[e] = f;
        }
    }
    // level schedule of the lower-neighbour DAG and the level-renumbered cell list
    std::vector<int> level(nCells, 0);
    nLevels = nCells ? 1 : 0;
    for (int f = 0; f < nFaces; f++) {
        const int l = level[lower[f]] + 1;
        if (l > level[upper[f]]) level[upper[f]] = l;
        nLevels = std::max(nLevels, l + 1);
    }
    h_lvlStart.assign(nLevels + 1, 0);
    for (int cI = 0; cI < nCells; cI++) h_lvlStart[level[cI] + 1]++;
    for (int L = 0; L < nLevels; L++) { maxLevelSize = std::max(maxLevelSize, h_lvlStart[L + 1]); h_lvlStart[L + 1] += h_lvlStart[L]; }
    std::vector<int> lvlCells(nCells);
    {
        std::vector<int> pos(h_lvlStart.begin(), h_lvlStart.end() - 1);
        for (int cI = 0; cI < nCells; cI++) lvlCells[pos[level[cI]]++] = cI;
    }
    h_level = level;
    {
        const char* e = std::getenv("B200_SWEEP");          // "coop": one cooperative launch with a grid barrier per level
        useFlow = !(e && std::string(e) == "coop");
    }
    cudaStream_t s = ctx->stream;
    if (useFlow && nCells > 0) {
        // dependency tables of the barrier-free sweeps, in position (= level) order with neighbour POSITIONS:
        // lower-side faces ascending (forward / factor), owned faces descending (backward)
        std::vector<int> posOf(nCells);
        for (int p = 0; p < nCells; p++) posOf[lvlCells[p]] = p;
        std::vector<int> fStart(nCells + 1, 0), bStart(nCells + 1, 0), fPos(nFaces), fFace(nFaces), bPos(nFaces), bFace(nFaces);
        for (int p = 0; p < nCells; p++) {
            const int cI = lvlCells[p];
            fStart[p + 1] = fStart[p] + (rowMid[cI] - rowStart[cI]);
            bStart[p + 1] = bStart[p] + (rowStart[cI + 1] - rowMid[cI]);
            int o = fStart[p];
            for (int e = rowStart[cI]; e < rowMid[cI]; e++, o++) { fPos[o] = posOf[rowCol[e]]; fFace[o] = rowFace[e]; }
            o = bStart[p];
            for (int e = rowStart[cI + 1] - 1; e >= rowMid[cI]; e--, o++) { bPos[o] = posOf[rowCol[e]]; bFace[o] = rowFace[e]; }
        }
        d_fStart = dev_alloc<int>(nCells + 1); d_bStart = dev_alloc<int>(nCells + 1);
        d_fPos = dev_alloc<int>(std::max(nFaces, 1)); d_fFace = dev_alloc<int>(std::max(nFaces, 1));
        d_bPos = dev_alloc<int>(std::max(nFaces, 1)); d_bFace = dev_alloc<int>(std::max(nFaces, 1));
        h2d(d_fStart, fStart.data(), nCells + 1, s); h2d(d_bStart, bStart.data(), nCells + 1, s);
        h2d(d_fPos, fPos.data(), nFaces, s); h2d(d_fFace, fFace.data(), nFaces, s);
        h2d(d_bPos, bPos.data(), nFaces, s); h2d(d_bFace, bFace.data(), nFaces, s);
        ctx->sync();                                         // the staging vectors go out of scope here
    }
    d_lower = dev_alloc<int>(nFaces); d_upper = dev_alloc<int>(nFaces);
    d_rowStart = dev_alloc<int>(nCells + 1); d_rowMid = dev_alloc<int>(nCells);
    d_rowCol = dev_alloc<int>(2 * (size_t)nFaces); d_rowFace = dev_alloc<int>(2 * (size_t)nFaces);
    d_lvlCells = dev_alloc<int>(nCells); d_lvlStart = dev_alloc<int>(nLevels + 1);
    h2d(d_lower, lower, nFaces, s); h2d(d_upper, upper, nFaces, s);
    h2d(d_rowStart, rowStart.data(), nCells + 1, s); h2d(d_rowMid, rowMid.data(), nCells, s);
    h2d(d_rowCol, rowCol.data(), rowCol.size(), s); h2d(d_rowFace, rowFace.data(), rowFace.size(), s);
    h2d(d_lvlCells, lvlCells.data(), nCells, s); h2d(d_lvlStart, h_lvlStart.data(), nLevels + 1, s);
    B200_CHECK(nInterfaces == 0 || ctx->nranks > 1, "coupled interfaces need a multi-rank context (b200_ctx_create with nranks > 1)");
    for (int p = 0; p < nInterfaces; p++) {
        DevInterface it;
        B200_CHECK(ifaceNeighbRank[p] >= 0 && ifaceNeighbRank[p] < ctx->nranks && ifaceNeighbRank[p] != ctx->rank, "interface neighbour rank out of range");
        it.size = ifaceSize[p]; it.neighbRank = ifaceNeighbRank[p];
        it.faceCells = dev_alloc<int>(it.size);
        h2d(it.faceCells, ifaceFaceCells[p], it.size, s);
        it.recvBuf = dev_alloc<double>(it.size); it.sendBuf = dev_alloc<double>(it.size);
        std::vector<int> sorted(ifaceFaceCells[p], ifaceFaceCells[p] + it.size);
        std::sort(sorted.begin(), sorted.end());
        it.hasDuplicates = std::adjacent_find(sorted.begin(), sorted.end()) != sorted.end();
        ifaces.push_back(it);
    }
    ctx->sync();
}

Ldu::Ldu(Ctx* c, BlockDims d) : ctx(c), dims(d)
{
    B200_CHECK(d.nCells() < (int64_t)INT32_MAX && d.nFaces() < (int64_t)INT32_MAX, "block too large for int32 labels");
    nCells = (int)d.nCells(); nFaces = (int)d.nFaces();
    isBlock = true; hasGeneral = false;
    nLevels = d.nx + d.ny + d.nz - 2;
    const char* e = std::getenv("B200_SWEEP");
    useWave = !(e && std::string(e) == "coop");
}

Ldu::~Ldu()
{
    dev_free(d_lower); dev_free(d_upper); dev_free(d_rowStart); dev_free(d_rowMid); dev_free(d_rowCol); dev_free(d_rowFace);
    dev_free(d_lvlCells); dev_free(d_lvlStart);
    for (auto& it : ifaces) { dev_free(it.faceCells); dev_free(it.recvBuf); dev_free(it.sendBuf); }
    for (auto& w : work) dev_free(w);
    dev_free(rD); dev_free(d_sumA);
    dev_free(d_state);
    if (h_state) cudaFreeHost(h_state);
    dev_free(red.partials); dev_free(red.ticket); dev_free(red.result);
    for (auto& p : stage) dev_free(p);
    for (auto& p : stageBou) dev_free(p);
    for (auto& p : wave) dev_free(p);
    dev_free(waveFlags); dev_free(wavePartial); dev_free(waveOrder); dev_free(waveTrace); dev_free(waveTraceG); dev_free(waveOrderC[0]); dev_free(waveOrderC[1]); dev_free(waveOrderC[2]);
    dev_free(d_fStart); dev_free(d_fPos); dev_free(d_fFace); dev_free(d_bStart); dev_free(d_bPos); dev_free(d_bFace);
    dev_free(flowCoefF); dev_free(flowCoefB); dev_free(flowNumF); dev_free(flowRD); dev_free(flowF); dev_free(flowB);
    dev_free(flowPartial); dev_free(flowTicket);
}

void Ldu::setBlock(int nx, int ny, int nz)
{
    BlockDims d; d.nx = nx; d.ny = ny; d.nz = nz;
    B200_CHECK(nx > 0 && ny > 0 && nz > 0, "block dimensions must be positive");
    B200_CHECK(d.nCells() == nCells && d.nFaces() == nFaces, "block dimensions do not match nCells/nFaces");
    // Coupled interfaces (one rank of a decomposed OpenFOAM case whose sub-domain is a block: `simple` / `hierarchical` decomposition
    // of a blockMesh block): the block kernels do the interior -- Amul without ghost planes, the block-local DIC/DILU sweeps of
    // [UPSTREAM] processor boundaries -- and the interface contributions are applied after the gather, as on general addressing.
    // expand the closed form on the device and compare bit-exactly with the addressing given at creation
    int* lo = dev_alloc<int>(nFaces); int* up = dev_alloc<int>(nFaces);
    k_block_addressing<<<std::min(ny * nz, 4096), BLK, 0, ctx->stream>>>(d, lo, up);
    B200_LAUNCH_CHECK();
    std::vector<int> hl(nFaces), hu(nFaces);
    d2h(hl.data(), lo, nFaces, ctx->stream); d2h(hu.data(), up, nFaces, ctx->stream);
    ctx->sync();
    dev_free(lo); dev_free(up);
    B200_CHECK(hl == h_lower && hu == h_upper, "addressing is not a single blockMesh hex block of these dimensions");
    dims = d; isBlock = true;
    const char* e = std::getenv("B200_SWEEP");
    useWave = !(e && std::string(e) == "coop");
    for (auto& w : work) dev_free(w);       // ghost padding changes
    dev_free(rD);
    // the wave-order arrays, the tile order and the gathered coefficients belong to the block shape
    for (auto& p : wave) dev_free(p);
    dev_free(waveFlags); dev_free(wavePartial); dev_free(waveOrder); dev_free(waveTrace); dev_free(waveTraceG); dev_free(waveOrderC[0]); dev_free(waveOrderC[1]); dev_free(waveOrderC[2]);
    waveEpoch = -1; waveUpper = nullptr; waveLower = nullptr;
}

void Ldu::levels(int* levelOfCell, int* nLev)
{
    if (hasGeneral) {
        std::memcpy(levelOfCell, h_level.data(), sizeof(int) * nCells);
    } else {
        for (int k = 0; k < dims.nz; k++) for (int j = 0; j < dims.ny; j++) for (int i = 0; i < dims.nx; i++)
            levelOfCell[i + dims.nx * (j + dims.ny * k)] = i + j + k;
    }
    *nLev = nLevels;
}

void Ldu::ensureWork()
{
    if (!d_state) {
        d_state = dev_alloc<SolveState>(1);
        B200_CUDA(cudaMallocHost(&h_state, sizeof(SolveState)));
        red.partials = dev_alloc<double>((size_t)S_NSLOTS * RED_MAX_BLOCKS);
        red.ticket = dev_alloc<unsigned>(1);
        red.result = dev_alloc<double>(S_NSLOTS);
        B200_CUDA(cudaMemsetAsync(red.ticket, 0, sizeof(unsigned), ctx->stream));
        B200_CUDA(cudaMemsetAsync(red.result, 0, sizeof(double) * S_NSLOTS, ctx->stream));
    }
    if (!work[0]) {
        ghost = isBlock ? (int64_t)dims.nx * dims.ny : 0;
        const size_t len = (size_t)nCells + 2 * (size_t)ghost;
        for (auto& w : work) { w = dev_alloc<double>(len); B200_CUDA(cudaMemsetAsync(w, 0, len * sizeof(double), ctx->stream)); }
        rD = dev_alloc<double>(len);
        B200_CUDA(cudaMemsetAsync(rD, 0, len * sizeof(double), ctx->stream));
    }
    redBlocks = std::min(ctx->persistentBlocks(), red_blocks_for(std::max(nCells, 1)));
}

// ------------------------------------------------------------------ halo / interface update
void Ldu::exchange(const double* x)
{
    if (ctx->nranks <= 1) return;
    ProfScope ps_(ctx, PC_COMM);
    if (isBlock && (ghostBelow || ghostAbove)) {
        const size_t plane = (size_t)dims.nx * dims.ny;
        double* xw = const_cast<double*>(x);
        ctx->sendrecv(x, xw - plane, ghostBelow ? ctx->rank - 1 : -1, x + (size_t)(dims.nz - 1) * plane, xw + (size_t)dims.nz * plane,
                      ghostAbove ? ctx->rank + 1 : -1, plane);
        return;
    }
    if (ifaces.empty()) return;
    for (auto& it : ifaces) {
        k_iface_gather<<<std::max(1, std::min((it.size + BLK - 1) / BLK, 1024)), BLK, 0, ctx->stream>>>(it.size, it.faceCells, x, it.sendBuf);
        B200_LAUNCH_CHECK();
    }
    B200_NCCL(ncclGroupStart());
    for (auto& it : ifaces) {
        B200_NCCL(ncclSend(it.sendBuf, it.size, ncclDouble, it.neighbRank, ctx->comm, ctx->stream));
        B200_NCCL(ncclRecv(it.recvBuf, it.size, ncclDouble, it.neighbRank, ctx->comm, ctx->stream));
    }
    B200_NCCL(ncclGroupEnd());
}

// After a kernel left a local sum in red.result[slot0 .. slot0+n): make it global and run the Krylov control step `phase` that
// consumes it.  One rank: the producing kernel's last block already ran the step (redFor(phase)) -- nothing to launch.  Several
// ranks: the peer all-reduce kernel runs it (one launch for both); over NCCL a k_ctrl launch follows the library call.
void Ldu::reduceCtrl(int slot0, int n, int phase)
{
    if (ctx->nranks <= 1) return;
    if (ctx->peerOK) {
        ProfScope ps_(ctx, PC_COMM);
        ctx->allreduceCtrl(red.result, slot0, n, phase, d_state);
        return;
    }
    { ProfScope ps_(ctx, PC_COMM); ctx->allreduceSum(red.result + slot0, n); }
    if (phase != C_NONE) {
        ProfScope ps_(ctx, PC_CTRL);
        k_ctrl<<<1, 1, 0, ctx->stream>>>(phase, d_state, red.result);
        B200_LAUNCH_CHECK();
    }
}

RedWork Ldu::redFor(int phase) const
{
    RedWork r = red;
    if (ctx->nranks <= 1) { r.phase = phase; r.state = d_state; }
    return r;
}

// ------------------------------------------------------------------ Amul
template <int MODE> static void launch_amul(Ldu& L, const MatrixView& A, const double* x, double* y, const double* aux, const SolveState* st, const RedWork& red)
{
    cudaStream_t s = L.ctx->stream;
    ProfScope ps_(L.ctx, PC_AMUL);
    if (L.isBlock) {
        const double* bb = (L.ghostBelow && A.bouCoeffs) ? A.bouCoeffs[0] : nullptr;
        const double* ba = (L.ghostAbove && A.bouCoeffs) ? A.bouCoeffs[L.ghostBelow ? 1 : 0] : nullptr;
        const int64_t items = (int64_t)((L.dims.nx + BLK - 1) / BLK) * L.dims.ny * L.dims.nz;
        const int grid = (int)std::min<int64_t>(items, L.ctx->persistentBlocks());
        k_amul_block<MODE><<<grid, BLK, 0, s>>>(L.dims, A.diag, A.upper, A.lower, x, y, aux, bb, ba, red, st);
    } else {
        k_amul_general<MODE><<<L.redBlocks, BLK, 0, s>>>(L.nCells, L.d_rowStart, L.d_rowMid, L.d_rowCol, L.d_rowFace, A.diag, A.upper,
                                                         A.lower, x, y, aux, red, st);
    }
    B200_LAUNCH_CHECK();
}

// y = A x with the sum `mode` fused in; `phase` = the control step that consumes that sum (the caller follows up with
// reduceCtrl(slot, n, phase), which is a no-op on one rank)
void Ldu::amulFused(const MatrixView& A, const double* x, double* y, int mode, const double* aux, int phase)
{
    const SolveState* st = d_state;
    exchange(x);
    const bool post = !ifaces.empty();       // general coupled interfaces: apply after the gather, reduce separately
    const int m = post ? AM_NONE : mode;
    const RedWork rw = redFor(phase);
    switch (m) {
    case AM_NONE: launch_amul<AM_NONE>(*this, A, x, y, aux, st, rw); break;
    case AM_SUMX: launch_amul<AM_SUMX>(*this, A, x, y, aux, st, rw); break;
    case AM_DOT_YX: launch_amul<AM_DOT_YX>(*this, A, x, y, aux, st, rw); break;
    case AM_DOT_AUXY: launch_amul<AM_DOT_AUXY>(*this, A, x, y, aux, st, rw); break;
    case AM_TT_TS: launch_amul<AM_TT_TS>(*this, A, x, y, aux, st, rw); break;
    }
    if (!post) return;
    cudaStream_t s = ctx->stream;
    for (size_t p = 0; p < ifaces.size(); p++) {
        auto& it = ifaces[p];
        const int grid = it.hasDuplicates ? 1 : std::max(1, std::min((it.size + BLK - 1) / BLK, 1024));
        k_iface_apply<<<grid, BLK, 0, s>>>(it.size, it.faceCells, A.bouCoeffs[p], it.recvBuf, y, it.hasDuplicates ? 1 : 0, st);
        B200_LAUNCH_CHECK();
    }
    const int64_t n = nCells;      // (coupled interfaces mean several ranks: the control step runs in reduceCtrl)
    if (mode == AM_SUMX) { k_sum<<<redBlocks, BLK, 0, s>>>(n, x, red, S_SUMPSI); B200_LAUNCH_CHECK(); }
    else if (mode == AM_DOT_YX) { k_dot<<<redBlocks, BLK, 0, s>>>(n, y, x, red, S_DEN, st); B200_LAUNCH_CHECK(); }
    else if (mode == AM_DOT_AUXY) { k_dot<<<redBlocks, BLK, 0, s>>>(n, aux, y, red, S_DEN, st); B200_LAUNCH_CHECK(); }
    else if (mode == AM_TT_TS) {
        { ProfScope ps_(ctx, PC_VECTOR); k_dot<<<redBlocks, BLK, 0, s>>>(n, y, y, red, S_TT, st); B200_LAUNCH_CHECK(); }
        { ProfScope ps_(ctx, PC_VECTOR); k_dot<<<redBlocks, BLK, 0, s>>>(n, y, aux, red, S_TS, st); B200_LAUNCH_CHECK(); }
    }
}

void Ldu::amul(const MatrixView& A, const double* psi, double* Apsi)
{
    ensureWork();
    B200_CUDA(cudaMemsetAsync(d_state, 0, sizeof(SolveState), ctx->stream));
    amulFused(A, psi, Apsi, AM_NONE, nullptr, C_NONE);
}

// ------------------------------------------------------------------ tile-wavefront sweeps (block addressing)
static WaveGeom wave_geom(const BlockDims& d)
{
    WaveGeom g;
    g.nx = d.nx; g.ny = d.ny; g.nz = d.nz;
    g.nJ = (d.ny + WTJ - 1) / WTJ; g.nK = (d.nz + WTK - 1) / WTK; g.nTiles = g.nJ * g.nK;
    g.nSteps = d.nx + WTJ + WTK - 2;
    g.nStepsPad = (g.nSteps + 2 * GS - 1) / (2 * GS) * (2 * GS);      // an even number of GS-step (and WS-step) groups
    g.nxny = (int64_t)d.nx * d.ny;
    return g;
}

void Ldu::ensureWave(bool dilu)
{
    const WaveGeom g = wave_geom(dims);
    waveLen = (int64_t)g.nTiles * g.nStepsPad * 32;
    for (int q = 0; q < 14; q++)
        if (!wave[q] && (q < 10 || q == 13 || dilu)) { wave[q] = dev_alloc<double>((size_t)waveLen); B200_CUDA(cudaMemsetAsync(wave[q], 0, (size_t)waveLen * sizeof(double), ctx->stream)); }
    if (!waveFlags) {
        waveFlags = dev_alloc<int>(4); wavePartial = dev_alloc<double>(g.nTiles);
        B200_CUDA(cudaMemsetAsync(waveFlags, 0, 4 * sizeof(int), ctx->stream));     // [0] tile ticket, [1] redo exactly, [2] veto, [3] block count
        // Tiles are handed out along the wavefront, not row by row.  Tile (J, K) can run only a fixed number of steps behind
        // (J-1, K) and (J, K-1) (the lane skew of the tile edge plus the look-ahead of the loads, ~15 and ~11 steps, plus the
        // store-to-poll latency), so its earliest start is about 9J + 7K in those units.  Handing tiles out in that order keeps
        // the resident ones within a fraction of a tile length of each other; in row-major order a row of 50 tiles alone spans
        // more than the ~1000 steps of a tile, and most resident warps would sit waiting for their upstream neighbour.  Any
        // order increasing in both J and K is topological, which is all the deadlock argument needs.
        std::vector<int> order(g.nTiles);
        for (int t = 0; t < g.nTiles; t++) order[t] = t;
        const int nJ = g.nJ;
        static const int wJ = [] { const char* e = std::getenv("B200_WAVE_ORDER_J"); return e ? std::atoi(e) : 9; }();
        static const int wK = [] { const char* e = std::getenv("B200_WAVE_ORDER_K"); return e ? std::atoi(e) : 7; }();
        std::stable_sort(order.begin(), order.end(), [&](int x, int y) {
            const long kx = (long)wJ * (x % nJ) + (long)wK * (x / nJ), ky = (long)wJ * (y % nJ) + (long)wK * (y / nJ);
            return kx < ky;
        });
        waveOrder = dev_alloc<int>(g.nTiles);
        B200_CUDA(cudaMemcpyAsync(waveOrder, order.data(), sizeof(int) * g.nTiles, cudaMemcpyHostToDevice, ctx->stream));
        // k_wave3: CTA tiles of NW tiles in J.  A J-hop between CTAs costs about NW tile lags, a K-hop one: hand out by 3*Jc + K
        for (int q = 0; q < 3; q++) {
            const int NW = q == 0 ? 4 : q == 1 ? 2 : 1;
            const int nJc = (g.nJ + NW - 1) / NW, nC = nJc * g.nK;
            std::vector<int> oc(nC);
            for (int t = 0; t < nC; t++) oc[t] = t;
            // relative start times of the CTA tiles: a J-hop costs about NW tile lags of WTJ-1 steps each, a K-hop WTK-1 steps
            const long cJ = NW == 1 ? 9 : 3 * NW, cK = NW == 1 ? 7 : 4;
            std::stable_sort(oc.begin(), oc.end(), [&](int x, int y) {
                return cJ * (x % nJc) + cK * (x / nJc) < cJ * (y % nJc) + cK * (y / nJc);
            });
            waveOrderC[q] = dev_alloc<int>(nC);
            B200_CUDA(cudaMemcpyAsync(waveOrderC[q], oc.data(), sizeof(int) * nC, cudaMemcpyHostToDevice, ctx->stream));
            B200_CUDA(cudaStreamSynchronize(ctx->stream));
        }
        B200_CUDA(cudaStreamSynchronize(ctx->stream));        // `order` is a local
    }
}

// persistent grid of the transposing vector kernels (wave_kernels.cuh): exactly what is resident
int Ldu::waveVecGrid()
{
    static int perSM = 0;
    if (!perSM) {
        int a = 0, b = 0, c = 0;
        B200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, k_wave_pcg_update, 256, 0));
        B200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, k_wave_pcg_p, 256, 0));
        B200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c, k_wave_from_natural, 256, 0));
        perSM = std::max(1, std::min(a, std::min(b, c)));
    }
    const WaveGeom g = wave_geom(dims);
    const int64_t items = (int64_t)g.nTiles * ((g.nStepsPad + TCH - 1) / TCH);
    return (int)std::min<int64_t>(std::max<int64_t>(1, std::min<int64_t>(items, (int64_t)ctx->smCount * perSM)), RED_MAX_BLOCKS);
}

// op: 0 factor (after the gather), 1 forward, 2 backward
void Ldu::waveSweep(int op, bool dilu, bool dot, const double* rA, double* wA, bool wio, int phase)
{
    WaveArgs a;
    a.g = wave_geom(dims);
    a.A0 = wave[0]; a.PX = wave[1]; a.PY = wave[2]; a.PZ = wave[3]; a.QX = wave[4]; a.QY = wave[5]; a.QZ = wave[6];
    a.WF = wave[7]; a.RF = wave[8]; a.WB = wave[9]; a.NX = wave[10]; a.NY = wave[11]; a.NZ = wave[12]; a.DG = wave[13];
    a.rA = rA; a.wA = wA;
    a.ticket = waveFlags; a.order = waveOrder; a.tilePartial = wavePartial; a.st = d_state;
    a.trace = nullptr; a.traceG = nullptr; a.nCtas = 0; a.orderC = nullptr;
    { static const int hd = [] { const char* e = std::getenv("B200_WAVE4_DEPTH"); return e ? std::atoi(e) : 1; }(); a.haloDepth = hd; }
    static const bool traceOn = [] { const char* e = std::getenv("B200_WAVE_TRACE"); return e && std::atoi(e) != 0; }();
    if (traceOn) {
        if (!waveTrace) { waveTrace = dev_alloc<long long>((size_t)a.g.nTiles * WTRACE); B200_CUDA(cudaMemsetAsync(waveTrace, 0, (size_t)a.g.nTiles * WTRACE * sizeof(long long), ctx->stream)); }
        a.trace = waveTrace;
        static const bool traceAll = [] { const char* e = std::getenv("B200_WAVE_TRACE"); return e && std::atoi(e) >= 2; }();
        if (traceAll) {
            const size_t nG = (size_t)a.g.nTiles * (a.g.nStepsPad / 4);
            if (!waveTraceG) waveTraceG = dev_alloc<long long>(nG);
            B200_CUDA(cudaMemsetAsync(waveTraceG, 0, nG * sizeof(long long), ctx->stream));
            a.traceG = waveTraceG;
        }
    }
    {
        // keep what the resident tiles have prefetched but not yet consumed well inside L2 (126 MB): about 32 MB in flight
        const int64_t resident = std::min<int64_t>(a.g.nTiles, (int64_t)ctx->smCount * 12);
        const int64_t perGroup = resident * 5 * WS * 32 * (int64_t)sizeof(double);
        int pf = (int)((32ll << 20) / std::max<int64_t>(perGroup, 1));
        pf = std::min(12, pf / 4 * 4);
        static const char* e = std::getenv("B200_WAVE_PF");
        a.pfGroups = e ? std::atoi(e) : pf;
    }
    cudaStream_t s = ctx->stream;
    B200_CUDA(cudaMemsetAsync(waveFlags, 0, (op == 0 ? 2 : 1) * sizeof(int), s));      // tile ticket (+ the factor sweep's out-of-range report)
    // the factor sweep exchanges raw rD through RF: arm it with the sentinel (all-ones NaN); the factor sweep itself arms
    // WF and WB for the substitution sweeps, which then re-arm each other's array as they go
    if (op == 0) B200_CUDA(cudaMemsetAsync(wave[8], 0xFF, (size_t)waveLen * sizeof(double), s));
    const int grid = a.g.nTiles;
    // Two implementations of the same sweep (bit-identical results): the warp-specialised kernel (memory warp + compute warp
    // per tile) is the default at every size now that tiles are handed out along the wavefront; the one-warp kernel stays
    // selectable (B200_WAVE=1warp) as the reference point the profiles compare against.
    // B200_WAVE: "1warp" the one-warp kernel, "single" k_wave2 (one tile per CTA), "multi" k_wave3 (NW tiles per CTA); default:
    // k_wave3 where the dependency chain across tiles bounds the sweep (few tiles: thin slabs, small blocks), k_wave2 where
    // bandwidth does (thousands of tiles)
    const int forced = [] { const char* e = std::getenv("B200_WAVE"); return !e ? 0 : std::string(e) == "1warp" ? 1 : std::string(e) == "multi" ? 3 : std::string(e) == "tri" ? 4 : 2; }();
    const bool oneWarp = forced == 1;
    // default: where the dependency chain across tiles bounds the sweep (few tiles: thin slabs, small blocks) the three-role kernel
    // k_wave4, whose tiles follow their producers at a third less lag; where bandwidth does k_wave2.  Measured inside a PCG/DIC solve
    // (tools/perf_mid.sh, profiles/r2_mid_slab_sweeps.txt): 400 tiles (1000 x 400 x 31) 0.30 / 0.32 ms against 0.40 / 0.43; 800 tiles
    // (1000 x 400 x 63) 0.40 / 0.49 against 0.41 / 0.42 and the factor sweep 1.22 against 1.09 ms; 1600 and more: k_wave2.
    const bool tri = forced == 4 || (forced == 0 && a.g.nTiles <= 600);
    const bool multi = forced == 3;
    if (op == 0 && !oneWarp) {      // the warp-specialised factor sweep streams the diagonal from wave order as well
        k_wave_from_natural<<<waveVecGrid(), 256, 0, s>>>(a.g, rA, wave[13], 1.0, d_state); B200_LAUNCH_CHECK();
    }
    if (oneWarp) {
        if (op == 0) { if (dilu) k_wave<0, true, false, false><<<grid, 32, 0, s>>>(a); else k_wave<0, false, false, false><<<grid, 32, 0, s>>>(a); }
        else if (op == 1) { if (wio) k_wave<1, false, false, true><<<grid, 32, 0, s>>>(a); else k_wave<1, false, false, false><<<grid, 32, 0, s>>>(a); }
        else if (wio) { if (dot) k_wave<2, false, true, true><<<grid, 32, 0, s>>>(a); else k_wave<2, false, false, true><<<grid, 32, 0, s>>>(a); }
        else if (dot) k_wave<2, false, true, false><<<grid, 32, 0, s>>>(a);
        else k_wave<2, false, false, false><<<grid, 32, 0, s>>>(a);
    } else if (tri) {
        // k_wave4 (wave4_kernels.cuh): compute / halo / stream warps per tile; ring geometry as for k_wave2
        auto launch4 = [&](auto kernel, size_t shm) {
            static std::set<const void*> configured;
            if (configured.insert((const void*)kernel).second)
                B200_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm));
            kernel<<<grid, 96, shm, s>>>(a);
        };
        const char* eg = std::getenv("B200_WAVE2_RING");
        const int cfg = eg ? std::atoi(eg) : (a.g.nTiles > 1536 ? 83 : 86);
#define B200_W4(G_, R_)                                                                                                                        \
        do {                                                                                                                                   \
            if (op == 0) { if (dilu) launch4(k_wave4<0, true, false, true, true, G_, R_>, sizeof(WaveSmem4<0, true, true, G_, R_>));                     \
                           else launch4(k_wave4<0, false, false, true, true, G_, R_>, sizeof(WaveSmem4<0, false, true, G_, R_>)); }                      \
            else if (op == 1) { if (wio) launch4(k_wave4<1, false, false, true, false, G_, R_>, sizeof(WaveSmem4<1, false, true, G_, R_>));               \
                                else launch4(k_wave4<1, false, false, false, false, G_, R_>, sizeof(WaveSmem4<1, false, false, G_, R_>)); }               \
            else if (wio) { if (dot) launch4(k_wave4<2, false, true, true, false, G_, R_>, sizeof(WaveSmem4<2, false, true, G_, R_>));                    \
                            else launch4(k_wave4<2, false, false, true, false, G_, R_>, sizeof(WaveSmem4<2, false, true, G_, R_>)); }                     \
            else if (dot) launch4(k_wave4<2, false, true, false, false, G_, R_>, sizeof(WaveSmem4<2, false, false, G_, R_>));                             \
            else launch4(k_wave4<2, false, false, false, false, G_, R_>, sizeof(WaveSmem4<2, false, false, G_, R_>));                                     \
            if (op == 0) {                                                                                                                     \
                k_wave_refactor_prepare<<<ctx->smCount * 4, 256, 0, s>>>(a, waveLen);                                                          \
                if (dilu) launch4(k_wave4<0, true, false, true, false, G_, R_>, sizeof(WaveSmem4<0, true, true, G_, R_>));                                \
                else launch4(k_wave4<0, false, false, true, false, G_, R_>, sizeof(WaveSmem4<0, false, true, G_, R_>));                                   \
            }                                                                                                                                  \
        } while (0)
        if (cfg == 83) B200_W4(8, 3);
        else if (cfg == 84) B200_W4(8, 4);
        else if (cfg == 48) B200_W4(4, 8);
        else if (cfg == 412) B200_W4(4, 12);
        else B200_W4(8, 6);
#undef B200_W4
    } else if (multi) {
        // geometry by B200_WAVE3 = <NW><GSZ><RNG> (tiles per CTA, steps per ring slot, ring slots); default 4 x 4 x 6
        const char* eg = std::getenv("B200_WAVE3");
        const int cfg = eg ? std::atoi(eg) : 446;
        const int NWsel = cfg / 100 == 2 ? 2 : cfg / 100 == 1 ? 1 : 4;
        a.nCtas = ((a.g.nJ + NWsel - 1) / NWsel) * a.g.nK;
        a.orderC = waveOrderC[NWsel == 4 ? 0 : NWsel == 2 ? 1 : 2];
        auto launch3 = [&](auto kernel, size_t shm, int threads) {
            static std::set<const void*> configured;
            if (configured.insert((const void*)kernel).second)
                B200_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm));
            kernel<<<a.nCtas, threads, shm, s>>>(a);
        };
#define B200_W3(G_, R_, RD_, N_)                                                                                                                \
        do {                                                                                                                                   \
            if (op == 0) { if (dilu) launch3(k_wave3<0, true, false, true, true, G_, RD_, N_>, sizeof(WaveSmem3<0, true, true, G_, RD_, N_>), 64 * N_);   \
                           else launch3(k_wave3<0, false, false, true, true, G_, R_, N_>, sizeof(WaveSmem3<0, false, true, G_, R_, N_>), 64 * N_); }      \
            else if (op == 1) { if (wio) launch3(k_wave3<1, false, false, true, false, G_, R_, N_>, sizeof(WaveSmem3<1, false, true, G_, R_, N_>), 64 * N_); \
                                else launch3(k_wave3<1, false, false, false, false, G_, R_, N_>, sizeof(WaveSmem3<1, false, false, G_, R_, N_>), 64 * N_); } \
            else if (wio) { if (dot) launch3(k_wave3<2, false, true, true, false, G_, R_, N_>, sizeof(WaveSmem3<2, false, true, G_, R_, N_>), 64 * N_);   \
                            else launch3(k_wave3<2, false, false, true, false, G_, R_, N_>, sizeof(WaveSmem3<2, false, true, G_, R_, N_>), 64 * N_); }    \
            else if (dot) launch3(k_wave3<2, false, true, false, false, G_, R_, N_>, sizeof(WaveSmem3<2, false, false, G_, R_, N_>), 64 * N_);            \
            else launch3(k_wave3<2, false, false, false, false, G_, R_, N_>, sizeof(WaveSmem3<2, false, false, G_, R_, N_>), 64 * N_);                    \
            if (op == 0) {                                                                                                                     \
                k_wave_refactor_prepare<<<ctx->smCount * 4, 256, 0, s>>>(a, waveLen);                                                          \
                if (dilu) launch3(k_wave3<0, true, false, true, false, G_, RD_, N_>, sizeof(WaveSmem3<0, true, true, G_, RD_, N_>), 64 * N_);             \
                else launch3(k_wave3<0, false, false, true, false, G_, R_, N_>, sizeof(WaveSmem3<0, false, true, G_, R_, N_>), 64 * N_);                  \
            }                                                                                                                                  \
        } while (0)
        if (cfg == 444) B200_W3(4, 4, 4, 4);
        else if (cfg == 248) B200_W3(4, 8, 8, 2);
        else if (cfg == 183) B200_W3(8, 3, 3, 1);
        else if (cfg == 146) B200_W3(4, 6, 6, 1);
        else B200_W3(4, 6, 5, 4);
#undef B200_W3
    } else {
        auto launch = [&](auto kernel, size_t shm) {
            static std::set<const void*> configured;     // every k_wave2 instantiation has the same pointer type
            if (configured.insert((const void*)kernel).second)
                B200_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm));
            kernel<<<grid, 64, shm, s>>>(a);
        };
        // ring geometry: few tiles => deep slots (latency of the chain); many tiles => small slots, more CTAs per SM
        const char* eg = std::getenv("B200_WAVE2_RING");
        // many tiles: 3 slots -> 4 CTAs per SM keep HBM busy; few tiles (thin slab, latency-bound): 6 slots of look-ahead
        const int cfg = eg ? std::atoi(eg) : (a.g.nTiles > 1536 ? 83 : 86);
#define B200_W2(G_, R_)                                                                                                                        \
        do {                                                                                                                                   \
            if (op == 0) { if (dilu) launch(k_wave2<0, true, false, true, true, G_, R_>, sizeof(WaveSmem<0, true, true, G_, R_>));                      \
                           else launch(k_wave2<0, false, false, true, true, G_, R_>, sizeof(WaveSmem<0, false, true, G_, R_>)); }                       \
            else if (op == 1) { if (wio) launch(k_wave2<1, false, false, true, false, G_, R_>, sizeof(WaveSmem<1, false, true, G_, R_>));                \
                                else launch(k_wave2<1, false, false, false, false, G_, R_>, sizeof(WaveSmem<1, false, false, G_, R_>)); }                \
            else if (wio) { if (dot) launch(k_wave2<2, false, true, true, false, G_, R_>, sizeof(WaveSmem<2, false, true, G_, R_>));                     \
                            else launch(k_wave2<2, false, false, true, false, G_, R_>, sizeof(WaveSmem<2, false, true, G_, R_>)); }                      \
            else if (dot) launch(k_wave2<2, false, true, false, false, G_, R_>, sizeof(WaveSmem<2, false, false, G_, R_>));                              \
            else launch(k_wave2<2, false, false, false, false, G_, R_>, sizeof(WaveSmem<2, false, false, G_, R_>));                                      \
        } while (0)
        if (cfg == 44) B200_W2(4, 4);
        else if (cfg == 83) B200_W2(8, 3);
        else if (cfg == 82) B200_W2(8, 2);
        else if (cfg == 86) B200_W2(8, 6);
        else if (cfg == 412) B200_W2(4, 12);
        else if (cfg == 48) B200_W2(4, 8);
        else B200_W2(8, 4);
#undef B200_W2
        if (op == 0) {
            // the exact factor sweep (plain FP64 operators) behind the fast one: both kernels return at once unless the fast
            // sweep was vetoed by the coefficients or reported a diagonal entry outside its exponent range
            k_wave_refactor_prepare<<<ctx->smCount * 4, 256, 0, s>>>(a, waveLen);
            if (dilu) launch(k_wave2<0, true, false, true, false, 8, 4>, sizeof(WaveSmem<0, true, true, 8, 4>));
            else launch(k_wave2<0, false, false, true, false, 8, 4>, sizeof(WaveSmem<0, false, true, 8, 4>));
        }
    }
    B200_LAUNCH_CHECK();
    if (traceOn) {          // diagnostic: per-tile time stamps of this sweep, relative to the earliest, on stderr
        std::vector<long long> h((size_t)a.g.nTiles * WTRACE);
        d2h(h.data(), waveTrace, h.size(), s);
        ctx->sync();
        long long t0 = h[11];
        for (int t = 0; t < a.g.nTiles; t++) t0 = std::min(t0, h[WTRACE * t + 11]);
        if (a.traceG) {         // every group of every tile of the first K-row and of the middle column: where the lag between tiles grows
            const int nG = a.g.nStepsPad / 8;
            std::vector<long long> hg((size_t)a.g.nTiles * (a.g.nStepsPad / 4));
            d2h(hg.data(), waveTraceG, hg.size(), s);
            ctx->sync();
            std::fprintf(stderr, "WAVE_GROUPS op %d nJ %d nK %d nG %d\n", op, a.g.nJ, a.g.nK, nG);
            for (int t = 0; t < a.g.nTiles; t++) {
                if (t / a.g.nJ != 0 && t % a.g.nJ != a.g.nJ / 2 && t / a.g.nJ != a.g.nK - 1) continue;
                std::fprintf(stderr, "WG %d %d", t % a.g.nJ, t / a.g.nJ);
                for (int q = 0; q < nG; q++) std::fprintf(stderr, " %lld", hg[(size_t)t * nG + q] ? hg[(size_t)t * nG + q] - t0 : -1);
                std::fprintf(stderr, "\n");
            }
        }
        std::fprintf(stderr, "WAVE_TRACE op %d nJ %d nK %d\n", op, a.g.nJ, a.g.nK);
        for (int t = 0; t < a.g.nTiles; t++) {
            std::fprintf(stderr, "WT %d %d", t % a.g.nJ, t / a.g.nJ);
            for (int q = 0; q < 12; q++) std::fprintf(stderr, " %lld", h[WTRACE * t + q] - t0);
            for (int q = 12; q < WTRACE; q++) std::fprintf(stderr, " %lld", h[WTRACE * t + q]);
            std::fprintf(stderr, "\n");
        }
    }
    if (op == 2 && dot) { k_wave_sum<<<1, 256, 0, s>>>(a.g.nTiles, wavePartial, red.result, (int)S_RHO, d_state, ctx->nranks <= 1 ? phase : (int)C_NONE); B200_LAUNCH_CHECK(); }
}

// ------------------------------------------------------------------ preconditioners
void Ldu::factor(int kind, const MatrixView& A)
{
    cudaStream_t s = ctx->stream;
    const SolveState* st = d_state;
    if (kind == 0) return;
    ProfScope ps_(ctx, PC_FACTOR);
    if (kind == 1) {
        k_copy<<<redBlocks, BLK, 0, s>>>((int64_t)nCells, A.diag, rD + ghost, st); B200_LAUNCH_CHECK();
    } else {
        const bool dilu = kind == 3;
        double* rd = rD + ghost;
        const double* nul = nullptr; double* nulw = nullptr;
        if (isBlock && useWave) {
            ensureWave(dilu);
            // the unscaled wave-order coefficients depend on the off-diagonals only: regather when they may have changed
            const bool same = A.offDiagEpoch >= 0 && A.offDiagEpoch == waveEpoch && A.upper == waveUpper && A.lower == waveLower
                           && dilu == waveDilu;
            if (!same) {
                WaveArgs a;
                a.g = wave_geom(dims);
                a.A0 = wave[0]; a.PX = wave[1]; a.PY = wave[2]; a.PZ = wave[3]; a.QX = wave[4]; a.QY = wave[5]; a.QZ = wave[6];
                a.WF = wave[7]; a.RF = wave[8]; a.WB = wave[9]; a.NX = wave[10]; a.NY = wave[11]; a.NZ = wave[12]; a.st = st;
                a.ticket = waveFlags;
                B200_CUDA(cudaMemsetAsync(waveFlags + 2, 0, sizeof(int), s));      // the veto of the fast factor sweep belongs to the off-diagonals
                const int64_t rows = (int64_t)a.g.nTiles * a.g.nStepsPad;
                int perSM = 8;
                if (dilu) B200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, k_wave_gather<true>, 256, 0));
                else B200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, k_wave_gather<false>, 256, 0));
                const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((rows + 7) / 8, (int64_t)ctx->smCount * std::max(perSM, 1)));
                if (dilu) k_wave_gather<true><<<grid, 256, 0, s>>>(dims, a, A.upper, A.lower);
                else k_wave_gather<false><<<grid, 256, 0, s>>>(dims, a, A.upper, A.lower);
                B200_LAUNCH_CHECK();
                waveEpoch = A.offDiagEpoch; waveUpper = A.upper; waveLower = A.lower; waveDilu = dilu;
            }
            waveSweep(0, dilu, false, A.diag, nullptr, false, C_NONE);
            return;          // rD lives in wave order (A0)
        } else if (isBlock) {
            if (dilu) coop_launch(ctx, k_sweep_block<0, true, false>, (int64_t)dims.ny * dims.nz, dims, A.diag, A.upper, A.lower, rd, nul, nulw, nul, red, 0, st);
            else coop_launch(ctx, k_sweep_block<0, false, false>, (int64_t)dims.ny * dims.nz, dims, A.diag, A.upper, A.lower, rd, nul, nulw, nul, red, 0, st);
        } else if (useFlow) {
            // coefficients into the position-ordered tables, then the raw recurrence into a sentinel-filled rD; the sweeps'
            // exchange arrays are reset here too, once per solve
            const int grid = std::max(1, std::min(redBlocks, (nCells + BLK - 1) / BLK));
            if (!flowF) {
                const size_t nn = (size_t)std::max(nCells, 1), nf = (size_t)std::max(nFaces, 1);
                flowF = dev_alloc<double>(nn); flowB = dev_alloc<double>(nn); flowRD = dev_alloc<double>(nn);
                flowCoefF = dev_alloc<double>(nf); flowCoefB = dev_alloc<double>(nf); flowNumF = dev_alloc<double>(nf);
            }
            if (nFaces > 0) {
                const int gg = std::max(1, std::min(redBlocks, (nFaces + BLK - 1) / BLK));
                if (dilu) k_flow_gather<true><<<gg, BLK, 0, s>>>(nFaces, d_fFace, d_bFace, A.upper, A.lower, flowCoefF, flowCoefB, flowNumF, st);
                else k_flow_gather<false><<<gg, BLK, 0, s>>>(nFaces, d_fFace, d_bFace, A.upper, A.lower, flowCoefF, flowCoefB, flowNumF, st);
                B200_LAUNCH_CHECK();
            }
            k_fill_sentinel<<<grid, BLK, 0, s>>>((int64_t)nCells, flowRD, st); B200_LAUNCH_CHECK();
            k_fill_sentinel<<<grid, BLK, 0, s>>>((int64_t)nCells, flowF, st); B200_LAUNCH_CHECK();
            k_fill_sentinel<<<grid, BLK, 0, s>>>((int64_t)nCells, flowB, st); B200_LAUNCH_CHECK();
            flowSweep(0, dilu, false, A, nullptr, nullptr, nullptr, C_NONE);
            k_reciprocal<<<redBlocks, BLK, 0, s>>>((int64_t)nCells, flowRD, st); B200_LAUNCH_CHECK();
            return;              // rD lives in position order (flowRD)
        } else {
            const int* ls = d_lvlStart;
            if (dilu) coop_launch(ctx, k_sweep_general<0, true, false>, (int64_t)maxLevelSize, nLevels, ls, (const int*)d_lvlCells, (const int*)d_rowStart, (const int*)d_rowMid, (const int*)d_rowCol, (const int*)d_rowFace, A.diag, A.upper, A.lower, rd, nul, nulw, nul, red, 0, st);
            else coop_launch(ctx, k_sweep_general<0, false, false>, (int64_t)maxLevelSize, nLevels, ls, (const int*)d_lvlCells, (const int*)d_rowStart, (const int*)d_rowMid, (const int*)d_rowCol, (const int*)d_rowFace, A.diag, A.upper, A.lower, rd, nul, nulw, nul, red, 0, st);
        }
    }
    k_reciprocal<<<redBlocks, BLK, 0, s>>>((int64_t)nCells, rD + ghost, st); B200_LAUNCH_CHECK();
}

// what: 0 factor recurrence (rD sentinel-filled by the caller), 1 forward, 2 backward (+ dot into S_RHO)
void Ldu::flowSweep(int what, bool dilu, bool dot, const MatrixView& A, const double* rA, double* wA, const double* dotVec, int phase)
{
    cudaStream_t s = ctx->stream;
    const int nChunks = (nCells + 31) / 32;              // one per warp ticket
    if (nChunks == 0) return;
    if (!flowTicket) { flowTicket = dev_alloc<int>(1); flowPartial = dev_alloc<double>((size_t)nChunks); }
    FlowArgs a;
    (void)dilu;                    // the coefficient tables were gathered for DIC or DILU by factor()
    a.n = nCells; a.perm = d_lvlCells; a.fStart = d_fStart; a.fPos = d_fPos; a.bStart = d_bStart; a.bPos = d_bPos;
    a.coefF = flowCoefF; a.coefB = flowCoefB; a.numF = flowNumF;
    a.diag = A.diag; a.rD = flowRD; a.rA = rA; a.F = flowF; a.B = flowB; a.wA = wA; a.dotVec = dotVec;
    a.chunkPartial = flowPartial; a.ticket = flowTicket; a.st = d_state;
    static int perSM = 0;
    if (!perSM) {
        B200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, k_sweep_flow<2, true>, BLK, 0));
        perSM = std::max(perSM, 1);
    }
    static const int envBlocks = [] { const char* e = std::getenv("B200_FLOW_BLOCKS"); return e ? std::atoi(e) : 0; }();
    static const int envSleep = [] { const char* e = std::getenv("B200_FLOW_SLEEP"); return e ? std::atoi(e) : 0; }();
    a.sleepNs = envSleep;
    // few warps in flight: the ones far ahead of the front only poll, and their polls are L2 traffic the front has to queue behind
    const int grid = std::min((nChunks + BLK / 32 - 1) / (BLK / 32), ctx->smCount * std::min(envBlocks > 0 ? envBlocks : 2, perSM));
    B200_CUDA(cudaMemsetAsync(flowTicket, 0, sizeof(int), s));
    if (what == 0) k_sweep_flow<0, false><<<grid, BLK, 0, s>>>(a);
    else if (what == 1) k_sweep_flow<1, false><<<grid, BLK, 0, s>>>(a);
    else if (dot) k_sweep_flow<2, true><<<grid, BLK, 0, s>>>(a);
    else k_sweep_flow<2, false><<<grid, BLK, 0, s>>>(a);
    B200_LAUNCH_CHECK();
    if (what == 2 && dot) { k_sum<<<red_blocks_for(nChunks), BLK, 0, s>>>((int64_t)nChunks, flowPartial, redFor(phase), (int)S_RHO, d_state); B200_LAUNCH_CHECK(); }
}

// wA = M^-1 rA, optionally with sum(wA . dotVec) into slot S_RHO
void Ldu::applyPrecond(int kind, const MatrixView& A, const double* rA, double* wA, int doDot, const double* dotVec, bool wio, int phase)
{
    cudaStream_t s = ctx->stream;
    const SolveState* st = d_state;
    double* rd = rD + ghost;
    const RedWork rw = redFor(doDot ? phase : (int)C_NONE);
    if (kind == 0) { ProfScope ps_(ctx, PC_VECTOR); k_precond_pointwise<false><<<redBlocks, BLK, 0, s>>>((int64_t)nCells, rd, rA, wA, dotVec, rw, S_RHO, doDot, st); B200_LAUNCH_CHECK(); return; }
    if (kind == 1) { ProfScope ps_(ctx, PC_VECTOR); k_precond_pointwise<true><<<redBlocks, BLK, 0, s>>>((int64_t)nCells, rd, rA, wA, dotVec, rw, S_RHO, doDot, st); B200_LAUNCH_CHECK(); return; }
    const bool dilu = kind == 3;
    const double* nul = nullptr;
    if (isBlock && useWave) {
        { ProfScope ps_(ctx, PC_SWEEP_FWD); waveSweep(1, dilu, false, rA, wA, wio, C_NONE); }
        { ProfScope ps_(ctx, PC_SWEEP_BWD); waveSweep(2, dilu, doDot != 0, rA, wA, wio, phase); }
        B200_CHECK(!doDot || dotVec == rA, "fused preconditioner dot product is defined against rA");
    } else if (isBlock) {
        const int64_t work_ = (int64_t)dims.ny * dims.nz;
        if (dilu) { ProfScope ps_(ctx, PC_SWEEP_FWD); coop_launch(ctx, k_sweep_block<1, true, false>, work_, dims, A.diag, A.upper, A.lower, rd, rA, wA, nul, red, 0, st); }
        else { ProfScope ps_(ctx, PC_SWEEP_FWD); coop_launch(ctx, k_sweep_block<1, false, false>, work_, dims, A.diag, A.upper, A.lower, rd, rA, wA, nul, red, 0, st); }
        if (doDot) { ProfScope ps_(ctx, PC_SWEEP_BWD); coop_launch(ctx, k_sweep_block<2, false, true>, work_, dims, A.diag, A.upper, A.lower, rd, rA, wA, dotVec, rw, (int)S_RHO, st); }
        else { ProfScope ps_(ctx, PC_SWEEP_BWD); coop_launch(ctx, k_sweep_block<2, false, false>, work_, dims, A.diag, A.upper, A.lower, rd, rA, wA, nul, red, 0, st); }
    } else if (useFlow) {
        { ProfScope ps_(ctx, PC_SWEEP_FWD); flowSweep(1, dilu, false, A, rA, wA, nullptr, C_NONE); }
        { ProfScope ps_(ctx, PC_SWEEP_BWD); flowSweep(2, dilu, doDot != 0, A, rA, wA, dotVec, phase); }
    } else {
        const int* ls = d_lvlStart; const int* lc = d_lvlCells; const int* rs = d_rowStart; const int* rm = d_rowMid;
        const int* rc = d_rowCol; const int* rf = d_rowFace;
        const int64_t work_ = maxLevelSize;
        if (dilu) { ProfScope ps_(ctx, PC_SWEEP_FWD); coop_launch(ctx, k_sweep_general<1, true, false>, work_, nLevels, ls, lc, rs, rm, rc, rf, A.diag, A.upper, A.lower, rd, rA, wA, nul, red, 0, st); }
        else { ProfScope ps_(ctx, PC_SWEEP_FWD); coop_launch(ctx, k_sweep_general<1, false, false>, work_, nLevels, ls, lc, rs, rm, rc, rf, A.diag, A.upper, A.lower, rd, rA, wA, nul, red, 0, st); }
        if (doDot) { ProfScope ps_(ctx, PC_SWEEP_BWD); coop_launch(ctx, k_sweep_general<2, false, true>, work_, nLevels, ls, lc, rs, rm, rc, rf, A.diag, A.upper, A.lower, rd, rA, wA, dotVec, rw, (int)S_RHO, st); }
        else { ProfScope ps_(ctx, PC_SWEEP_BWD); coop_launch(ctx, k_sweep_general<2, false, false>, work_, nLevels, ls, lc, rs, rm, rc, rf, A.diag, A.upper, A.lower, rd, rA, wA, nul, red, 0, st); }
    }
}

void Ldu::precondition(int kind, const MatrixView& A, const double* rA, double* wA)
{
    ensureWork();
    B200_CUDA(cudaMemsetAsync(d_state, 0, sizeof(SolveState), ctx->stream));
    factor(kind, A);
    applyPrecond(kind, A, rA, wA, 0, nullptr, false, C_NONE);
}

// ------------------------------------------------------------------ Krylov drivers
static void ctrl(Ldu& L, int phase)
{
    ProfScope ps_(L.ctx, PC_CTRL);
    k_ctrl<<<1, 1, 0, L.ctx->stream>>>(phase, L.d_state, L.red.result);
    B200_LAUNCH_CHECK();
}

SolvePerf Ldu::solve(const MatrixView& A, const double* source, double* psi, const SolveControls& sc)
{
    ensureWork();
    cudaStream_t s = ctx->stream;
    const int64_t n = nCells;
    SolvePerf perf;
    if (sc.solver == 2) {       // diagonalSolver
        k_diag_solve<<<redBlocks, BLK, 0, s>>>(n, A.diag, source, psi); B200_LAUNCH_CHECK();
        perf.converged = 1;
        return perf;
    }
    std::memset(h_state, 0, sizeof(SolveState));
    h_state->tol = sc.tolerance; h_state->relTol = sc.relTol; h_state->minIter = sc.minIter; h_state->maxIter = sc.maxIter;
    if (nGlobalCells < 0) {
        // global cell count for gAverage(psi): one tiny allreduce, once per addressing
        double nGlobal = (double)n;
        if (ctx->nranks > 1) {
            B200_CUDA(cudaMemcpyAsync(red.result + S_NSLOTS - 1, &nGlobal, sizeof(double), cudaMemcpyHostToDevice, s));
            ctx->allreduceSum(red.result + S_NSLOTS - 1, 1);
            B200_CUDA(cudaMemcpyAsync(&nGlobal, red.result + S_NSLOTS - 1, sizeof(double), cudaMemcpyDeviceToHost, s));
            ctx->sync();
        }
        nGlobalCells = nGlobal;
    }
    const double nGlobal = nGlobalCells;
    h_state->nGlobal = nGlobal;
    static_assert(sizeof(SolveState) <= 256 && sizeof(SolveState) % 4 == 0, "SolveState travels as a kernel argument");
    words_to_device(ctx, d_state, h_state, sizeof(SolveState));
    const SolveState* st = d_state;

    const bool pcg = sc.solver == 0;
    double *wA = vec(0), *pA = vec(1), *rA = vec(2), *qA = vec(3);
    double *yA = vec(0), *AyA = vec(3), *sA = vec(4), *zA = vec(5), *tA = vec(6), *rA0 = vec(7);

    // wA = A psi (+ sum psi); rA = source - wA; normFactor.  Every global sum is followed by the control step that consumes it,
    // fused into the kernel that completes the sum (reduceCtrl / redFor).
    amulFused(A, psi, wA, AM_SUMX, nullptr, C_AVG);
    reduceCtrl(S_SUMPSI, 1, C_AVG);
    if (isBlock && ifaces.empty()) {
        const double* bb = (ghostBelow && A.bouCoeffs) ? A.bouCoeffs[0] : nullptr;
        const double* ba = (ghostAbove && A.bouCoeffs) ? A.bouCoeffs[ghostBelow ? 1 : 0] : nullptr;
        const int64_t items = (int64_t)((dims.nx + BLK - 1) / BLK) * dims.ny * dims.nz;
        const int grid = (int)std::min<int64_t>(items, ctx->persistentBlocks());
        ProfScope ps_(ctx, PC_NORMFACTOR);
        k_normfactor_block<<<grid, BLK, 0, s>>>(dims, A.diag, A.upper, A.lower, bb, ba, source, wA, rA, redFor(C_INIT), st);
        B200_LAUNCH_CHECK();
    } else {
        ProfScope ps_(ctx, PC_NORMFACTOR);
        if (!d_sumA) d_sumA = dev_alloc<double>(n);
        k_sumA_general<<<redBlocks, BLK, 0, s>>>(nCells, d_rowStart, d_rowMid, d_rowFace, A.diag, A.upper, A.lower, d_sumA);
        B200_LAUNCH_CHECK();
        for (size_t p = 0; p < ifaces.size(); p++) {
            auto& it = ifaces[p];
            const int grid = it.hasDuplicates ? 1 : std::max(1, std::min((it.size + BLK - 1) / BLK, 1024));
            k_iface_sumA<<<grid, BLK, 0, s>>>(it.size, it.faceCells, A.bouCoeffs[p], d_sumA, it.hasDuplicates ? 1 : 0);
            B200_LAUNCH_CHECK();
        }
        k_normfactor_general<<<redBlocks, BLK, 0, s>>>(nCells, d_sumA, source, wA, rA, redFor(C_INIT), st);
        B200_LAUNCH_CHECK();
    }
    reduceCtrl(S_NORM, 2, C_INIT);
    factor(sc.precond, A);
    // PCG with a DIC/DILU wave sweep keeps the residual (RF) and the preconditioned residual (WB) in wave order: the sweeps
    // then stream whole rows only, and the two vector updates transpose through shared memory (wave_kernels.cuh)
    static const bool wioOff = [] { const char* e = std::getenv("B200_WAVE_IO"); return e && std::string(e) == "natural"; }();
    const bool wio = pcg && isBlock && useWave && sc.precond >= 2 && !wioOff;
    WaveGeom wg{};
    int wioGrid = 1;
    if (wio) {
        wg = wave_geom(dims);
        wioGrid = waveVecGrid();
        ProfScope ps_(ctx, PC_VECTOR);
        k_wave_from_natural<<<wioGrid, 256, 0, s>>>(wg, rA, wave[8], 0.0, st); B200_LAUNCH_CHECK();
    }
    if (!pcg) { { ProfScope ps_(ctx, PC_VECTOR); k_copy<<<redBlocks, BLK, 0, s>>>(n, rA, rA0, st); B200_LAUNCH_CHECK(); } }

    const int hardMax = std::max(sc.maxIter + 1, sc.minIter);
    // The host only looks at the device's `done` flag (one tiny D2H + sync); iterations enqueued past convergence are no-ops.
    // How many iterations go out before the first look, and per look afterwards, must be the same on every rank -- each
    // iteration enqueues halo exchanges and all-reduces that pair up across ranks -- so the policy uses rank-invariant
    // numbers only: the iteration count of the previous solve on this addressing (a global result; the resident stepper's
    // solves with the shipped controls all run to maxIter + 1) and the global cell count per rank.
    const bool large = nGlobalCells / ctx->nranks >= (double)(1 << 20);
    int enqueued = 0;
    const size_t profFirst = ctx->profRecs.size();
    int chunk = lastIterations > 0 ? std::min(lastIterations, hardMax) : (large ? 1 : 2);
    for (;;) {
        for (int it = 0; it < chunk && enqueued < hardMax; it++, enqueued++) {
            ctx->profTag = enqueued;
            if (pcg) {
                applyPrecond(sc.precond, A, rA, wA, 1, rA, wio, C_PCG_BETA);    // wA = M^-1 rA ; wArA
                reduceCtrl(S_RHO, 1, C_PCG_BETA);
                { ProfScope ps_(ctx, PC_VECTOR);
                  if (wio) k_wave_pcg_p<<<wioGrid, 256, 0, s>>>(wg, wave[9], pA, st);
                  else k_pcg_p<<<redBlocks, BLK, 0, s>>>(n, wA, pA, st);
                  B200_LAUNCH_CHECK(); }
                amulFused(A, pA, qA, AM_DOT_YX, nullptr, C_PCG_ALPHA);          // q = A pA ; wApA
                reduceCtrl(S_DEN, 1, C_PCG_ALPHA);
                { ProfScope ps_(ctx, PC_VECTOR);
                  if (wio) k_wave_pcg_update<<<wioGrid, 256, 0, s>>>(wg, pA, qA, psi, wave[8], redFor(C_ITER_END), st);
                  else k_pcg_update<<<redBlocks, BLK, 0, s>>>(n, pA, qA, psi, rA, redFor(C_ITER_END), st);
                  B200_LAUNCH_CHECK(); }
                reduceCtrl(S_MAGR, 1, C_ITER_END);
            } else {
                { ProfScope ps_(ctx, PC_VECTOR); k_dot<<<redBlocks, BLK, 0, s>>>(n, rA0, rA, redFor(C_BICG_RHO), S_RHO, st); B200_LAUNCH_CHECK(); }
                reduceCtrl(S_RHO, 1, C_BICG_RHO);
                { ProfScope ps_(ctx, PC_VECTOR); k_bicg_p<<<redBlocks, BLK, 0, s>>>(n, rA, AyA, pA, st); B200_LAUNCH_CHECK(); }
                applyPrecond(sc.precond, A, pA, yA, 0, nullptr, false, C_NONE);
                amulFused(A, yA, AyA, AM_DOT_AUXY, rA0, C_BICG_ALPHA);          // rA0 . AyA
                reduceCtrl(S_DEN, 1, C_BICG_ALPHA);
                { ProfScope ps_(ctx, PC_VECTOR); k_bicg_s<<<redBlocks, BLK, 0, s>>>(n, rA, AyA, sA, redFor(C_BICG_SCHECK), st); B200_LAUNCH_CHECK(); }
                reduceCtrl(S_MAGR, 1, C_BICG_SCHECK);
                { ProfScope ps_(ctx, PC_VECTOR); k_bicg_early<<<redBlocks, BLK, 0, s>>>(n, yA, psi, st, &d_state->early); B200_LAUNCH_CHECK(); }
                ctrl(*this, C_BICG_EARLY_END);
                applyPrecond(sc.precond, A, sA, zA, 0, nullptr, false, C_NONE);
                amulFused(A, zA, tA, AM_TT_TS, sA, C_BICG_OMEGA);               // tA.tA, tA.sA
                reduceCtrl(S_TT, 2, C_BICG_OMEGA);
                { ProfScope ps_(ctx, PC_VECTOR); k_bicg_update<<<redBlocks, BLK, 0, s>>>(n, yA, zA, sA, tA, psi, rA, redFor(C_ITER_END2), st); B200_LAUNCH_CHECK(); }
                reduceCtrl(S_MAGR2, 1, C_ITER_END2);
            }
        }
        ctx->profTag = -1;
        words_to_host(ctx, h_state, d_state, sizeof(SolveState));
        ctx->sync();
        if (h_state->done || enqueued >= hardMax) break;
        chunk = large ? 1 : std::min(std::max(chunk, 1) * 2, 32);
    }
    perf.initialResidual = h_state->initRes; perf.finalResidual = h_state->finalRes;
    perf.nIterations = h_state->nIter; perf.converged = h_state->converged; perf.singular = h_state->singular;
    lastIterations = perf.nIterations;
    if (ctx->profOn) ctx->profDropFrom(profFirst, perf.nIterations);       // iterations enqueued past convergence were no-ops
    return perf;
}

unsigned long long selftest_division(Ctx* ctx, unsigned long long n, unsigned long long seed, int expRange)
{
    unsigned long long* d = dev_alloc<unsigned long long>(1);
    unsigned long long h = 0;
    B200_CUDA(cudaMemsetAsync(d, 0, sizeof(h), ctx->stream));
    k_selftest_division<<<ctx->smCount * 8, 256, 0, ctx->stream>>>(n, seed, expRange, d);
    B200_LAUNCH_CHECK();
    d2h(&h, d, 1, ctx->stream);
    ctx->sync();
    dev_free(d);
    return h;
}

} // namespace b200
