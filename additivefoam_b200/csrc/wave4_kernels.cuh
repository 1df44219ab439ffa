// Tile-wavefront DIC/DILU sweeps, three warp roles per tile (k_wave4).  Same tiles, same wave-order arrays, same exchange
// protocol (the datum is its own flag) and -- step for step -- the same arithmetic as k_wave2 (wave_kernels.cuh), hence the same
// bits; what changes is who does what, and when.  B200_WAVE_TRACE on k_wave2 (profiles/r2_wave_trace.md) showed where a sweep of
// a thin slab (1000 x 400 x 31 cells: 50 x 8 tiles, the per-GPU share of the decomposed 100 M-cell case) spent its 0.40 ms:
//   * the COMPUTE warp needs ~960 cycles per 8 steps although the FP64 chain of a step (shuffle, select, DMUL, 3 DADD) is ~60:
//     the 56 LDS + 32 DMUL that bring a group's operands into registers ran before the chain, not under it;
//   * the MEMORY warp needs ~1150 cycles per 8 steps (bulk-copy issue 450, 24 predicated loads, 24 st.shared, 8-16 arming
//     stores), so a tile that follows its producer closely -- every tile of a thin slab -- has no time left to ask again for
//     neighbour elements that were not there yet: each retry costs an L2 round trip (~900 cycles), the tile falls behind
//     until its early loads succeed, and the lag per tile-to-tile hop settles at ~5 us where the dependency needs < 1.5.
// Here:
//   * COMPUTE warp: a slot is consumed in two halves; the operands of the next half are fetched (LDS + the products that do
//     not depend on the chain) in the same basic block as the chain of the current half, so the compiler schedules them into
//     the chain's stall slots.  The next slot is only prefetched when its barrier has already completed (mbarrier.test_wait);
//     a tile that is waiting for its producer falls back to wait-then-fetch.
//   * HALO warp: nothing but the neighbour tiles' elements.  One element per lane (32 lanes cover the 8 steps x 4 y-edge
//     elements of a group, twice 32 the 8 x 8 z-edge elements), so first loads, a retry pass and the hand-over are 3
//     instructions each instead of 16-24, and DEPTH groups of look-ahead cost 3 registers each.  A retry pass asks for every
//     missing element of every group in flight at once.
//   * STREAM warp: everything that is off the dependency chain -- the bulk async copies of the tile's rows into the ring (one
//     lane per array), the natural-order residual / diagonal, the arming of the exchange arrays.
#pragma once
#include "wave_kernels.cuh"

namespace b200 {

// operands of HS consecutive steps of one lane, in registers
template <int OP, bool DILU, bool FASTDIV, int HS> struct WaveOps {
    double c0[HS], c1[HS], c2[HS], c3[HS], in[HS];
    double n1[(OP == 0 && DILU) ? HS : 1], n2[(OP == 0 && DILU) ? HS : 1], n3[(OP == 0 && DILU) ? HS : 1];
};

// ring slot -> registers, with everything that does not depend on the chain formed on the way ((rD*coef) exactly as the
// reference's loops round it; rD*rA)
template <int OP, bool DILU, bool WIO, bool FASTDIV, int NARR, int HS, class Slot>
__device__ __forceinline__ void wave4_fetch_one(const Slot& S, int s0, int lane, WaveOps<OP, DILU, FASTDIV, HS>& r, int s)
{
    const int o = (s0 + s) * 32 + lane;
    if (OP == 0) {
        r.c0[s] = WIO ? S.arr[NARR - 1][o] : S.in[o]; r.c1[s] = S.arr[0][o]; r.c2[s] = S.arr[1][o]; r.c3[s] = S.arr[2][o]; r.in[s] = 0.0;
        if (DILU) { r.n1[s] = S.arr[3][o]; r.n2[s] = S.arr[4][o]; r.n3[s] = S.arr[5][o]; }
    } else {
        const double rd = OP == 1 ? S.arr[0][o] : S.arr[5][o];
        r.in[s] = (OP == 1 && !WIO) ? S.in[o] : S.arr[4][o];
        r.c0[s] = OP == 1 ? rd * r.in[s] : S.arr[0][o];
        r.c1[s] = rd * S.arr[1][o]; r.c2[s] = rd * S.arr[2][o]; r.c3[s] = rd * S.arr[3][o];
    }
}
template <int OP, bool DILU, bool WIO, bool FASTDIV, int NARR, int HS, class Slot>
__device__ __forceinline__ void wave4_fetch(const Slot& S, int s0, int lane, WaveOps<OP, DILU, FASTDIV, HS>& r)
{
#pragma unroll
    for (int s = 0; s < HS; s++) wave4_fetch_one<OP, DILU, WIO, FASTDIV, NARR, HS, Slot>(S, s0, lane, r, s);
}

// HS steps of the recurrence (t0 = lowest step of the half; the backward sweep walks it downwards).  The arithmetic of a step is
// that of wave2_compute, operation for operation.
// PRE: the operands of the NEXT half (slot Sn, first step sn0) are fetched into `nxt` step by step between the steps of this chain --
// written interleaved in the source because the compiler keeps a separate fetch loop in front of the chain (the divergence check
// of the first shuffle ends its scheduling window), where the warp then waits out the LDS latency and 16 DMULs before every half.
// halo operands of HS steps (the neighbour tiles' results on this tile's upstream edges; for the factor sweep with their refined reciprocals)
template <int OP, bool FASTDIV, int HS> struct WaveHalo {
    double hy[HS], hz[HS];
    double yY[(OP == 0 && FASTDIV) ? HS : 1], yZ[(OP == 0 && FASTDIV) ? HS : 1];
};
template <int OP, bool DILU, bool DOT, bool RAMP, bool WIO, bool FASTDIV, int HS, bool PRE = false, int NARR = 1, class Slot = int>
__device__ __forceinline__ void wave4_chain(const WaveOps<OP, DILU, FASTDIV, HS>& r, const WaveHalo<OP, FASTDIV, HS>& h, const WaveOut<OP>& out, int t0, int lane,
                                            int skew, int nx, bool pencil, bool edgeInY, bool edgeInZ, bool hasY, bool hasZ, double& last, double& acc, double& held,
                                            unsigned& bad, const Slot* Sn = nullptr, int sn0 = 0, WaveOps<OP, DILU, FASTDIV, HS>* nxt = nullptr)
{
    constexpr bool bwd = OP == 2;
    const unsigned full = 0xffffffffu;
    const size_t gofs = (size_t)t0 * 32;
    double* xg = out.X + gofs;
    const double sent = __longlong_as_double((long long)WSENT);
#pragma unroll
    for (int ss = 0; ss < HS; ss++) {
        const int s = bwd ? HS - 1 - ss : ss;
        if constexpr (PRE) wave4_fetch_one<OP, DILU, WIO, FASTDIV, NARR, HS, Slot>(*Sn, sn0, lane, *nxt, ss);
        bool act = pencil;
        if (RAMP) { const int i = t0 + s - skew; act = act && i >= 0 && i < nx; }
        const double c0 = r.c0[s], c1 = r.c1[s], c2 = r.c2[s], c3 = r.c3[s];
        double vy = bwd ? __shfl_down_sync(full, last, 1) : __shfl_up_sync(full, last, 1);
        double vz = bwd ? __shfl_down_sync(full, last, WTJ) : __shfl_up_sync(full, last, WTJ);
        vy = edgeInY ? h.hy[s] : vy;
        vz = edgeInZ ? h.hz[s] : vz;
        double res;
        if (OP == 0) {
            // rD[u] -= upper*upper/rD[l] (DIC) or upper*lower/rD[l] (DILU), faces ascending: z, y, x lower neighbours
            const double numz = DILU ? r.n3[s] : c3 * c3, numy = DILU ? r.n2[s] : c2 * c2, numx = DILU ? r.n1[s] : c1 * c1;
            const bool hasX = RAMP ? (t0 + s - skew > 0 && t0 + s - skew < nx) : true;
            double tz, ty, tx;
            if (FASTDIV) {
                const double nz_ = hasZ ? numz : 0.0, ny_ = hasY ? numy : 0.0, nx_ = hasX ? numx : 0.0;
                double yy = bwd ? __shfl_down_sync(full, held, 1) : __shfl_up_sync(full, held, 1);
                double yz = bwd ? __shfl_down_sync(full, held, WTJ) : __shfl_up_sync(full, held, WTJ);
                yy = edgeInY ? h.yY[s] : yy;
                yz = edgeInZ ? h.yZ[s] : yz;
                tz = div_by(nz_, vz, yz); ty = div_by(ny_, vy, yy); tx = div_by(nx_, last, held);
            } else {
                tz = (hasZ ? numz : 1.0) / vz; ty = (hasY ? numy : 1.0) / vy; tx = (hasX ? numx : 1.0) / last;
                tz = hasZ ? tz : 0.0; ty = hasY ? ty : 0.0; tx = hasX ? tx : 0.0;
            }
            double dd = c0;
            dd -= tz;
            dd -= ty;
            dd -= tx;
            res = RAMP ? (act ? dd : 1.0) : dd;
            xg[s * 32] = publishable(res);               // publish the raw diagonal
            double rd;
            if (FASTDIV) {
                bad |= div_in_range(res, false) ? 0u : 1u;
                held = div_refine(res);                  // on the chain: the next step divides by this entry
                rd = rcp_fast(res);                      // rD = 1/rD (off the chain)
            } else rd = 1.0 / res;
            (out.A0 + gofs)[s * 32] = rd;
        } else {
            // forward:  wA = rD*rA; wA[u] -= rD[u]*coef[f]*wA[l], faces ascending (z, y, x lower neighbours)
            // backward: wA[l] -= rD[l]*upper[f]*wA[u], faces descending (z, y, x upper neighbours)
            const double m3 = c3 * vz, m2 = c2 * vy, m1 = c1 * last;
            double w = c0;
            w -= m3;
            w -= m2;
            w -= m1;
            res = RAMP ? (act ? w : 0.0) : w;
            xg[s * 32] = publishable(res);               // publish: the next tile is waiting for exactly this row
            if (OP == 2) {
                if (DOT) acc += res * r.in[s];
                if (act && !WIO) (out.wA + t0)[s] = res;
                (out.WF + gofs)[s * 32] = sent;          // re-arm the forward sweep's exchange array
            }
        }
        last = res;
    }
}

// ring slot of k_wave4: the streamed arrays (+ the natural-order input where there is one); the halo has its own channel
template <int NARR, int GSZ, bool NEEDIN> struct alignas(128) WaveSlot4 {
    double arr[NARR][GSZ * 32];
    double in[NEEDIN ? GSZ * 32 : 16];
};
template <int OP, bool DILU, bool WIO, int GSZ, int RNG> struct WaveSmem4 {
    static constexpr int NARR = OP == 0 ? (DILU ? 6 : 3) + (WIO ? 1 : 0) : (OP == 1 ? (WIO ? 5 : 4) : 6);
    static constexpr bool NEEDIN = !WIO && OP != 2;
    WaveSlot4<NARR, GSZ, NEEDIN> slot[RNG];
    unsigned long long full[RNG], empty[RNG];
    int tile;
};

template <int OP, bool DILU, bool DOT, bool WIO, bool FASTDIV, int GSZ, int RNG>
__global__ void __launch_bounds__(96, (RNG <= 3 ? 4 : RNG <= 4 ? 3 : 2)) k_wave4(WaveArgs a)
{
    if (a.st != nullptr && a.st->done) return;
    if (OP == 0 && FASTDIV && a.ticket[2] != 0) return;
    if (OP == 0 && !FASTDIV && a.ticket[1] == 0) return;
    using Smem = WaveSmem4<OP, DILU, WIO, GSZ, RNG>;
    constexpr int NARR = Smem::NARR;
    using Slot = WaveSlot4<NARR, GSZ, Smem::NEEDIN>;
    constexpr bool NEEDIN = Smem::NEEDIN;          // natural-order residual (forward) / matrix diagonal (factor)
    extern __shared__ __align__(128) unsigned char smemRaw[];
    Smem& sm = *reinterpret_cast<Smem*>(smemRaw);
    __shared__ const double* srcS[8];              // the streamed arrays of this sweep in slot order, indexable by lane
    // The halo channel: the neighbour tiles' edge elements, step by step (ring over the steps in processing order), and two
    // progress counters -- steps whose elements have all arrived (halo warp -> compute warp), steps consumed (back-pressure).
    // Finer than the ring slots on purpose: a slot-sized hand-over made every group wait for a whole poll pass, its hand-over and
    // the compute warp's wake-up in sequence, which a tile close behind its producer could not sustain at the producer's pace.
    constexpr int RH = 64;
    __shared__ double hyRing[RH * WTK], hzRing[RH * WTJ];
    __shared__ unsigned haloProg, compProg;
    const WaveGeom& g = a.g;
    const unsigned full = 0xffffffffu;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tj = lane & (WTJ - 1), tk = lane / WTJ;
    constexpr bool bwd = OP == 2;
    if (threadIdx.x == 0) {
        sm.tile = atomicAdd(a.ticket, 1);                  // topological hand-out: one tile per CTA
        // full: the stream warp's arrive.expect_tx (bulk bytes); empty: the compute warp's arrive.  The halo has its own channel.
        for (int r = 0; r < RNG; r++) { mbar_init(&sm.full[r], 1); mbar_init(&sm.empty[r], 1); }
        haloProg = 0u; compProg = 0u;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (OP == 2) { srcS[0] = a.WF; srcS[1] = a.QX; srcS[2] = a.QY; srcS[3] = a.QZ; srcS[4] = a.RF; srcS[5] = a.A0; }
        else if (OP == 1) { srcS[0] = a.A0; srcS[1] = a.PX; srcS[2] = a.PY; srcS[3] = a.PZ; if (WIO) srcS[NARR - 1] = a.RF; }
        else {
            srcS[0] = a.PX; srcS[1] = a.PY; srcS[2] = a.PZ;
            if (DILU) { srcS[3] = a.NX; srcS[4] = a.NY; srcS[5] = a.NZ; }
            if (WIO) srcS[NARR - 1] = a.DG;
        }
    }
    __syncthreads();
    int T = sm.tile;
    if (T >= g.nTiles) return;
    T = a.order[bwd ? g.nTiles - 1 - T : T];
    const bool tracing = a.trace != nullptr;
    auto stamp = [&](int what) {
        if (tracing && lane == 0) { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); a.trace[WTRACE * T + what] = t; }
    };
    if (warp == 0) stamp(11);
    const int K = T / g.nJ, J = T - K * g.nJ;
    const size_t tileBase = (size_t)T * g.nStepsPad * 32;
    double* X = OP == 0 ? a.RF : OP == 1 ? a.WF : a.WB;      // exchange array of this sweep
    const int nGroups = g.nStepsPad / GSZ;
    const double one_or_zero = OP == 0 ? 1.0 : 0.0;
    auto tbOf = [&](int grp) { return bwd ? g.nStepsPad - GSZ * (grp + 1) : GSZ * grp; };      // lowest step of the group
    const int upJ = bwd ? T + 1 : T - 1, upK = bwd ? T + g.nJ : T - g.nJ;
    const int dtY = bwd ? -(WTJ - 1) : WTJ - 1, dtZ = bwd ? -(WTK - 1) : WTK - 1;
    long long ph[4] = {0, 0, 0, 0};

    if (warp == 1) {
        // ------------------------------------------------------------------ halo warp: one neighbour element per lane
        constexpr int NEY = GSZ * WTK, NEZ = GSZ * WTJ;            // y-edge / z-edge elements of a group
        constexpr int NZ = (NEZ + 31) / 32;
        static_assert(NEY <= 32, "one y-edge element per lane");
        constexpr int DEPTH = 4;
        // y-edge element of this lane: step sY of the group, consumer lane (tjE, tkY)
        const int sY = lane / WTK, tkY = lane % WTK, tjE = bwd ? WTJ - 1 : 0;
        const int jY = J * WTJ + tjE, kY = K * WTK + tkY;
        const bool wantY = lane < NEY && jY < g.ny && kY < g.nz && (bwd ? jY < g.ny - 1 : jY > 0);
        const int skewY = tjE + tkY;
        const double* XYp = X + ((size_t)(wantY ? upJ : T) * g.nStepsPad + (wantY ? dtY : 0)) * 32 + (bwd ? WTJ * tkY : WTJ - 1 + WTJ * tkY);
        // z-edge elements: e = lane + 32 r -> step e / WTJ, consumer lane (tjZ, tkE)
        const int tjZ = lane % WTJ, tkE = bwd ? WTK - 1 : 0;
        const int jZ = J * WTJ + tjZ, kZ = K * WTK + tkE;
        const bool wantZ = jZ < g.ny && kZ < g.nz && (bwd ? kZ < g.nz - 1 : kZ > 0);
        const int skewZ = tjZ + tkE;
        const double* XZp = X + ((size_t)(wantZ ? upK : T) * g.nStepsPad + (wantZ ? dtZ : 0)) * 32 + (bwd ? tjZ : tjZ + WTJ * (WTK - 1));
        const int sZ0 = lane / WTJ;
        double vY[DEPTH], vZ[DEPTH][NZ];
        auto first = [&](int grp, double& y, double (&z)[NZ]) {
            const int tb = tbOf(grp);
            y = one_or_zero;
            { const int t = tb + sY, i = t - skewY; ld_relaxed_if(y, XYp + (size_t)t * 32, wantY && i >= 0 && i < g.nx); }
#pragma unroll
            for (int r = 0; r < NZ; r++) {
                z[r] = one_or_zero;
                const int sZ = sZ0 + r * (32 / WTJ);
                const int t = tb + sZ, i = t - skewZ;
                ld_relaxed_if(z[r], XZp + (size_t)t * 32, wantZ && sZ < GSZ && i >= 0 && i < g.nx);
            }
        };
        // ask again for what has not arrived (only a load that was made can have returned the sentinel)
        auto again = [&](int grp, double& y, double (&z)[NZ]) {
            const bool valid = grp < nGroups;
            const int tb = tbOf(valid ? grp : 0);
            ld_relaxed_if(y, XYp + (size_t)(tb + sY) * 32, valid && is_sentinel(y));
#pragma unroll
            for (int r = 0; r < NZ; r++) ld_relaxed_if(z[r], XZp + (size_t)(tb + sZ0 + r * (32 / WTJ)) * 32, valid && is_sentinel(z[r]));
        };
#pragma unroll
        for (int q = 0; q < DEPTH; q++) {
            vY[q] = 0.0;
#pragma unroll
            for (int r = 0; r < NZ; r++) vZ[q][r] = 0.0;
        }
        // How many groups ahead the first loads go (1 = none ahead, at most DEPTH).  NOT "the more the better": loads that go out
        // before the producer has written come back as the sentinel and cost a retry, and a consumer that retries is slower than
        // its producer -- it falls back until its early loads succeed.  The look-ahead distance is therefore the lag a tile settles
        // at behind its producer (profiles/r2_wave_trace.md: 4.4 us per hop with 4 groups, all tiles at the same pace so that the
        // lag never shrinks again).
        const int hd = a.haloDepth < 1 ? 1 : a.haloDepth > DEPTH ? DEPTH : a.haloDepth;
#pragma unroll
        for (int q = 0; q < DEPTH - 1; q++) if (q < hd - 1 && q < nGroups) first(q, vY[q], vZ[q]);
        // processing-order offset (within its group) of this lane's elements
        const int pY = bwd ? GSZ - 1 - sY : sY;
        for (int grp0 = 0; grp0 < nGroups; grp0 += DEPTH) {
#pragma unroll
            for (int q = 0; q < DEPTH; q++) {
                const int grp = grp0 + q;
                if (grp < nGroups) {
#pragma unroll
                    for (int d = 1; d <= DEPTH; d++)
                        if (d == hd && grp + d - 1 < nGroups) first(grp + d - 1, vY[(q + d - 1) % DEPTH], vZ[(q + d - 1) % DEPTH]);
                    const int pBase = grp * GSZ;
                    int delivered = 0;
                    long long c0 = tracing ? clock64() : 0;
                    for (;;) {
                        // leading steps (processing order) of this group whose elements are all here
                        const unsigned yOk = __ballot_sync(full, !is_sentinel(vY[q]));
                        unsigned zOk[NZ];
#pragma unroll
                        for (int r = 0; r < NZ; r++) zOk[r] = __ballot_sync(full, !is_sentinel(vZ[q][r]));
                        int k = 0;
#pragma unroll
                        for (int i = 0; i < GSZ; i++) {
                            const int sM = bwd ? GSZ - 1 - i : i;                                    // memory-order offset of processing step i
                            const bool okY = ((yOk >> (WTK * sM)) & ((1u << WTK) - 1u)) == ((1u << WTK) - 1u);
                            const bool okZ = ((zOk[sM / (32 / WTJ)] >> (WTJ * (sM % (32 / WTJ)))) & ((1u << WTJ) - 1u)) == ((1u << WTJ) - 1u);
                            if (okY && okZ && k == i) k = i + 1;
                        }
                        if (k > delivered) {
                            // back-pressure: the ring entries about to be overwritten have been consumed (never waits in practice)
                            while ((int)ld_acquire_cta_u32(&compProg) + RH < pBase + k) __nanosleep(200);      // (a tight spin on shared memory slows the compute warp's LDS / SHFL)
                            if (lane < NEY && pY >= delivered && pY < k) hyRing[((pBase + pY) & (RH - 1)) * WTK + tkY] = vY[q];
#pragma unroll
                            for (int r = 0; r < NZ; r++) {
                                const int sZ = sZ0 + r * (32 / WTJ);
                                const int pZ = bwd ? GSZ - 1 - sZ : sZ;
                                if (sZ < GSZ && pZ >= delivered && pZ < k) hzRing[((pBase + pZ) & (RH - 1)) * WTJ + tjZ] = vZ[q][r];
                            }
                            __syncwarp();
                            if (lane == 0) st_release_cta_u32(&haloProg, (unsigned)(pBase + k));
                            delivered = k;
                        }
                        if (k == GSZ) break;
                        // everything that is missing of every group in flight, oldest first (the youngest group's first loads may
                        // still be on their way: its predicate is evaluated last)
#pragma unroll
                        for (int d = 0; d < DEPTH; d++) if (d < hd) again(grp + d, vY[(q + d) % DEPTH], vZ[(q + d) % DEPTH]);
                        if (tracing) ph[1] += 1;
                    }
                    if (tracing) { const long long c1 = clock64(); ph[0] += c1 - c0; }
                    if (grp < 5) stamp(grp);
                }
            }
        }
        if (tracing && lane == 0) { a.trace[WTRACE * T + 15] = ph[0]; a.trace[WTRACE * T + 18] = ph[1]; a.trace[WTRACE * T + 12] = ph[2]; }
        return;
    }

    // per-lane geometry of the compute and stream warps
    const int j = J * WTJ + tj, k = K * WTK + tk;
    const bool pencil = j < g.ny && k < g.nz;
    const int skew = tj + tk;
    const int64_t cbase = (int64_t)g.nx * (j + (int64_t)g.ny * k) - skew;        // natural index c = cbase + t

    if (warp == 2) {
        // ------------------------------------------------------------------ stream warp: bulk copies, natural-order input, arming
        const double sent = __longlong_as_double((long long)WSENT);
        auto loadIn = [&](int grp, double (&in)[GSZ]) {
            const int tb = tbOf(grp);
            // a private stream per lane (one 128-byte line per 16 steps): pull it into L2 well ahead
            if (pencil) { const int tp = tb + 12 * GSZ, ip = tp - skew; if (ip >= 0 && ip < g.nx) prefetch_l2(a.rA + cbase + tp); }
#pragma unroll
            for (int s = 0; s < GSZ; s++) {
                const int t = tb + s, i = t - skew;
                in[s] = OP == 0 ? 1.0 : 0.0;
                ldnc_if(in[s], a.rA + cbase + t, pencil && i >= 0 && i < g.nx);
            }
        };
        auto body = [&](int grp, double (&cur)[GSZ], double (&nxt)[GSZ]) {
            if (grp >= nGroups) return;
            const int slot = grp % RNG;
            const int tb = tbOf(grp);
            mbar_wait(&sm.empty[slot], (((unsigned)(grp / RNG)) & 1u) ^ 1u);
            Slot& S = sm.slot[slot];
            if (lane < NARR) bulk_g2s(&S.arr[0][0] + lane * (GSZ * 32), srcS[lane] + tileBase + (size_t)tb * 32, GSZ * 32 * sizeof(double), &sm.full[slot]);
            if (NEEDIN) {
                if (grp + 1 < nGroups) loadIn(grp + 1, nxt);
#pragma unroll
                for (int s = 0; s < GSZ; s++) S.in[s * 32 + lane] = cur[s];
                __syncwarp();
            }
            if (lane == 0) mbar_arrive_expect_tx(&sm.full[slot], NARR * GSZ * 32 * sizeof(double));
#pragma unroll
            for (int s = 0; s < GSZ; s++) {
                const size_t row = tileBase + (size_t)(tb + s) * 32 + lane;
                if (OP == 0) { a.WF[row] = sent; a.WB[row] = sent; }
                if (OP == 1) { if (!WIO) a.RF[row] = cur[s]; a.WB[row] = sent; }
            }
        };
        double inA[NEEDIN ? GSZ : 1], inB[NEEDIN ? GSZ : 1];
        if (NEEDIN) loadIn(0, reinterpret_cast<double (&)[GSZ]>(inA));
        for (int grp = 0; grp < nGroups; grp += 2) {
            body(grp, reinterpret_cast<double (&)[GSZ]>(inA), reinterpret_cast<double (&)[GSZ]>(inB));
            body(grp + 1, reinterpret_cast<double (&)[GSZ]>(inB), reinterpret_cast<double (&)[GSZ]>(inA));
        }
        return;
    }

    // ---------------------------------------------------------------------- compute warp
    const bool hasY = pencil && (bwd ? j < g.ny - 1 : j > 0);
    const bool hasZ = pencil && (bwd ? k < g.nz - 1 : k > 0);
    const bool edgeInY = bwd ? tj == WTJ - 1 : tj == 0;
    const bool edgeInZ = bwd ? tk == WTK - 1 : tk == 0;
    double last = one_or_zero, acc = 0.0, held = 0.0;
    if (OP == 0 && FASTDIV) held = div_refine(last);      // factor sweep: `held` carries the refined reciprocal of `last`
    unsigned bad = 0;
    WaveOut<OP> out;
    out.X = X + tileBase + lane;
    out.wA = (OP == 2 && pencil && !WIO) ? a.wA + cbase : nullptr;
    out.WF = a.WF + tileBase + lane;
    out.A0 = a.A0 + tileBase + lane; out.PX = a.PX + tileBase + lane; out.PY = a.PY + tileBase + lane; out.PZ = a.PZ + tileBase + lane;
    out.QX = a.QX + tileBase + lane; out.QY = a.QY + tileBase + lane; out.QZ = a.QZ + tileBase + lane;
    constexpr int HS = GSZ / 2;
    using Ops = WaveOps<OP, DILU, FASTDIV, HS>;
    // shared-window addresses of the barriers, formed once (the generic-to-shared conversion costs an S2R every time it is redone)
    const unsigned fullBase = smem_u32(&sm.full[0]), emptyBase = smem_u32(&sm.empty[0]);
    auto waitFull = [&](int slot, unsigned parity) {
        asm volatile("{\n\t.reg .pred P1;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t@P1 bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}"
                     ::"r"(fullBase + 8u * (unsigned)slot), "r"(parity) : "memory");
    };
    auto testFull = [&](int slot, unsigned parity) {
        unsigned ok;
        asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.u32 %0, 1, 0, P1;\n\t}"
                     : "=r"(ok) : "r"(fullBase + 8u * (unsigned)slot), "r"(parity) : "memory");
        return ok != 0;
    };
    auto arriveEmpty = [&](int slot) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(emptyBase + 8u * (unsigned)slot) : "memory"); };
    // state of the walk over the ring: the slot and phase parity of the group in hand (no modulo in the loop)
    int slot = 0; unsigned parity = 0u;
    int tb = tbOf(0);
    using Halo = WaveHalo<OP, FASTDIV, HS>;
    int pDone = 0;                                // steps completed, processing order
    // ring -> registers for the HS steps from memory-order step t0
    auto loadHalo = [&](Halo& h, int t0) {
#pragma unroll
        for (int s = 0; s < HS; s++) {
            const int p = bwd ? g.nStepsPad - 1 - (t0 + s) : t0 + s;
            h.hy[s] = hyRing[(p & (RH - 1)) * WTK + tk];
            h.hz[s] = hzRing[(p & (RH - 1)) * WTJ + tj];
            if (OP == 0 && FASTDIV) { h.yY[s] = div_refine(h.hy[s]); h.yZ[s] = div_refine(h.hz[s]); }
        }
    };
    bool haveHalo = false;                        // the halo of the half about to start is in registers already
    // the halo of this half (wait for the halo warp if it was not prefetched), and -- if it has been delivered already -- the next
    // half's, whose loads then complete under this half's chain
    auto halos = [&](Halo& hcur, Halo& hnxt, int t0, int t0next, bool haveNextHalf) {
        const unsigned prog = ld_acquire_cta_u32(&haloProg);
        if (!haveHalo) {
            const unsigned need = (unsigned)(pDone + HS);
            if (prog < need) {
                const long long c0 = tracing ? clock64() : 0;
                while (ld_acquire_cta_u32(&haloProg) < need) {}
                if (tracing) ph[1] += clock64() - c0;
            }
            loadHalo(hcur, t0);
        }
        haveHalo = haveNextHalf && prog >= (unsigned)(pDone + 2 * HS);
        if (haveHalo) loadHalo(hnxt, t0next);
    };
    auto halfDone = [&](bool publish) {
        pDone += HS;
        if (publish) {          // back-pressure for the halo ring (RH steps deep): once per group is plenty
            // A plain (volatile) store, no release: what it guards are this warp's ring READS of positions < pDone, and those have
            // completed -- the chain consumed their values.  st.release here compiles to MEMBAR.ALL.CTA, which also waits for the
            // publishing global stores of the whole group: ~450 cycles per group on the critical path of every tile.
            __syncwarp();
            if (lane == 0) asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(smem_u32(&compProg)), "r"((unsigned)pDone) : "memory");
        }
    };
    // first half of group `grp` (operands in `cur`), prefetching the second half of the same slot
    auto firstHalf = [&](int grp, Ops& cur, Ops& nxt, Halo& h, Halo& hn) {
        const Slot& S = sm.slot[slot];
        const int s0 = bwd ? HS : 0, sn0 = bwd ? 0 : HS;
        const int t0 = tb + s0;
        halos(h, hn, t0, tb + sn0, true);
        if (t0 >= WTJ + WTK - 2 && t0 + HS - 1 < g.nx)
            wave4_chain<OP, DILU, DOT, false, WIO, FASTDIV, HS, true, NARR, Slot>(cur, h, out, t0, lane, skew, g.nx, pencil, edgeInY, edgeInZ, hasY, hasZ, last, acc, held, bad, &S, sn0, &nxt);
        else {
            wave4_chain<OP, DILU, DOT, true, WIO, FASTDIV, HS>(cur, h, out, t0, lane, skew, g.nx, pencil, edgeInY, edgeInZ, hasY, hasZ, last, acc, held, bad);
            wave4_fetch<OP, DILU, WIO, FASTDIV, NARR, HS, Slot>(S, sn0, lane, nxt);
        }
        halfDone(false);
    };
    // second half of group `grp`; the first half of the next group is prefetched only if its slot has landed already
    auto secondHalf = [&](int grp, Ops& cur, Ops& nxt, Halo& h, Halo& hn) {
        const int s0 = bwd ? 0 : HS, sn0 = bwd ? HS : 0;
        const int t0 = tb + s0;
        const bool haveNext = grp + 1 < nGroups;
        const int nslot = slot + 1 == RNG ? 0 : slot + 1;
        const unsigned nparity = slot + 1 == RNG ? parity ^ 1u : parity;
        const Slot& Sn = sm.slot[nslot];
        halos(h, hn, t0, tb + (bwd ? -GSZ : GSZ) + sn0, haveNext);
        const bool pre = haveNext && __all_sync(full, testFull(nslot, nparity));
        const bool interior = t0 >= WTJ + WTK - 2 && t0 + HS - 1 < g.nx;
        if (pre && interior)
            wave4_chain<OP, DILU, DOT, false, WIO, FASTDIV, HS, true, NARR, Slot>(cur, h, out, t0, lane, skew, g.nx, pencil, edgeInY, edgeInZ, hasY, hasZ, last, acc, held, bad, &Sn, sn0, &nxt);
        else {
            if (interior) wave4_chain<OP, DILU, DOT, false, WIO, FASTDIV, HS>(cur, h, out, t0, lane, skew, g.nx, pencil, edgeInY, edgeInZ, hasY, hasZ, last, acc, held, bad);
            else wave4_chain<OP, DILU, DOT, true, WIO, FASTDIV, HS>(cur, h, out, t0, lane, skew, g.nx, pencil, edgeInY, edgeInZ, hasY, hasZ, last, acc, held, bad);
            if (haveNext) {
                if (!pre) {
                    const long long c0 = tracing ? clock64() : 0;
                    waitFull(nslot, nparity);
                    if (tracing) ph[0] += clock64() - c0;
                }
                wave4_fetch<OP, DILU, WIO, FASTDIV, NARR, HS, Slot>(Sn, sn0, lane, nxt);
            }
        }
        halfDone(true);
        if (lane == 0) arriveEmpty(slot);
        if (grp < 5) stamp(5 + grp);
        if (grp == nGroups - 1) stamp(10);
        if (a.traceG != nullptr && lane == 0) { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); a.traceG[(size_t)T * nGroups + grp] = t; }
        slot = nslot; parity = nparity;
        tb += bwd ? -GSZ : GSZ;
    };
    Ops opsA, opsB;
    const long long cStart = tracing ? clock64() : 0;
    waitFull(0, 0u);
    wave4_fetch<OP, DILU, WIO, FASTDIV, NARR, HS, Slot>(sm.slot[0], bwd ? HS : 0, lane, opsA);
    Halo haloA, haloB;
    for (int grp = 0; grp < nGroups; grp++) {
        firstHalf(grp, opsA, opsB, haloA, haloB);
        secondHalf(grp, opsB, opsA, haloB, haloA);
    }
    if (DOT) {
        acc = warp_reduce<Sum>(acc);
        if (lane == 0) a.tilePartial[T] = acc;
    }
    if (OP == 0 && FASTDIV && __any_sync(full, bad != 0) && lane == 0) atomicOr(a.ticket + 1, 1);
    if (tracing && lane == 0) { a.trace[WTRACE * T + 20] = ph[0] + ph[1]; a.trace[WTRACE * T + 21] = clock64() - cStart - ph[0] - ph[1]; a.trace[WTRACE * T + 22] = ph[1]; }
}

} // namespace b200
