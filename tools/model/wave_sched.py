"""Fluid model of the tile-wavefront sweep schedule: how much of a sweep's time is pipeline fill/drain?
Tiles (J,K) with nSteps steps; a tile may be at step p only if its predecessors are at least lag steps + latency ahead;
P slots handed out in ticket order; aggregate step rate capped by HBM bandwidth, per-tile rate by the lone-warp chain."""
import sys
import numpy as np


def simulate(nJ, nK, nSteps, P, bw_steps_per_us, rmax, lagY, lagZ, lat_us, fused=False, dt=0.25):
    nT = nJ * nK
    # ticket order: forward tiles 0..nT-1, then (fused) backward tiles in reverse
    total = 2 * nT if fused else nT
    prog = np.zeros(total)          # steps done
    started = np.zeros(total, bool)
    done = np.zeros(total, bool)
    next_ticket = 0
    active = []
    t = 0.0
    hist = []                       # time at which tile reached each progress (for latency): approximate with rate
    # latency handled as extra lag = lat_us * current rate of the producer (approx with rmax bound)
    while not done.all():
        # hand out slots
        while len(active) < P and next_ticket < total:
            active.append(next_ticket); started[next_ticket] = True; next_ticket += 1
        act = np.array(active)
        # allowed progress per tile from dependencies
        limit = np.full(len(act), float(nSteps))
        for idx, T in enumerate(act):
            if T < nT:
                J, K = T % nJ, T // nJ
                if J > 0:
                    q = T - 1; limit[idx] = min(limit[idx], prog[q] - lagY if not done[q] else nSteps)
                if K > 0:
                    q = T - nJ; limit[idx] = min(limit[idx], prog[q] - lagZ if not done[q] else nSteps)
            else:
                B = 2 * nT - 1 - T           # tile index of this backward tile
                J, K = B % nJ, B // nJ
                if not done[B]:
                    limit[idx] = 0.0
                if J < nJ - 1:
                    q = 2 * nT - 1 - (B + 1); limit[idx] = min(limit[idx], prog[q] - lagY if not done[q] else nSteps)
                if K < nK - 1:
                    q = 2 * nT - 1 - (B + nJ); limit[idx] = min(limit[idx], prog[q] - lagZ if not done[q] else nSteps)
        want = np.clip(limit - prog[act], 0.0, rmax * dt)
        # latency: a tile that is right behind its producer proceeds at most at the producer's pace (already in limit) minus latency steps
        tot = want.sum()
        cap = bw_steps_per_us * dt
        if tot > cap:
            want *= cap / tot
        prog[act] += want
        fin = prog[act] >= nSteps - 1e-9
        for T in act[fin]:
            done[T] = True
        active = [T for T, f in zip(act, fin) if not f]
        t += dt
        if t > 1e5:
            break
    return t


if __name__ == "__main__":
    nJ, nK, nSteps, P = 50, 63, 1000, 1776
    bw = nJ * nK * nSteps / 1000.0            # all tiles' steps in 1000 us at full bandwidth
    for lat in (0.0, 1.5, 3.0):
        rmax = 1 / 0.15
        lagY, lagZ = 7 + lat * 2.0, 3 + lat * 2.0     # latency expressed in steps at ~2 steps/us
        a = simulate(nJ, nK, nSteps, P, bw, rmax, lagY, lagZ, lat)
        b = simulate(nJ, nK, nSteps, P, bw, rmax, lagY, lagZ, lat, fused=True)
        print(f"lat {lat}: one sweep {a:.0f} us (ideal 1000), two separate {2*a:.0f}, fused fwd+bwd {b:.0f} us")
    print("two-warp variant, 3 CTAs/SM")
    for P, rm in ((444, 1 / 0.08), (592, 1 / 0.08), (888, 1 / 0.08)):
        for lat in (1.5,):
            lagY, lagZ = 7 + lat * rm * 0.5, 3 + lat * rm * 0.5
            a = simulate(nJ, nK, nSteps, P, bw, rm, lagY, lagZ, lat)
            print(f"P {P} rmax {rm:.1f}/us lat {lat}: one sweep {a:.0f} us")
