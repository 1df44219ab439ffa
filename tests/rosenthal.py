"""Rosenthal's quasi-steady solution of a point source moving over a semi-infinite adiabatic-surface solid:
T - T0 = Q / (2 pi kappa R) * exp(-v (R + xi) / (2 alpha)),  xi = x - x_beam, R = |r - r_beam|.
BASELINE configs[1] (SURVEY.md 8d cfg 2) is this case; the reference does not ship it (finding F7)."""
import math

from additivefoam_b200 import cases

SPEED, POWER, ETA, KAPPA, RHO, CP = 0.4, 179.2, 0.33, 25.0, 7569.92, 600.0
N = (120, 50, 25)


def case():
    # dt = 4 us keeps implicit Euler's numerical diffusion v^2 dt / 2 at 6 % of alpha
    return cases.rosenthal(n=N, speed=SPEED, kappa=KAPPA, Cp=CP, eta=ETA, power=POWER, tolerance=1e-12, maxIter=100, adjustTimeStep=0,
                           deltaT=4e-6, endTime=0.0038)


def compare(T, time_now, mesh):
    """[(R, relative deviation of T - T0)] along the track behind the beam, at three depths."""
    T = T.reshape(N[2], N[1], N[0])
    dx = 20e-6
    xb = SPEED * time_now
    Q, alpha = ETA * POWER, KAPPA / (RHO * CP)
    out = []
    for xi in (-300e-6, -400e-6, -600e-6, -900e-6):
        for depth in (10e-6, 110e-6, 210e-6):
            i = int(round((xb + xi - mesh["min"][0]) / dx - 0.5))
            j = N[1] // 2
            k = N[2] - 1 - int(round(depth / dx - 0.5))
            xc, yc, zc = (mesh["min"][q] + (idx + 0.5) * dx for q, idx in enumerate((i, j, k)))
            xic = xc - xb
            R = math.sqrt(xic ** 2 + yc ** 2 + zc ** 2)
            Ta = Q / (2 * math.pi * KAPPA * R) * math.exp(-SPEED * (R + xic) / (2 * alpha))
            out.append((R, (T[k, j, i] - 300.0 - Ta) / Ta))
    return out


def check(dev):
    # a distributed (85 um Gaussian, 30 um deep) source and a 20 um mesh approach the point-source field from below with distance
    for R, d in dev:
        assert abs(d) < (0.06 if R > 550e-6 else 0.11), (R, d)
    far = [abs(d) for R, d in dev if R > 800e-6]
    near = [abs(d) for R, d in dev if R < 350e-6]
    assert max(far) < 0.045 and max(far) < max(near)
