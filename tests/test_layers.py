"""SURVEY.md 8f N4, the layer workflow: applications/utilities/multiLayer/runLayers:133-205 on one uniform block
(additivefoam_b200/layers.py).  CPU: the mapping itself (mapNearest, the powder box, times, the scan-path wait).  GPU: two layers run
through the resident stepper against the oracle doing the same restarts."""
import numpy as np
import pytest

from additivefoam_b200 import abi, cases, layers


def _case(n=(12, 6, 8), cell=10e-6):
    return cases.multi_layer_pbf(n=n, cell=cell, hatch=20e-6, nHatch=2, endTime=2e-5, writeInterval=0.0)


def test_next_layer_maps_nearest_and_lays_powder():
    case = _case()
    nx, ny, nz = case["mesh"]["n"]
    rng = np.random.default_rng(3)
    T = rng.uniform(300.0, 2000.0, nx * ny * nz)
    P = (rng.uniform(0, 1, nx * ny * nz) > 0.5).astype(np.float64)
    new, T2, P2 = layers.next_layer(case, T, P, layer_thickness=40e-6, n_cells_per_layer=4, start_time=2e-5)
    assert new["mesh"]["n"] == [nx, ny, nz + 4]
    assert new["mesh"]["max"][2] == case["mesh"]["max"][2] and np.isclose(new["mesh"]["min"][2], case["mesh"]["min"][2] - 40e-6, rtol=1e-12)
    T3, T0 = T2.reshape(nz + 4, ny, nx), T.reshape(nz, ny, nx)
    assert np.array_equal(T3[:nz], T0)                                    # old cells: the coinciding source cell
    for k in range(nz, nz + 4):
        assert np.array_equal(T3[k], T0[-1])                              # new cells: the old top cell of the column
    P3, P0 = P2.reshape(nz + 4, ny, nx), P.reshape(nz, ny, nx)
    assert np.array_equal(P3[:nz], P0) and np.all(P3[nz:] == 1.0)         # the box z >= -layerThickness is exactly the new layers
    assert new["controls"]["startTime"] == 2e-5 and np.isclose(new["controls"]["endTime"], 4e-5)
    assert new["beams"][0]["path"][0][5] == 2e-5 and case["beams"][0]["path"][0][5] == 0.0      # the caller's case is not touched
    assert new["beams"][0]["path"][1:] == case["beams"][0]["path"][1:]


def test_next_layer_refuses_a_second_cell_height():
    case = _case()
    n = np.prod(case["mesh"]["n"])
    with pytest.raises(layers.LayerError):
        layers.next_layer(case, np.full(n, 300.0), np.zeros(n), layer_thickness=40e-6, n_cells_per_layer=3, start_time=0.0)
    with pytest.raises(layers.LayerError):
        layers.next_layer(case, np.full(n, 300.0), np.zeros(n), layer_thickness=40e-6, n_cells_per_layer=0, start_time=0.0)


@pytest.mark.gpu
def test_gpu_two_layers_against_the_oracle(ctx, oracle):
    """Two layers of the configs[4] family at toy size: layer 0 from the case, layer 1 from the mapped fields with the beam waiting
    until the layer's start time.  The oracle is restarted from the same mapped fields; every step of both layers within the bound."""
    import oracle_lib as ol
    from additivefoam_b200 import lib
    case = _case(n=(80, 16, 8), cell=20e-6)          # 1.6 mm long: the raster starts 0.5 mm inside
    seen = []

    def check_layer(cur, T0, P0, nsteps):
        g, o = lib.Case(ctx, cur), ol.OracleCase(cur, lib=oracle)
        if T0 is not None:
            g.restart(T0, P0); o.restart(T0, P0)
        rg = None
        for s in range(nsteps):
            if not g.running():
                break
            rg, ro = g.step(), o.step()
            assert rg.nThermoCorr == ro.nThermoCorr and abs(rg.lastSolveIterations - ro.lastSolveIterations) <= 1
            assert abs(rg.time - ro.time) <= 1e-15 * max(1.0, abs(ro.time))
            eT = float(np.max(np.abs(g.field(abi.FIELD_T) - o.field(abi.FIELD_T))) / np.max(np.abs(o.field(abi.FIELD_T))))
            eA = float(np.max(np.abs(g.field(abi.FIELD_ALPHA_SOLID) - o.field(abi.FIELD_ALPHA_SOLID))))
            assert eT < 1e-9 and eA < 1e-9, (s, eT, eA)
        out = (g.field(abi.FIELD_T), g.field(abi.FIELD_ALPHA_POWDER), rg.time, rg.totalPower)
        g.close(); o.close()
        return out

    T, P, t_end, _ = check_layer(case, None, None, 60)
    assert T.max() > 2000.0
    seen.append(t_end)
    new, T1, P1 = layers.next_layer(case, T, P, layer_thickness=40e-6, n_cells_per_layer=2, start_time=t_end, layer_time=2e-5)
    T2, P2, t2, power = check_layer(new, T1, P1, 60)
    assert t2 > t_end and power > 0.0                    # the beam is on again in the second layer: its path waited for the start time
    assert T2.size == T.size + 2 * 80 * 16


@pytest.mark.gpu
def test_gpu_run_layers_driver(ctx):
    """layers.run_layers: the loop of runLayers:152-205 -- every layer to its end time on the device, the callback once per layer."""
    case = _case(n=(80, 16, 8), cell=20e-6)
    seen = []
    last, T, P = layers.run_layers(ctx, case, 3, 40e-6, 2, on_layer_end=lambda layer, c, T_, P_: seen.append((layer, c["mesh"]["n"][2], T_.size, float(T_.max()))))
    assert [s[:3] for s in seen] == [(0, 8, 80 * 16 * 8), (1, 10, 80 * 16 * 10), (2, 12, 80 * 16 * 12)]
    assert all(s[3] > 1000.0 for s in seen)                                        # the beam came on in every layer
    assert last["mesh"]["n"][2] == 12 and np.isclose(last["controls"]["startTime"], 4e-5) and np.isclose(last["controls"]["endTime"], 6e-5)
    assert last["beams"][0]["path"][0][5] == last["controls"]["startTime"]
    assert T.size == P.size == 80 * 16 * 12
    assert P.reshape(12, 16, 80)[:8].max() <= 1.0 and P.reshape(12, 16, 80)[10:].max() == 1.0      # powder left beside the tracks of the last layer
