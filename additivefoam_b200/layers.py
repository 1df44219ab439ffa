"""The reference's multi-layer workflow (applications/utilities/multiLayer/runLayers) for the resident stepper.

`runLayers` runs one additiveFoam case per deposited layer.  Between two layers it (:133-205)

  * extrudes the mesh upwards from the `top` patch by `nCellsPerLayer` cell layers of total height `layerThickness`
    (system/extrudeMeshDict: linearDirection (0 0 1), :140-150),
  * moves the previous end time into the new case as its start time and adds the layer time to the end time (:166-171),
  * rewrites the parameter of the first scan-path row to that start time (:173-177: the beam waits until the layer starts),
  * maps the previous layer's fields onto the new mesh with `mapFields -mapMethod mapNearest` (:180),
  * sets alpha.powder = 1 in the box z >= -layerThickness of the mesh translated so that its top is z = 0
    (:183-197 setFieldsDict, :201 transformPoints), and runs the solver on that translated mesh.

On one blockMesh block with uniform cells this is closed-form: the block grows by `nCellsPerLayer` layers (which must have the
cell height of the block -- the block kernels need one dz), every old cell keeps its values (its centre coincides with a source
centre), every new cell takes the values of the old top cell of its column (the nearest source centre), the powder box is the new
layers.  alpha.solid is never mapped: the reference does not write it, a run started from fields rebuilds it from T
(createFields.H:31-44,74-78; `lib.Case.restart`).  [UPSTREAM] extrudeMesh appends the new cells after the old ones and renumbers
faces; here the grown block is numbered like a blockMesh block, which changes the order of the matrix rows, not the equations.
"""
import copy

import numpy as np


class LayerError(ValueError):
    pass


def next_layer(case, T, alpha_powder, layer_thickness, n_cells_per_layer, start_time, layer_time=None):
    """(case', T', alpha.powder') of the layer that follows `case`, whose run ended at `start_time` with fields T, alpha.powder.

    `case` is a description as `cases.py` / `foamcase.read_case` build it; its mesh has its top at max[2] (z = 0 in the solver's
    frame, as after `transformPoints`).  `layer_time` defaults to the duration of `case`'s own run (runLayers:156)."""
    m = case["mesh"]
    nx, ny, nz = m["n"]
    dz = (m["max"][2] - m["min"][2]) / nz
    if n_cells_per_layer < 1 or layer_thickness <= 0:
        raise LayerError("layerThickness and nCellsPerLayer must be positive")
    if abs(layer_thickness / n_cells_per_layer - dz) > 1e-9 * dz:
        raise LayerError(f"the deposited cells ({layer_thickness / n_cells_per_layer:g} m high) must have the cell height of the block "
                         f"({dz:g} m): the block kernels address one uniform block")
    T = np.asarray(T, dtype=np.float64).reshape(nz, ny, nx)
    P = np.asarray(alpha_powder, dtype=np.float64).reshape(nz, ny, nx)
    add = int(n_cells_per_layer)
    # mapNearest: old cells onto themselves, new cells from the old top cell of their column
    T2 = np.concatenate([T, np.repeat(T[-1:], add, axis=0)], axis=0)
    P2 = np.concatenate([P, np.repeat(P[-1:], add, axis=0)], axis=0)
    # setFields boxToCell (-1 -1 -layerThickness) (1 1 1) on the mesh whose top is z = 0: cells whose CENTRE lies in the box
    nz2 = nz + add
    zc = m["max"][2] - (nz2 - np.arange(nz2) - 0.5) * dz           # centres in the new solver frame (top at max[2])
    P2[(zc - m["max"][2]) >= -layer_thickness] = 1.0
    new = copy.deepcopy(case)
    new["mesh"]["n"] = [nx, ny, nz2]
    new["mesh"]["min"] = [m["min"][0], m["min"][1], m["min"][2] - layer_thickness]     # the part sinks, the top stays at max[2]
    c = new["controls"]
    duration = layer_time if layer_time is not None else case["controls"]["endTime"] - case["controls"]["startTime"]
    c["startTime"] = float(start_time)
    c["endTime"] = float(start_time) + float(duration)
    for b in new["beams"]:
        if b["path"]:
            b["path"][0] = list(b["path"][0])
            b["path"][0][5] = float(start_time)                      # runLayers:173-177
    new["powderZMin"] = 1e300                                        # the powder field comes from the caller from now on
    return new, np.ascontiguousarray(T2.reshape(-1)), np.ascontiguousarray(P2.reshape(-1))


def run_layers(ctx, case, n_layers, layer_thickness, n_cells_per_layer, on_layer_end=None, max_steps=None):
    """runLayers:152-205 on one context: every layer runs to its end time on the device, the fields come to the host once per
    layer for the mapping.  Returns the last (case, T, alpha.powder).  `on_layer_end(layer, case, T, alpha_powder)` is the place
    for `foamcase.write_time_directory`."""
    from . import abi, lib
    cur, T, P = case, None, None
    for layer in range(n_layers):
        g = lib.Case(ctx, cur)
        if layer > 0:
            g.restart(T, P)
        steps, end = 0, cur["controls"]["startTime"]
        while g.running() and (max_steps is None or steps < max_steps):
            end = g.step().time
            steps += 1
        if not g.running():
            end = cur["controls"]["endTime"]                         # runLayers:160 takes the previous case's nominal end time
        T, P = g.field(abi.FIELD_T), g.field(abi.FIELD_ALPHA_POWDER)
        g.close()
        if on_layer_end is not None:
            on_layer_end(layer, cur, T, P)
        if layer + 1 < n_layers:
            cur, T, P = next_layer(cur, T, P, layer_thickness, n_cells_per_layer, end)
    return cur, T, P
