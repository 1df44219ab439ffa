// Closed-form LDU addressing of one blockMesh hex block: for cell (i, row = j + ny*k)
// the ids of the internal faces it touches, in OpenFOAM's upper-triangular face order.
#pragma once
#include "ldu.cuh"

namespace b200 {

constexpr int BLK = 256;

struct RowCtx {
    int j, k, hy, hz, stride, fbase, fbaseBelow, fbaseBack;
};
__device__ __forceinline__ RowCtx row_ctx(const BlockDims& d, int row)
{
    RowCtx r;
    r.k = row / d.ny; r.j = row - r.k * d.ny;
    r.hy = r.j < d.ny - 1; r.hz = r.k < d.nz - 1;
    r.stride = 1 + r.hy + r.hz;
    r.fbase = d.rowFaceBase(r.j, r.k);
    r.fbaseBelow = r.k > 0 ? d.rowFaceBase(r.j, r.k - 1) : 0;   // row (j, k-1): stride 2 + hy
    r.fbaseBack = r.j > 0 ? d.rowFaceBase(r.j - 1, r.k) : 0;    // row (j-1, k): stride 2 + hz
    return r;
}
// face ids touching cell (i, row): lower-side z,y,x then own x,y,z (-1 when absent)
__device__ __forceinline__ void cell_faces(const BlockDims& d, const RowCtx& r, int i, int& fzl, int& fyl, int& fxl, int& fx, int& fy,
                                           int& fz)
{
    const int hx = i < d.nx - 1;
    const int f0 = r.fbase + i * r.stride;
    fzl = r.k > 0 ? r.fbaseBelow + i * (2 + r.hy) + hx + r.hy : -1;
    fyl = r.j > 0 ? r.fbaseBack + i * (2 + r.hz) + hx : -1;
    fxl = i > 0 ? f0 - r.stride : -1;
    fx = hx ? f0 : -1;
    fy = r.hy ? f0 + hx : -1;
    fz = r.hz ? f0 + hx + r.hy : -1;
}


} // namespace b200
