// Driver that EXECUTES the solver shim of the drop-in seam B1, openfoam/bPCG/bPCG.{H,C}: the shim is compiled against the executable
// stand-in of oracle/ref_stubs/ldu (a real lduMatrix / lduAddressing, [UPSTREAM] lduMatrix::solver::New dispatch through run-time
// selection tables that the shim's own add...ConstructorToTable objects fill) and linked against the product library.  Request on
// stdin:   bpcg fieldName nCells nFaces hasLower nEntries (key value)...   lowerAddr upperAddr diag upper [lower] source psi0
// -> "solverName nIterations initialResidual finalResidual converged" then psi (hex floats).
// Built by oracle/Makefile into oracle/_ref/refBPCG; runs on the GPU box only (tests/test_gpu_shim_exec.py).  TEST INFRASTRUCTURE.
#include "bPCG.C"

#include <cstdio>

int main()
{
    using namespace Foam;
    std::string what, fieldName;
    if (!(std::cin >> what >> fieldName) || what != "bpcg") return 2;
    lduMatrix M;
    int n, nf, hasLower, nEntries;
    std::cin >> n >> nf >> hasLower >> nEntries;
    dictionary dict;
    for (int i = 0; i < nEntries; i++) { std::string k, v; std::cin >> k >> v; dict.e_[k] = v; }
    M.addr_.nCells_ = n;
    auto readL = [&](labelUList& l, int cnt) { for (int i = 0; i < cnt; i++) { label v; std::cin >> v; l.append(v); } };
    auto readS = [&](scalarField& f, int cnt) { for (int i = 0; i < cnt; i++) { scalar v; std::cin >> v; f.append(v); } };
    readL(M.addr_.lower_, nf); readL(M.addr_.upper_, nf);
    readS(M.diag_, n); readS(M.upper_, nf);
    if (hasLower) readS(M.lower_, nf);
    scalarField source, psi;
    readS(source, n); readS(psi, n);
    if (!std::cin) return 3;
    FieldField<Field, scalar> bou, in;
    lduInterfaceFieldPtrsList interfaces;
    try {
        solverPerformance perf = lduMatrix::solver::New(fieldName, M, bou, in, interfaces, dict)->solve(psi, source);
        std::printf("%s %d %a %a %d\n", perf.solverName().c_str(), perf.nIterations(), perf.initialResidual(), perf.finalResidual(), perf.converged() ? 1 : 0);
        for (int i = 0; i < n; i++) std::printf("%a ", psi[i]);
        std::printf("\n");
    } catch (const std::exception& e) { std::printf("ERROR %s\n", e.what()); return 4; }
    return 0;
}
