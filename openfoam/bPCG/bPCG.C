/*---------------------------------------------------------------------------*\
  See bPCG.H.  Every numerical step happens behind b200_ldu_solve(); this file only
  translates OpenFOAM's objects into the plain arrays of include/additivefoam_b200.h.
\*---------------------------------------------------------------------------*/
#include "bPCG.H"
#include "processorLduInterface.H"
#include "Pstream.H"
#include "addToRunTimeSelectionTable.H"
#include <map>

namespace Foam
{
    defineTypeNameAndDebug(bPCG, 0);
    lduMatrix::solver::addsymMatrixConstructorToTable<bPCG> addbPCGSymMatrixConstructorToTable_;
    lduMatrix::solver::addasymMatrixConstructorToTable<bPCG> addbPCGAsymMatrixConstructorToTable_;

    defineTypeNameAndDebug(bPBiCGStab, 0);
    lduMatrix::solver::addsymMatrixConstructorToTable<bPBiCGStab> addbPBiCGStabSymMatrixConstructorToTable_;
    lduMatrix::solver::addasymMatrixConstructorToTable<bPBiCGStab> addbPBiCGStabAsymMatrixConstructorToTable_;
}


static void b200Check(const int rc, const char* what)
{
    if (rc != 0)
    {
        FatalErrorInFunction
            << what << ": " << b200_last_error() << Foam::exit(Foam::FatalError);
    }
}




b200_ctx* Foam::b200SolverBase::ctx()
{
    static b200_ctx* c = nullptr;
    if (!c)
    {
        // one rank <-> one GPU of the node; ranks exchange the NCCL id through Pstream
        unsigned char id[128] = {0};
        if (Pstream::parRun())
        {
            if (Pstream::master())
            {
                b200Check(b200_nccl_unique_id(id), "b200_nccl_unique_id");
            }
            List<unsigned char> idList(128);
            forAll(idList, i) { idList[i] = id[i]; }
            Pstream::scatter(idList);
            forAll(idList, i) { id[i] = idList[i]; }
        }
        // the ranks of a node share its GPUs round-robin (one rank per GPU when there are as many)
        int nDevices = 1;
        b200Check(b200_device_count(&nDevices), "b200_device_count");
        b200Check
        (
            b200_ctx_create
            (
                Pstream::parRun() ? Pstream::myProcNo() % nDevices : 0,
                Pstream::parRun() ? id : nullptr,
                Pstream::parRun() ? Pstream::myProcNo() : 0,
                Pstream::parRun() ? Pstream::nProcs() : 1,
                &c
            ),
            "b200_ctx_create"
        );
    }
    return c;
}


b200_ldu* Foam::b200SolverBase::handle() const
{
    // addressing is cached per lduAddressing object; coefficients are fresh per solve.  The entry is validated before reuse
    // (sizes and the address arrays themselves): a mesh that was destroyed or changed topology can leave another lduAddressing
    // at the same address.
    struct Cached { b200_ldu* h; label nCells, nFaces; const label* lower; const label* upper; };
    static std::map<const lduAddressing*, Cached> cache;
    const lduAddressing& addr = matrix_.lduAddr();
    auto it = cache.find(&addr);
    if (it != cache.end())
    {
        const Cached& c = it->second;
        if
        (
            c.nCells == addr.size() && c.nFaces == addr.lowerAddr().size()
         && c.lower == addr.lowerAddr().begin() && c.upper == addr.upperAddr().begin()
        )
        {
            return c.h;
        }
        b200_ldu_destroy(c.h);
        cache.erase(it);
    }

    DynamicList<label> neighbRank, size;
    DynamicList<const label*> faceCells;
    forAll(interfaces_, patchi)
    {
        if (interfaces_.set(patchi))
        {
            const processorLduInterface& pli =
                refCast<const processorLduInterface>(interfaces_[patchi].interface());
            const labelUList& pa = addr.patchAddr(patchi);
            neighbRank.append(pli.neighbProcNo());
            size.append(pa.size());
            faceCells.append(pa.begin());
        }
    }

    b200_ldu* h = nullptr;
    b200Check
    (
        b200_ldu_create
        (
            ctx(),
            addr.size(),
            addr.lowerAddr().size(),
            addr.lowerAddr().begin(),
            addr.upperAddr().begin(),
            neighbRank.size(),
            neighbRank.begin(),
            size.begin(),
            faceCells.begin(),
            &h
        ),
        "b200_ldu_create"
    );

    // The addressing-free block kernels.  `blockNx/blockNy/blockNz` in the solver dictionary declare this rank's cells to be one
    // blockMesh hex block (an error if the addressing says otherwise).  Without them (`blockDetect yes`, the default) the shape is
    // read off the faces cell 0 owns -- in blockMesh cell order they lead to cells 1, nx and nx*ny -- and offered to the library,
    // which compares the closed form with the whole addressing bit for bit and keeps the general kernels if it differs.  That also
    // covers each rank of a `simple` / `hierarchical` decomposition of a block: its interior is a block, the processor patches stay
    // coupled interfaces.
    label bnx = controlDict_.lookupOrDefault<label>("blockNx", 0);
    label bny = controlDict_.lookupOrDefault<label>("blockNy", 0);
    label bnz = controlDict_.lookupOrDefault<label>("blockNz", 0);
    const bool declared = bnx > 0 && bny > 0 && bnz > 0;
    if (!declared && controlDict_.lookupOrDefault<Switch>("blockDetect", true) && addr.size() > 0)
    {
        const labelUList& l = addr.lowerAddr();
        const labelUList& u = addr.upperAddr();
        label owned = 0;
        while (owned < l.size() && owned < 3 && l[owned] == 0) owned++;
        const label nCells = addr.size();
        bnx = owned >= 2 ? u[1] : (owned == 1 ? nCells : 1);
        const label nxny = owned >= 3 ? u[2] : nCells;
        if (bnx > 0 && nxny % bnx == 0 && nCells % nxny == 0)
        {
            bny = nxny/bnx;
            bnz = nCells/nxny;
        }
        else
        {
            bnx = bny = bnz = 0;
        }
    }
    if (bnx > 0 && bny > 0 && bnz > 0)
    {
        const int rc = b200_ldu_set_block(h, bnx, bny, bnz);
        if (rc != 0 && declared)
        {
            b200Check(rc, "b200_ldu_set_block");
        }
    }
    cache[&addr] = Cached{h, addr.size(), addr.lowerAddr().size(), addr.lowerAddr().begin(), addr.upperAddr().begin()};
    return h;
}


Foam::b200SolverBase::b200SolverBase
(
    const word& fieldName,
    const lduMatrix& matrix,
    const FieldField<Field, scalar>& interfaceBouCoeffs,
    const FieldField<Field, scalar>& interfaceIntCoeffs,
    const lduInterfaceFieldPtrsList& interfaces,
    const dictionary& solverControls,
    const int solverType
)
:
    lduMatrix::solver
    (
        fieldName,
        matrix,
        interfaceBouCoeffs,
        interfaceIntCoeffs,
        interfaces,
        solverControls
    ),
    solverType_(solverType),
    preconditioner_(B200_PRECOND_NONE)
{
    const word pName(controlDict_.lookupOrDefault<word>("preconditioner", "none"));
    if (pName == "DIC") preconditioner_ = B200_PRECOND_DIC;
    else if (pName == "DILU") preconditioner_ = B200_PRECOND_DILU;
    else if (pName == "diagonal") preconditioner_ = B200_PRECOND_DIAGONAL;
    else if (pName == "none") preconditioner_ = B200_PRECOND_NONE;
    else
    {
        FatalIOErrorInFunction(controlDict_)
            << "Unknown preconditioner " << pName
            << " (DIC, DILU, diagonal, none)" << exit(FatalIOError);
    }
}


Foam::solverPerformance Foam::b200SolverBase::solveB200
(
    const word& solverName,
    scalarField& psi,
    const scalarField& source,
    const bool forceSymmetric
) const
{
    const scalar* lowerPtr = nullptr;
    if (matrix_.asymmetric() && !forceSymmetric)
    {
        lowerPtr = matrix_.lower().begin();
    }

    DynamicList<const scalar*> bouCoeffs;
    forAll(interfaces_, patchi)
    {
        if (interfaces_.set(patchi))
        {
            bouCoeffs.append(interfaceBouCoeffs_[patchi].begin());
        }
    }

    b200_solver_perf perf;
    b200Check
    (
        b200_ldu_solve
        (
            handle(),
            solverType_,
            preconditioner_,
            matrix_.diag().begin(),
            matrix_.upper().begin(),
            lowerPtr,
            bouCoeffs.size() ? bouCoeffs.begin() : nullptr,
            source.begin(),
            psi.begin(),
            tolerance_,
            relTol_,
            minIter_,
            maxIter_,
            &perf
        ),
        "b200_ldu_solve"
    );

    solverPerformance solverPerf(solverName, fieldName_);
    solverPerf.initialResidual() = perf.initialResidual;
    solverPerf.finalResidual() = perf.finalResidual;
    solverPerf.nIterations() = perf.nIterations;
    // converged_/singular_ are recomputed by checkConvergence/checkSingularity in upstream's
    // solverPerformance; the printed line ("DICbPCG: Solving for T, ...") uses the three above
    solverPerf.checkConvergence(tolerance_, relTol_);
    return solverPerf;
}


Foam::bPCG::bPCG
(
    const word& fieldName,
    const lduMatrix& matrix,
    const FieldField<Field, scalar>& interfaceBouCoeffs,
    const FieldField<Field, scalar>& interfaceIntCoeffs,
    const lduInterfaceFieldPtrsList& interfaces,
    const dictionary& solverControls
)
:
    b200SolverBase
    (
        fieldName, matrix, interfaceBouCoeffs, interfaceIntCoeffs, interfaces,
        solverControls, B200_SOLVER_PCG
    )
{}


Foam::solverPerformance Foam::bPCG::solve
(
    scalarField& psi,
    const scalarField& source,
    const direction
) const
{
    // PCG needs lower == upper.  With phi == 0 (nOuterCorrectors 0, finding F6) the TEqn
    // coefficients are numerically symmetric although both triangles are allocated.
    if (matrix_.asymmetric())
    {
        const scalarField& lower = matrix_.lower();
        const scalarField& upper = matrix_.upper();
        forAll(lower, facei)
        {
            if (lower[facei] != upper[facei])
            {
                FatalErrorInFunction
                    << "bPCG selected for field " << fieldName_
                    << " but the matrix is not symmetric (face " << facei
                    << "); use bPBiCGStab" << exit(FatalError);
            }
        }
    }
    return solveB200(lduMatrix::preconditioner::getName(controlDict_) + typeName, psi, source, true);
}


Foam::bPBiCGStab::bPBiCGStab
(
    const word& fieldName,
    const lduMatrix& matrix,
    const FieldField<Field, scalar>& interfaceBouCoeffs,
    const FieldField<Field, scalar>& interfaceIntCoeffs,
    const lduInterfaceFieldPtrsList& interfaces,
    const dictionary& solverControls
)
:
    b200SolverBase
    (
        fieldName, matrix, interfaceBouCoeffs, interfaceIntCoeffs, interfaces,
        solverControls, B200_SOLVER_PBICGSTAB
    )
{}


Foam::solverPerformance Foam::bPBiCGStab::solve
(
    scalarField& psi,
    const scalarField& source,
    const direction
) const
{
    return solveB200(lduMatrix::preconditioner::getName(controlDict_) + typeName, psi, source, false);
}
