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
        b200Check
        (
            b200_ctx_create
            (
                Pstream::parRun() ? Pstream::myProcNo() % 8 : 0,
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
    // addressing is cached per lduAddressing object; coefficients are fresh per solve
    static std::map<const lduAddressing*, b200_ldu*> cache;
    const lduAddressing& addr = matrix_.lduAddr();
    auto it = cache.find(&addr);
    if (it != cache.end())
    {
        return it->second;
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
    // optional: `blockNx/blockNy/blockNz` in the solver dictionary declares a single blockMesh hex block (serial runs); the
    // library verifies the addressing bit-exactly and switches to its addressing-free kernels
    const label bnx = controlDict_.lookupOrDefault<label>("blockNx", 0);
    const label bny = controlDict_.lookupOrDefault<label>("blockNy", 0);
    const label bnz = controlDict_.lookupOrDefault<label>("blockNz", 0);
    if (bnx > 0 && bny > 0 && bnz > 0 && !Pstream::parRun())
    {
        b200Check(b200_ldu_set_block(h, bnx, bny, bnz), "b200_ldu_set_block");
    }
    cache[&addr] = h;
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
