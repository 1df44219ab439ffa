#include "b200HeatSources.H"
#include "addToRunTimeSelectionTable.H"
#include "Pstream.H"

namespace Foam
{
namespace heatSourceModels
{
    defineTypeNameAndDebug(b200SuperGaussian, 0);
    addToRunTimeSelectionTable(heatSourceModel, b200SuperGaussian, dictionary);
    defineTypeNameAndDebug(b200ModifiedSuperGaussian, 0);
    addToRunTimeSelectionTable(heatSourceModel, b200ModifiedSuperGaussian, dictionary);
    defineTypeNameAndDebug(b200ProjectedGaussian, 0);
    addToRunTimeSelectionTable(heatSourceModel, b200ProjectedGaussian, dictionary);
}
}


static b200_ctx* b200HeatCtx()
{
    // one context per process; in a parallel run the ranks of a node share its GPUs round-robin.  No communicator is needed:
    // the only global quantity, the normalisation sum of heatSourceModel.C:359, goes through Foam::reduce (b200ReduceSum)
    static b200_ctx* c = nullptr;
    if (!c)
    {
        int nDevices = 1;
        if (b200_device_count(&nDevices) != 0 || nDevices < 1)
        {
            FatalErrorInFunction << b200_last_error() << Foam::exit(Foam::FatalError);
        }
        const int device = Foam::Pstream::parRun() ? Foam::Pstream::myProcNo() % nDevices : 0;
        if (b200_ctx_create(device, nullptr, 0, 1, &c) != 0)
        {
            FatalErrorInFunction << b200_last_error() << Foam::exit(Foam::FatalError);
        }
    }
    return c;
}


// fvc::domainIntegrate's reduction (heatSourceModel.C:359) for b200_beam_qdot_sub
static double b200ReduceSum(double local, void*)
{
    Foam::scalar s = local;
    Foam::reduce(s, Foam::sumOp<Foam::scalar>());
    return s;
}


// * * * * * * * * * * * * * * * * common part * * * * * * * * * * * * * * * //

Foam::heatSourceModels::b200HeatSource::b200HeatSource
(
    const word& type,
    const label shape,
    const word& sourceName,
    const dictionary& dict,
    const fvMesh& mesh
)
:
    heatSourceModel(type, sourceName, dict, mesh)
{
    beam_ = b200_beam{};
    beam_.heatSourceModel = shape;
    beam_.absorptionModel = B200_ABSORPTION_CONSTANT;
    beam_.hitPathIntervals = 1;

    // This rank's cells must be a sub-block of ONE uniform blockMesh hex block, in blockMesh cell order (x fastest): the whole
    // block in a serial run, one piece of a `simple` / `hierarchical` decomposition in a parallel one (decomposePar keeps the
    // cell order within a rank).  Spacing from the first cell centre, counts from the bounding boxes.
    const boundBox bbGlobal(mesh.points(), true);       // reduced over the ranks
    const boundBox bbLocal(mesh.points(), false);
    const vector c0 = mesh.cellCentres()[0];
    vector spacing;
    for (direction d = 0; d < 3; d++)
    {
        spacing[d] = 2.0*(c0[d] - bbLocal.min()[d]);
        block_.min[d] = bbGlobal.min()[d];
        block_.max[d] = bbGlobal.max()[d];
        nLocal_[d] = int32_t(round(bbLocal.span()[d]/spacing[d]));
        offset_[d] = int32_t(round((bbLocal.min()[d] - bbGlobal.min()[d])/spacing[d]));
    }
    block_.nx = int32_t(round(bbGlobal.span()[0]/spacing[0]));
    block_.ny = int32_t(round(bbGlobal.span()[1]/spacing[1]));
    block_.nz = int32_t(round(bbGlobal.span()[2]/spacing[2]));

    bool ok = label(nLocal_[0])*nLocal_[1]*nLocal_[2] == mesh.nCells();
    if (ok)
    {
        // the last cell and one in the middle sit where the closed form i + nx*(j + ny*k) puts them
        const label probes[2] = {mesh.nCells() - 1, mesh.nCells()/2};
        for (int q = 0; q < 2; q++)
        {
            const label c = probes[q];
            const label ijk[3] = {c % nLocal_[0], (c/nLocal_[0]) % nLocal_[1], c/(nLocal_[0]*nLocal_[1])};
            for (direction d = 0; d < 3; d++)
            {
                const scalar expected = bbLocal.min()[d] + (ijk[d] + 0.5)*spacing[d];
                ok = ok && mag(mesh.cellCentres()[c][d] - expected) < 1e-6*spacing[d];
            }
        }
    }
    reduce(ok, andOp<bool>());
    if (!ok)
    {
        FatalErrorInFunction
            << type << " needs a mesh that is one uniform blockMesh hex block, undecomposed or cut by the"
            << " simple / hierarchical method (each rank a sub-block in blockMesh cell order)"
            << exit(FatalError);
    }
}


const b200_beam& Foam::heatSourceModels::b200HeatSource::currentBeam()
{
    // staticDimensions_ in the record (projectedGaussian derives its exponent from them, projectedGaussian.C:99-103);
    // the current, possibly transient, dimensions_ travel as a separate argument
    for (direction d = 0; d < 3; d++)
    {
        beam_.dimensions[d] = staticDimensions_[d];
        beam_.nPoints[d] = nPoints_[d];
    }
    beam_.transient = transient_;
    beam_.isoValue = isoValue_;
    // heatSourceModel.C:246-256: eta of the reference's own absorption model, handed over as a constant
    const scalar aspectRatio = dimensions_.z()/min(dimensions_.x(), dimensions_.y());
    beam_.eta = absorptionModel_->eta(aspectRatio);
    return beam_;
}


Foam::scalar Foam::heatSourceModels::b200HeatSource::weight(const vector& d)
{
    const double dims[3] = {dimensions_.x(), dimensions_.y(), dimensions_.z()};
    const double dist[3] = {d.x(), d.y(), d.z()};
    double w = 0;
    if (b200_beam_shape_weight(&currentBeam(), dims, dist, &w) != 0)
    {
        FatalErrorInFunction << b200_last_error() << exit(FatalError);
    }
    return w;
}


Foam::dimensionedScalar Foam::heatSourceModels::b200HeatSource::V0()
{
    const double dims[3] = {dimensions_.x(), dimensions_.y(), dimensions_.z()};
    double v = 0;
    if (b200_beam_shape_v0(&currentBeam(), dims, &v, nullptr) != 0)
    {
        FatalErrorInFunction << b200_last_error() << exit(FatalError);
    }
    return dimensionedScalar("V0", dimVolume, v);
}


Foam::tmp<Foam::volScalarField> Foam::heatSourceModels::b200HeatSource::qDot()
{
    tmp<volScalarField> tqDot
    (
        new volScalarField
        (
            IOobject("qDot_", mesh_.time().timeName(), mesh_, IOobject::NO_READ, IOobject::NO_WRITE),
            mesh_,
            dimensionedScalar("Zero", dimPower/dimVolume, 0.0)
        )
    );
    volScalarField& q = tqDot.ref();

    const vector p = movingBeam_->position();
    const double pos[3] = {p.x(), p.y(), p.z()};
    const double dims[3] = {dimensions_.x(), dimensions_.y(), dimensions_.z()};
    double sumWeights, volume, absorbed;
    if
    (
        b200_beam_qdot_sub
        (
            b200HeatCtx(), &block_, offset_, nLocal_, &currentBeam(), pos, dims, movingBeam_->power(),
            Pstream::parRun() ? &b200ReduceSum : nullptr, nullptr,
            q.primitiveFieldRef().begin(), &sumWeights, &volume, &absorbed
        ) != 0
    )
    {
        FatalErrorInFunction << b200_last_error() << exit(FatalError);
    }
    return tqDot;
}


bool Foam::heatSourceModels::b200HeatSource::read()
{
    if (heatSourceModel::read())
    {
        heatSourceModelCoeffs_ = optionalSubDict(type() + "Coeffs");
        readCoeffs();
        return true;
    }
    return false;
}


// * * * * * * * * * * * * * * * * the three shapes * * * * * * * * * * * * * //

#define B200_DEFINE_HEAT_SOURCE(Type, SHAPE)                                                           \
Foam::heatSourceModels::Type::Type(const word& sourceName, const dictionary& dict, const fvMesh& mesh) \
:                                                                                                      \
    b200HeatSource(typeName, SHAPE, sourceName, dict, mesh)                                            \
{                                                                                                      \
    readCoeffs();                                                                                      \
}

B200_DEFINE_HEAT_SOURCE(b200SuperGaussian, B200_SUPER_GAUSSIAN)
B200_DEFINE_HEAT_SOURCE(b200ModifiedSuperGaussian, B200_MODIFIED_SUPER_GAUSSIAN)
B200_DEFINE_HEAT_SOURCE(b200ProjectedGaussian, B200_PROJECTED_GAUSSIAN)


void Foam::heatSourceModels::b200SuperGaussian::readCoeffs()
{
    beam_.k = heatSourceModelCoeffs_.lookup<scalar>("k");
}


void Foam::heatSourceModels::b200ModifiedSuperGaussian::readCoeffs()
{
    beam_.k = heatSourceModelCoeffs_.lookup<scalar>("k");
    beam_.m = heatSourceModelCoeffs_.lookup<scalar>("m");
}


void Foam::heatSourceModels::b200ProjectedGaussian::readCoeffs()
{
    beam_.A = heatSourceModelCoeffs_.lookup<scalar>("A");
    beam_.B = heatSourceModelCoeffs_.lookup<scalar>("B");
}
