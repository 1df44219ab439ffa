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
    static b200_ctx* c = nullptr;
    if (!c && b200_ctx_create(0, nullptr, 0, 1, &c) != 0)
    {
        FatalErrorInFunction << b200_last_error() << Foam::exit(Foam::FatalError);
    }
    return c;
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

    // a uniform single block: spacing from the first cell centre, counts from the bounding box
    const boundBox bb(mesh.points(), false);
    const vector c0 = mesh.cellCentres()[0];
    label n[3];
    for (direction d = 0; d < 3; d++)
    {
        const scalar spacing = 2.0*(c0[d] - bb.min()[d]);
        n[d] = label(round(bb.span()[d]/spacing));
        block_.min[d] = bb.min()[d];
        block_.max[d] = bb.max()[d];
    }
    block_.nx = n[0]; block_.ny = n[1]; block_.nz = n[2];

    if (n[0]*n[1]*n[2] != mesh.nCells() || Pstream::parRun())
    {
        FatalErrorInFunction
            << type << " needs a serial run on a single uniform blockMesh hex block"
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
        b200_beam_qdot
        (
            b200HeatCtx(), &block_, &currentBeam(), pos, dims, movingBeam_->power(),
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
