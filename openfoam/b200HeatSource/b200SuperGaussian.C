#include "b200SuperGaussian.H"
#include "addToRunTimeSelectionTable.H"
#include "Pstream.H"

namespace Foam
{
namespace heatSourceModels
{
    defineTypeNameAndDebug(b200SuperGaussian, 0);
    addToRunTimeSelectionTable(heatSourceModel, b200SuperGaussian, dictionary);
}
}

using Foam::constant::mathematical::pi;

static b200_ctx* b200HeatCtx()
{
    static b200_ctx* c = nullptr;
    if (!c && b200_ctx_create(0, nullptr, 0, 1, &c) != 0)
    {
        FatalErrorInFunction << b200_last_error() << Foam::exit(Foam::FatalError);
    }
    return c;
}


Foam::heatSourceModels::b200SuperGaussian::b200SuperGaussian
(
    const word& sourceName,
    const dictionary& dict,
    const fvMesh& mesh
)
:
    heatSourceModel(typeName, sourceName, dict, mesh),
    mesh_(mesh)
{
    k_ = heatSourceModelCoeffs_.lookup<scalar>("k");

    // uniform single block: counts from the number of distinct cell-centre coordinates
    const boundBox bb(mesh_.points(), false);
    const vectorField& C = mesh_.cellCentres();
    const scalar dx = 2.0*(C[0].x() - bb.min().x());
    const scalar dy = 2.0*(C[0].y() - bb.min().y());
    const scalar dz = 2.0*(C[0].z() - bb.min().z());
    block_.nx = label(round(bb.span().x()/dx));
    block_.ny = label(round(bb.span().y()/dy));
    block_.nz = label(round(bb.span().z()/dz));
    for (direction d = 0; d < 3; d++)
    {
        block_.min[d] = bb.min()[d];
        block_.max[d] = bb.max()[d];
    }
    if (label(block_.nx)*block_.ny*block_.nz != mesh_.nCells() || Pstream::parRun())
    {
        FatalErrorInFunction
            << "b200SuperGaussian needs a serial, single uniform blockMesh hex block"
            << exit(FatalError);
    }
}


Foam::scalar Foam::heatSourceModels::b200SuperGaussian::weight(const vector& d)
{
    vector s = dimensions_ / Foam::pow(2.0, 1.0/k_);
    scalar x = Foam::pow(magSqr(cmptDivide(d, s)), k_/2.0);
    return Foam::exp(-x);
}


Foam::dimensionedScalar Foam::heatSourceModels::b200SuperGaussian::V0()
{
    vector s = dimensions_ / Foam::pow(2.0, 1.0/k_);
    return dimensionedScalar
    (
        "V0", dimVolume, (2.0/3.0)*s.x()*s.y()*s.z()*pi*Foam::tgamma(1.0 + 3.0/k_)
    );
}


Foam::tmp<Foam::volScalarField> Foam::heatSourceModels::b200SuperGaussian::qDot()
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

    b200_beam beam;
    beam.heatSourceModel = B200_SUPER_GAUSSIAN;
    beam.k = k_; beam.m = 0; beam.A = 0; beam.B = 0;
    for (direction d = 0; d < 3; d++)
    {
        beam.dimensions[d] = dimensions_[d];
        beam.nPoints[d] = nPoints_[d];
    }
    beam.transient = transient_; beam.isoValue = isoValue_;
    // the absorbed fraction stays with the reference's absorptionModel: pass it as a constant
    const scalar aspectRatio = dimensions_.z()/min(dimensions_.x(), dimensions_.y());
    beam.absorptionModel = B200_ABSORPTION_CONSTANT;
    beam.eta = absorptionModel_->eta(aspectRatio);
    beam.kellyGeometry = 0; beam.eta0 = 0; beam.etaMin = 0;
    beam.deltaT = 0; beam.hitPathIntervals = 1; beam.nSegments = 0; beam.path = nullptr;

    const vector p = movingBeam_->position();
    const double pos[3] = {p.x(), p.y(), p.z()};
    const double dims[3] = {dimensions_.x(), dimensions_.y(), dimensions_.z()};
    double sumWeights, volume, absorbed;
    if
    (
        b200_beam_qdot
        (
            b200HeatCtx(), &block_, &beam, pos, dims, movingBeam_->power(),
            q.primitiveFieldRef().begin(), &sumWeights, &volume, &absorbed
        ) != 0
    )
    {
        FatalErrorInFunction << b200_last_error() << exit(FatalError);
    }
    return tqDot;
}


bool Foam::heatSourceModels::b200SuperGaussian::read()
{
    if (heatSourceModel::read())
    {
        heatSourceModelCoeffs_ = optionalSubDict(type() + "Coeffs");
        heatSourceModelCoeffs_.lookup("k") >> k_;
        return true;
    }
    return false;
}
