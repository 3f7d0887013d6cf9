//! TEST INFRASTRUCTURE (oracle): a user problem file for the UNMODIFIED reference, written against
//! its public pgen API (Mesh::InitUserMeshData / MeshBlock::ProblemGenerator, EnrollUserExplicit
//! SourceFunction, mesh/mesh.hpp:400-430).  The reference ships no Cartesian problem whose user
//! source term runs without a shearing box or self-gravity, so parity of that hook (task SRC_TERM,
//! hydro/srcterms/hydro_srcterms.cpp:151-154) is pinned with this one: a smooth density / pressure
//! bump in a periodic box, driven by a time-dependent body force along x3 and a cooling term.
//! The same problem is built into the B200 host (athena_b200/csrc/host/pgen.cpp, "usrsrc").
#include <cmath>

#include "../athena.hpp"
#include "../athena_arrays.hpp"
#include "../coordinates/coordinates.hpp"
#include "../eos/eos.hpp"
#include "../field/field.hpp"
#include "../hydro/hydro.hpp"
#include "../mesh/mesh.hpp"
#include "../parameter_input.hpp"

namespace {
Real g0, tcool, lz, zmin;
}

void DrivenCooling(MeshBlock *pmb, const Real time, const Real dt, const AthenaArray<Real> &prim,
                   const AthenaArray<Real> &prim_scalar, const AthenaArray<Real> &bcc,
                   AthenaArray<Real> &cons, AthenaArray<Real> &cons_scalar) {
  for (int k = pmb->ks; k <= pmb->ke; ++k) {
    for (int j = pmb->js; j <= pmb->je; ++j) {
      for (int i = pmb->is; i <= pmb->ie; ++i) {
        const Real den = prim(IDN,k,j,i);
        const Real x3 = pmb->pcoord->x3v(k);
        const Real acc = g0*(1.0 + time)*std::sin(TWO_PI*(x3 - zmin)/lz);
        cons(IM3,k,j,i) -= dt*den*acc;
        cons(IEN,k,j,i) -= dt*den*acc*prim(IVZ,k,j,i);
        cons(IEN,k,j,i) -= dt*prim(IPR,k,j,i)/tcool;
      }
    }
  }
}

void Mesh::InitUserMeshData(ParameterInput *pin) {
  g0 = pin->GetReal("problem", "g0");
  tcool = pin->GetReal("problem", "tcool");
  lz = mesh_size.x3max - mesh_size.x3min;
  zmin = mesh_size.x3min;
  EnrollUserExplicitSourceFunction(DrivenCooling);
}

void MeshBlock::ProblemGenerator(ParameterInput *pin) {
  const Real gm1 = peos->GetGamma() - 1.0;
  const Real amp = pin->GetReal("problem", "amp");
  const Real b0 = MAGNETIC_FIELDS_ENABLED ? pin->GetReal("problem", "b0") : 0.0;
  RegionSize &ms = pmy_mesh->mesh_size;
  for (int k = ks; k <= ke; ++k) {
    for (int j = js; j <= je; ++j) {
      for (int i = is; i <= ie; ++i) {
        const Real s1 = std::sin(TWO_PI*(pcoord->x1v(i) - ms.x1min)/(ms.x1max - ms.x1min));
        const Real s2 = std::cos(TWO_PI*(pcoord->x2v(j) - ms.x2min)/(ms.x2max - ms.x2min));
        const Real s3 = std::sin(TWO_PI*(pcoord->x3v(k) - ms.x3min)/(ms.x3max - ms.x3min));
        const Real d = 1.0 + amp*s1*s2;
        const Real p = 1.0 + amp*s2*s3;
        phydro->u(IDN,k,j,i) = d;
        phydro->u(IM1,k,j,i) = 0.1*d*s3;
        phydro->u(IM2,k,j,i) = 0.0;
        phydro->u(IM3,k,j,i) = -0.05*d*s1;
        phydro->u(IEN,k,j,i) = p/gm1 + 0.5*(SQR(phydro->u(IM1,k,j,i)) + SQR(phydro->u(IM3,k,j,i)))/d;
      }
    }
  }
  if (MAGNETIC_FIELDS_ENABLED) {
    for (int k = ks; k <= ke; ++k)
      for (int j = js; j <= je; ++j)
        for (int i = is; i <= ie+1; ++i) pfield->b.x1f(k,j,i) = b0;
    for (int k = ks; k <= ke; ++k)
      for (int j = js; j <= je+1; ++j)
        for (int i = is; i <= ie; ++i) pfield->b.x2f(k,j,i) = 0.5*b0;
    for (int k = ks; k <= ke+1; ++k)
      for (int j = js; j <= je; ++j)
        for (int i = is; i <= ie; ++i) pfield->b.x3f(k,j,i) = 0.0;
    for (int k = ks; k <= ke; ++k)
      for (int j = js; j <= je; ++j)
        for (int i = is; i <= ie; ++i) phydro->u(IEN,k,j,i) += 0.5*(SQR(b0) + SQR(0.5*b0));
  }
}
