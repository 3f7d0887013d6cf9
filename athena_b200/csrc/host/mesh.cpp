// mesh.cpp -- Mesh / MeshBlock / Coordinates of the host side.
// Follows src/mesh/mesh.cpp:72-567 (constructor: <mesh>/<meshblock> parsing, root level, Z-ordered
// block list), :1622-1967 (Initialize), :1236-1282 (NewTimeStep), :1984-2060
// (SetBlockSizeAndBoundaries), src/mesh/meshblock.cpp:62-261, src/coordinates/coordinates.cpp:
// 100-324 + cartesian.cpp:24-60, mesh.hpp:457-473,535-560 (uniform mesh generator).
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <limits>

#include "athena.hpp"

Config g_config;

void AthkCheck(int rc, const char *where) {
  if (rc != 0) {
    std::stringstream msg;
    msg << "### FATAL ERROR in " << where << std::endl << "libathk: " << athk_last_error();
    ATHENA_ERROR(msg);
  }
}

// mesh.hpp:457-473 (sym_interval = true) and :535-538
static Real ComputeMeshGeneratorX(std::int64_t index, std::int64_t nrange) {
  const std::int64_t noffset = index - (nrange) / 2;
  const std::int64_t noffset_ceil = index - (nrange + 1) / 2;
  return static_cast<Real>(noffset + noffset_ceil) / (2.0 * nrange);
}
static Real UniformMeshGenerator(Real x, Real xmin, Real xmax) {
  return static_cast<Real>(0.5) * (xmin + xmax) + (x * xmax - x * xmin);
}

static BoundaryFlag GetBoundaryFlag(const std::string &s) {   // bvals/utils/boundary_flag.cpp
  if (s == "periodic") return BoundaryFlag::periodic;
  if (s == "reflecting") return BoundaryFlag::reflect;
  if (s == "outflow") return BoundaryFlag::outflow;
  if (s == "user") return BoundaryFlag::user;
  if (s == "none") return BoundaryFlag::undef;
  std::stringstream msg;
  msg << "### FATAL ERROR in GetBoundaryFlag" << std::endl
      << "Input string=" << s << " is an invalid boundary type";
  ATHENA_ERROR(msg);
}

static std::uint64_t MortonKey(const LogicalLocation &l) {
  std::uint64_t k = 0;
  for (int bit = 0; bit < 21; ++bit) {
    k |= ((std::uint64_t)((l.lx1 >> bit) & 1)) << (3 * bit);
    k |= ((std::uint64_t)((l.lx2 >> bit) & 1)) << (3 * bit + 1);
    k |= ((std::uint64_t)((l.lx3 >> bit) & 1)) << (3 * bit + 2);
  }
  return k;
}

//----------------------------------------------------------------------------------------
Coordinates::Coordinates(MeshBlock *pmb, Mesh *pm, bool coarse) {
  const int ng = NGHOST;
  const int cf = coarse ? 2 : 1;   // coordinates.cpp:108-160: a coarse cell spans two fine ones
  const RegionSize &ms = pm->mesh_size;
  const RegionSize &bs = pmb->block_size;
  struct Dir { AthenaArray<Real> *xf, *xv, *dxf, *dxv; int nx, nc; std::int64_t lx; Real bmin, bmax, mmin, mmax; int mnx; };
  Dir dirs[3] = {
      {&x1f, &x1v, &dx1f, &dx1v, bs.nx1 / cf, bs.nx1 / cf + 2 * ng, pmb->loc.lx1, bs.x1min, bs.x1max, ms.x1min, ms.x1max, ms.nx1},
      {&x2f, &x2v, &dx2f, &dx2v, bs.nx2 / cf, bs.nx2 / cf + 2 * ng, pmb->loc.lx2, bs.x2min, bs.x2max, ms.x2min, ms.x2max, ms.nx2},
      {&x3f, &x3v, &dx3f, &dx3v, bs.nx3 / cf, bs.nx3 / cf + 2 * ng, pmb->loc.lx3, bs.x3min, bs.x3max, ms.x3min, ms.x3max, ms.nx3}};
  for (Dir &d : dirs) {
    d.xf->NewAthenaArray(d.nc + 1);
    d.xv->NewAthenaArray(d.nc);
    d.dxf->NewAthenaArray(d.nc);
    d.dxv->NewAthenaArray(d.nc);
    const int il = ng, iu = ng + d.nx - 1;
    const std::int64_t nrootmesh = (std::int64_t)d.mnx * (1LL << (pmb->loc.level - pm->root_level));   // coordinates.cpp:95
    // coordinates.cpp:143-152 (uniform mesh generator branch)
    const Real dx = (d.bmax - d.bmin) / (iu - il + 1);
    for (int i = il - ng; i <= iu + ng + 1; ++i) {
      const std::int64_t noffset = static_cast<std::int64_t>((i - il) * cf + d.lx * d.nx * cf);
      const Real rx = ComputeMeshGeneratorX(noffset, nrootmesh);
      (*d.xf)(i) = UniformMeshGenerator(rx, d.mmin, d.mmax);
    }
    (*d.xf)(il) = d.bmin;
    (*d.xf)(iu + 1) = d.bmax;
    for (int i = il - ng; i <= iu + ng; ++i) (*d.dxf)(i) = dx;
    // cartesian.cpp:27-38
    for (int i = il - ng; i <= iu + ng; ++i) (*d.xv)(i) = 0.5 * ((*d.xf)(i + 1) + (*d.xf)(i));
    for (int i = il - ng; i <= iu + ng - 1; ++i) (*d.dxv)(i) = (*d.dxf)(i);
  }
}

Hydro::Hydro(MeshBlock *pmb) {
  u.NewAthenaArray(NHYDRO, pmb->ncells3, pmb->ncells2, pmb->ncells1);
  w.NewAthenaArray(NHYDRO, pmb->ncells3, pmb->ncells2, pmb->ncells1);
  if (RELATIVISTIC_DYNAMICS) w1.NewAthenaArray(NHYDRO, pmb->ncells3, pmb->ncells2, pmb->ncells1);
}
Field::Field(MeshBlock *pmb) {
  if (!MAGNETIC_FIELDS_ENABLED) return;
  b.x1f.NewAthenaArray(pmb->ncells3, pmb->ncells2, pmb->ncells1 + 1);
  b.x2f.NewAthenaArray(pmb->ncells3, pmb->ncells2 + 1, pmb->ncells1);
  b.x3f.NewAthenaArray(pmb->ncells3 + 1, pmb->ncells2, pmb->ncells1);
  bcc.NewAthenaArray(3, pmb->ncells3, pmb->ncells2, pmb->ncells1);
}
PassiveScalars::PassiveScalars(MeshBlock *pmb) {
  if (NSCALARS == 0) return;
  s.NewAthenaArray(NSCALARS, pmb->ncells3, pmb->ncells2, pmb->ncells1);
  r.NewAthenaArray(NSCALARS, pmb->ncells3, pmb->ncells2, pmb->ncells1);
}
EquationOfState::EquationOfState(ParameterInput *pin) {   // eos/adiabatic_*.cpp:26-31
  const Real float_min = std::numeric_limits<float>::min();
  if (g_config.isothermal) {   // eos/isothermal_hydro.cpp:25-30, isothermal_mhd.cpp:26-31
    iso_sound_speed_ = pin->GetReal("hydro", "iso_sound_speed");   // error if missing!
    gamma_ = 1.0;   // (never used)
    dfloor_ = pin->GetOrAddReal("hydro", "dfloor", std::sqrt(1024 * float_min));
    pfloor_ = std::sqrt(1024 * float_min);
    sfloor_ = pin->GetOrAddReal("hydro", "sfloor", std::sqrt(1024 * float_min));
    return;
  }
  gamma_ = pin->GetReal("hydro", "gamma");
  dfloor_ = pin->GetOrAddReal("hydro", "dfloor", std::sqrt(1024 * float_min));
  pfloor_ = pin->GetOrAddReal("hydro", "pfloor", std::sqrt(1024 * float_min));
  sfloor_ = pin->GetOrAddReal("hydro", "sfloor", std::sqrt(1024 * float_min));
  if (RELATIVISTIC_DYNAMICS) gamma_max_ = pin->GetOrAddReal("hydro", "gamma_max", 1000.0);   // adiabatic_hydro_sr.cpp:40
}

MeshBlock::MeshBlock(int igid, int ilid, LogicalLocation iloc, RegionSize bs, Mesh *pm, ParameterInput *pin)
    : pmy_mesh(pm), loc(iloc), block_size(bs), gid(igid), lid(ilid), new_block_dt_(0.0) {
  is = NGHOST; ie = is + bs.nx1 - 1;
  js = NGHOST; je = js + bs.nx2 - 1;
  ks = NGHOST; ke = ks + bs.nx3 - 1;
  ncells1 = bs.nx1 + 2 * NGHOST;
  ncells2 = bs.nx2 + 2 * NGHOST;
  ncells3 = bs.nx3 + 2 * NGHOST;
  for (auto &a : nblevel) for (auto &b : a) for (int &v : b) v = -1;   // (set by SearchAndSetNeighbors on refined meshes)
  pcoord = new Coordinates(this, pm);
  phydro = new Hydro(this);
  pfield = new Field(this);
  pscalars = new PassiveScalars(this);
  peos = new EquationOfState(pin);
}
MeshBlock::~MeshBlock() {
  delete pcoarsec;
  delete pcoord; delete phydro; delete pfield; delete pscalars; delete peos;
}
void MeshBlock::ProblemGenerator(ParameterInput *pin) {
  if (pmy_mesh->pgen && pmy_mesh->pgen->problem_generator) pmy_mesh->pgen->problem_generator(this, pin);
}

//----------------------------------------------------------------------------------------
Mesh::Mesh(ParameterInput *pin) : athk(nullptr) {
  std::stringstream msg;
  mesh_size.nx1 = pin->GetInteger("mesh", "nx1");
  mesh_size.nx2 = pin->GetInteger("mesh", "nx2");
  mesh_size.nx3 = pin->GetInteger("mesh", "nx3");
  mesh_size.x1min = pin->GetReal("mesh", "x1min"); mesh_size.x1max = pin->GetReal("mesh", "x1max");
  mesh_size.x2min = pin->GetReal("mesh", "x2min"); mesh_size.x2max = pin->GetReal("mesh", "x2max");
  mesh_size.x3min = pin->GetReal("mesh", "x3min"); mesh_size.x3max = pin->GetReal("mesh", "x3max");
  mesh_size.x1rat = pin->GetOrAddReal("mesh", "x1rat", 1.0);
  mesh_size.x2rat = pin->GetOrAddReal("mesh", "x2rat", 1.0);
  mesh_size.x3rat = pin->GetOrAddReal("mesh", "x3rat", 1.0);
  const char *bn[6] = {"ix1_bc", "ox1_bc", "ix2_bc", "ox2_bc", "ix3_bc", "ox3_bc"};
  for (int n = 0; n < 6; ++n) mesh_bcs[n] = GetBoundaryFlag(pin->GetOrAddString("mesh", bn[n], "none"));
  start_time = pin->GetOrAddReal("time", "start_time", 0.0);
  time = start_time;
  tlim = pin->GetReal("time", "tlim");
  dt = std::numeric_limits<Real>::max();
  nlim = pin->GetOrAddInteger("time", "nlim", -1);
  ncycle = 0;
  ncycle_out = pin->GetOrAddInteger("time", "ncycle_out", 1);
  f2 = mesh_size.nx2 > 1;
  f3 = mesh_size.nx3 > 1;
  ndim = f3 ? 3 : (f2 ? 2 : 1);
  if (ndim != 3) {
    msg << "### FATAL ERROR in Mesh constructor" << std::endl
        << "The B200 path implements 3-D MeshBlocks only (nx1,nx2,nx3 > 1)";
    ATHENA_ERROR(msg);
  }
  if (mesh_size.x1rat != 1.0 || mesh_size.x2rat != 1.0 || mesh_size.x3rat != 1.0) {
    msg << "### FATAL ERROR in Mesh constructor" << std::endl
        << "The B200 path implements uniform Cartesian meshes (x?rat = 1)";
    ATHENA_ERROR(msg);
  }
  {   // mesh.cpp:226-233
    const std::string ref = pin->GetOrAddString("mesh", "refinement", "none");
    if (ref == "static") {
      multilevel = true;
    } else if (ref != "none") {
      msg << "### FATAL ERROR in Mesh constructor" << std::endl
          << "The B200 path implements static mesh refinement (refinement = static); '" << ref << "' is not";
      ATHENA_ERROR(msg);
    }
  }
  if (multilevel) {

    if (NGHOST % 2 == 1) {   // prolongation fills 2*(cnghost-1) ghost cells: more than an odd NGHOST holds
      msg << "### FATAL ERROR in Mesh constructor" << std::endl
          << "Mesh refinement needs an even number of ghost zones (nghost=" << NGHOST << ")";
      ATHENA_ERROR(msg);
    }
  }
  for (int n = 0; n < 6; ++n) {
    if (mesh_bcs[n] != BoundaryFlag::periodic && mesh_bcs[n] != BoundaryFlag::reflect &&
        mesh_bcs[n] != BoundaryFlag::outflow && mesh_bcs[n] != BoundaryFlag::user) {
      msg << "### FATAL ERROR in Mesh constructor" << std::endl
          << "The B200 path implements periodic, reflecting, outflow and user mesh boundaries (" << bn[n] << ")";
      ATHENA_ERROR(msg);
    }
  }
  for (int d = 0; d < 3; ++d) {   // bvals_base.cpp: periodic boundaries come in pairs
    if ((mesh_bcs[2 * d] == BoundaryFlag::periodic) != (mesh_bcs[2 * d + 1] == BoundaryFlag::periodic)) {
      msg << "### FATAL ERROR in Mesh constructor" << std::endl
          << "When periodic boundaries are in use, both sides must be periodic." << std::endl;
      ATHENA_ERROR(msg);
    }
  }
  RegionSize bs = mesh_size;
  bs.nx1 = pin->GetOrAddInteger("meshblock", "nx1", mesh_size.nx1);
  bs.nx2 = pin->GetOrAddInteger("meshblock", "nx2", mesh_size.nx2);
  bs.nx3 = pin->GetOrAddInteger("meshblock", "nx3", mesh_size.nx3);
  if (mesh_size.nx1 % bs.nx1 || mesh_size.nx2 % bs.nx2 || mesh_size.nx3 % bs.nx3) {
    msg << "### FATAL ERROR in Mesh constructor" << std::endl
        << "the Mesh must be evenly divisible by the MeshBlock";   // mesh.cpp:243-248
    ATHENA_ERROR(msg);
  }
  nrbx1 = mesh_size.nx1 / bs.nx1;
  nrbx2 = mesh_size.nx2 / bs.nx2;
  nrbx3 = mesh_size.nx3 / bs.nx3;
  const int nbmax = std::max(nrbx1, std::max(nrbx2, nrbx3));
  for (root_level = 0; (1 << root_level) < nbmax; root_level++) {}   // mesh.cpp:265
  // Z-ordered list of the root-level blocks (meshblock_tree.cpp:336-350)
  for (std::int64_t k = 0; k < nrbx3; ++k)
    for (std::int64_t j = 0; j < nrbx2; ++j)
      for (std::int64_t i = 0; i < nrbx1; ++i) loclist.push_back(LogicalLocation{i, j, k, root_level});
  std::sort(loclist.begin(), loclist.end(),
            [](const LogicalLocation &a, const LogicalLocation &b) { return MortonKey(a) < MortonKey(b); });
  if (multilevel) {   // mesh.cpp:340-492: the block tree with its <refinementN> regions, leaves in Z-order
    if (bs.nx1 % 2 == 1 || bs.nx2 % 2 == 1 || bs.nx3 % 2 == 1) {
      msg << "### FATAL ERROR in Mesh constructor" << std::endl
          << "The size of MeshBlock must be divisible by 2 in order to use SMR or AMR." << std::endl;
      ATHENA_ERROR(msg);
    }
    BuildRefinedTree(pin);
  }
  nbtotal = static_cast<int>(loclist.size());
  nblocal = nbtotal;   // one process drives every GPU: all MeshBlocks are "local" to the host
  // Mesh::CalculateLoadBalance (amr_loadbalance.cpp:73-131) for equal costs: contiguous Z-order
  // ranges, the remainder handed out from the last rank backwards
  ngpu = pin->GetOrAddInteger("mesh", "num_gpus", 1);
  if (ngpu < 1 || ngpu > nbtotal) {
    msg << "### FATAL ERROR in Mesh constructor" << std::endl
        << "<mesh> num_gpus = " << ngpu << " must be between 1 and the number of MeshBlocks (" << nbtotal << ")";
    ATHENA_ERROR(msg);
  }
  ranklist.assign(nbtotal, 0);
  nslist.assign(ngpu, 0);
  nblist.assign(ngpu, 0);
  {
    const int base = nbtotal / ngpu, rem = nbtotal % ngpu;
    int s = 0;
    for (int r = 0; r < ngpu; ++r) {
      nslist[r] = s;
      nblist[r] = base + (r >= ngpu - rem ? 1 : 0);
      for (int n = s; n < s + nblist[r]; ++n) ranklist[n] = r;
      s += nblist[r];
    }
  }

  pgen = FindProblemGenerator(g_config.problem);
  if (pgen == nullptr) {
    msg << "### FATAL ERROR in Mesh constructor" << std::endl
        << "problem generator '" << g_config.problem << "' is not built into this binary";
    ATHENA_ERROR(msg);
  }
  fused_stage = pin->GetOrAddBoolean("time", "fused_stage", true);
  // the flux correction between levels replaces stored face fluxes before they are differenced:
  // refined meshes run the operator sequence (the fused stage never stores fluxes)
  if (multilevel) fused_stage = false;

  for (int n = 0; n < nbtotal; ++n) {
    const LogicalLocation &l = loclist[n];
    RegionSize b = bs;
    // SetBlockSizeAndBoundaries, mesh.cpp:1984-2060
    const int lsh = l.level - root_level;   // nrbx_ll = nrbx << (ll - root_level)
    struct { std::int64_t lx; std::int64_t nrb; Real mmin, mmax; Real *bmin, *bmax; } d[3] = {
        {l.lx1, (std::int64_t)nrbx1 << lsh, mesh_size.x1min, mesh_size.x1max, &b.x1min, &b.x1max},
        {l.lx2, (std::int64_t)nrbx2 << lsh, mesh_size.x2min, mesh_size.x2max, &b.x2min, &b.x2max},
        {l.lx3, (std::int64_t)nrbx3 << lsh, mesh_size.x3min, mesh_size.x3max, &b.x3min, &b.x3max}};
    for (auto &q : d) {
      *q.bmin = (q.lx == 0) ? q.mmin : UniformMeshGenerator(ComputeMeshGeneratorX(q.lx, q.nrb), q.mmin, q.mmax);
      *q.bmax = (q.lx == q.nrb - 1) ? q.mmax
                                    : UniformMeshGenerator(ComputeMeshGeneratorX(q.lx + 1, q.nrb), q.mmin, q.mmax);
    }
    blocks_.push_back(new MeshBlock(n, n - nslist[ranklist[n]], l, b, this, pin));
    blocks_.back()->rank = ranklist[n];
    for (int q = 0; q < 3; ++q) {   // SetBlockSizeAndBoundaries: the mesh's flag on the mesh edge, else "block"
      blocks_.back()->block_bcs[2 * q] = (d[q].lx == 0) ? mesh_bcs[2 * q] : BoundaryFlag::block;
      blocks_.back()->block_bcs[2 * q + 1] = (d[q].lx == d[q].nrb - 1) ? mesh_bcs[2 * q + 1] : BoundaryFlag::block;
    }
  }

  // ---- device contexts: one per GPU, each for the MeshBlocks of its range ------------------
  athk_config cfg;
  std::memset(&cfg, 0, sizeof(cfg));
  cfg.nx1 = bs.nx1; cfg.nx2 = bs.nx2; cfg.nx3 = bs.nx3;
  cfg.nghost = NGHOST;
  cfg.mhd = MAGNETIC_FIELDS_ENABLED ? 1 : 0;
  cfg.nscalars = NSCALARS;
  // configure.py's defaults: hlld with -b, else hllc, or hlle for the isothermal EOS
  const std::string flux = g_config.flux.empty()
                               ? (MAGNETIC_FIELDS_ENABLED ? "hlld" : (g_config.isothermal ? "hlle" : "hllc"))
                               : g_config.flux;
  cfg.eos = g_config.isothermal ? ATHK_EOS_ISOTHERMAL : ATHK_EOS_ADIABATIC;
  if (RELATIVISTIC_DYNAMICS) {   // configure.py:391-409
    if (g_config.isothermal) {
      msg << "### FATAL ERROR in Mesh constructor" << std::endl << "Isothermal EOS is incompatible with relativity";
      ATHENA_ERROR(msg);
    }
    cfg.eos = ATHK_EOS_ADIABATIC_SR;
    cfg.gamma_max = blocks_[0]->peos->GetGammaMax();
  }
  cfg.iso_sound_speed = g_config.isothermal ? blocks_[0]->peos->GetIsoSoundSpeed() : 0.0;
  if (flux == "hllc") cfg.rsolver = ATHK_RS_HLLC;
  else if (flux == "hlld") cfg.rsolver = ATHK_RS_HLLD;
  else if (flux == "roe") cfg.rsolver = ATHK_RS_ROE;
  else if (flux == "hlle") cfg.rsolver = ATHK_RS_HLLE;
  else if (flux == "llf") cfg.rsolver = ATHK_RS_LLF;
  else {
    msg << "### FATAL ERROR in Mesh constructor" << std::endl << "unknown Riemann solver '" << flux << "'";
    ATHENA_ERROR(msg);
  }
  {   // <hydro> gpu_arithmetic = strict (default: the reference's arithmetic, bit for bit) | fast
    const std::string ar = pin->GetOrAddString("hydro", "gpu_arithmetic", "strict");
    if (ar != "strict" && ar != "fast") {
      msg << "### FATAL ERROR in Mesh constructor" << std::endl << "<hydro> gpu_arithmetic must be strict or fast";
      ATHENA_ERROR(msg);
    }
    cfg.arith = (ar == "fast") ? ATHK_ARITH_FAST : ATHK_ARITH_STRICT;
  }
  const int dev0 = pin->GetOrAddInteger("mesh", "gpu_device", 0);
  gpu_device0_ = dev0;
  cfg.gamma = blocks_[0]->peos->GetGamma();
  cfg.dfloor = blocks_[0]->peos->GetDensityFloor();
  cfg.pfloor = blocks_[0]->peos->GetPressureFloor();
  cfg.sfloor = blocks_[0]->peos->GetScalarFloor();
  ctx.assign(ngpu, nullptr);
  for (int r = 0; r < ngpu; ++r) {
    cfg.device = dev0 + r;
    cfg.nblocks = nblist[r];
    AthkCheck(athk_create(&cfg, &ctx[r]), "Mesh constructor (athk_create)");
  }
  athk = ctx[0];
  // each MeshBlock's own EquationOfState: blocks built after the first read the floors back from
  // the 6-digit strings the first one stored (see ParameterInput::GetOrAddReal)
  for (int n = 0; n < nblocal; ++n) {
    const EquationOfState *pe = blocks_[n]->peos;
    // (sfloor = 0, as the reference's SR inputs set it, is not expressible in athk_config, where <= 0
    //  selects the default)
    if (pe->GetDensityFloor() != cfg.dfloor || pe->GetPressureFloor() != cfg.pfloor ||
        pe->GetScalarFloor() != cfg.sfloor || !(cfg.sfloor > 0.0))
      AthkCheck(athk_set_block_floors(ctx_of(blocks_[n]), blocks_[n]->lid, pe->GetDensityFloor(),
                                      pe->GetPressureFloor(), pe->GetScalarFloor()),
                "Mesh constructor (athk_set_block_floors)");
  }
  {   // HydroSourceTerms constructor (hydro/srcterms/hydro_srcterms.cpp:42-46)
    const Real g1 = pin->GetOrAddReal("hydro", "grav_acc1", 0.0);
    const Real g2 = pin->GetOrAddReal("hydro", "grav_acc2", 0.0);
    const Real g3 = pin->GetOrAddReal("hydro", "grav_acc3", 0.0);
    if (g1 != 0.0 || g2 != 0.0 || g3 != 0.0)
      for (athk_ctx *a : ctx)
        AthkCheck(athk_set_constant_acceleration(a, g1, g2, g3), "Mesh constructor (athk_set_constant_acceleration)");
  }
  for (int n = 0; n < nblocal; ++n) {
    Coordinates *pc = blocks_[n]->pcoord;
    AthkCheck(athk_set_block_coords(ctx_of(blocks_[n]), blocks_[n]->lid, pc->x1f.data(), pc->x1v.data(),
                                    pc->dx1f.data(), pc->x2f.data(), pc->x2v.data(), pc->dx2f.data(),
                                    pc->x3f.data(), pc->x3v.data(), pc->dx3f.data()),
              "Mesh constructor (athk_set_block_coords)");
  }
  // same-level neighbour tables, per-GPU remote links, peer-to-peer buffers (multigpu.cpp)
  if (multilevel) SetupRefinement();   // neighbours across levels, coarse buffers, exchange boxes
  else SetupRemoteLinks();
  // BoundaryValues::block_bcs (bvals.cpp:60-120): mesh flag on the mesh edge, "block" elsewhere
  for (int n = 0; n < nblocal; ++n) {
    const LogicalLocation &l = loclist[n];
    const int lsh = l.level - root_level;
    const std::int64_t lx[3] = {l.lx1, l.lx2, l.lx3},
                       nr[3] = {(std::int64_t)nrbx1 << lsh, (std::int64_t)nrbx2 << lsh, (std::int64_t)nrbx3 << lsh};
    int fl[6];
    bool any = false;
    for (int d = 0; d < 3; ++d)
      for (int side = 0; side < 2; ++side) {
        const BoundaryFlag bf = mesh_bcs[2 * d + side];
        const bool edge = side ? (lx[d] == nr[d] - 1) : (lx[d] == 0);
        fl[2 * d + side] = (!edge || bf == BoundaryFlag::periodic) ? ATHK_BC_NONE
                           : (bf == BoundaryFlag::reflect ? ATHK_BC_REFLECT
                                                          : (bf == BoundaryFlag::user ? ATHK_BC_USER : ATHK_BC_OUTFLOW));
        any = any || fl[2 * d + side] != ATHK_BC_NONE;
      }
    if (any) AthkCheck(athk_set_boundary_flags(ctx_of(blocks_[n]), blocks_[n]->lid, fl),
                       "Mesh constructor (athk_set_boundary_flags)");
  }
  if (pgen->init_user_mesh_data) pgen->init_user_mesh_data(this, pin);   // mesh.cpp:601
  for (int n = 0; n < 6; ++n)   // bvals.cpp:349-356
    if (mesh_bcs[n] == BoundaryFlag::user && BoundaryFunction_[n] == nullptr) {
      msg << "### FATAL ERROR in BoundaryValues constructor" << std::endl
          << "A user-defined boundary is specified but the actual boundary function "
          << "is not enrolled in direction " << n << " (in [0,6])." << std::endl;
      ATHENA_ERROR(msg);
    }
}

MeshBlock *Mesh::FindLocalBlock(int rank, int lid) {
  for (MeshBlock *p : blocks_)
    if (p->rank == rank && p->lid == lid) return p;
  return nullptr;
}

// The library's side of a user boundary function is a host callback that enqueues device work
// (athk_enroll_user_boundary); a BValFunc written for the reference works on host arrays, so this
// adapter brings the block's primitives / face field to the host mirrors, runs it, and sends them
// back.  Synchronous and per block: the cost of keeping the reference's host-side function as is.
namespace {
struct UserBCSlot { Mesh *pm; int rank; };
int HostBValAdapter(void *user, const athk_bval_args *a) {
  UserBCSlot *slot = static_cast<UserBCSlot *>(user);
  Mesh *pm = slot->pm;
  MeshBlock *pmb = pm->FindLocalBlock(slot->rank, a->lb);
  if (pmb == nullptr || pm->BoundaryFunction_[a->face] == nullptr) return 1;
  athk_ctx *c = pm->ctx[slot->rank];
  if (athk_download(c, ATHK_W, a->lb, pmb->phydro->w.data())) return 1;
  if (MAGNETIC_FIELDS_ENABLED) {
    if (athk_download(c, ATHK_B_X1F, a->lb, pmb->pfield->b.x1f.data())) return 1;
    if (athk_download(c, ATHK_B_X2F, a->lb, pmb->pfield->b.x2f.data())) return 1;
    if (athk_download(c, ATHK_B_X3F, a->lb, pmb->pfield->b.x3f.data())) return 1;
  }
  pm->BoundaryFunction_[a->face](pmb, pmb->pcoord, pmb->phydro->w, pmb->pfield->b, a->time, a->dt, a->il, a->iu,
                                 a->jl, a->ju, a->kl, a->ku, a->ngh);
  if (athk_upload(c, ATHK_W, a->lb, pmb->phydro->w.data())) return 1;
  if (MAGNETIC_FIELDS_ENABLED) {
    if (athk_upload(c, ATHK_B_X1F, a->lb, pmb->pfield->b.x1f.data())) return 1;
    if (athk_upload(c, ATHK_B_X2F, a->lb, pmb->pfield->b.x2f.data())) return 1;
    if (athk_upload(c, ATHK_B_X3F, a->lb, pmb->pfield->b.x3f.data())) return 1;
  }
  return 0;
}
}  // namespace

namespace {
int HostSrcAdapter(void *user, const athk_src_args *a) {
  UserBCSlot *slot = static_cast<UserBCSlot *>(user);
  Mesh *pm = slot->pm;
  MeshBlock *pmb = pm->FindLocalBlock(slot->rank, a->lb);
  if (pmb == nullptr || pm->UserSourceTerm_ == nullptr) return 1;
  athk_ctx *c = pm->ctx[slot->rank];
  if (athk_download(c, ATHK_W, a->lb, pmb->phydro->w.data())) return 1;
  if (athk_download(c, ATHK_U, a->lb, pmb->phydro->u.data())) return 1;
  if (MAGNETIC_FIELDS_ENABLED && athk_download(c, ATHK_BCC, a->lb, pmb->pfield->bcc.data())) return 1;
  if (NSCALARS > 0) {
    if (athk_download(c, ATHK_R, a->lb, pmb->pscalars->r.data())) return 1;
    if (athk_download(c, ATHK_S, a->lb, pmb->pscalars->s.data())) return 1;
  }
  pm->UserSourceTerm_(pmb, a->time, a->dt, pmb->phydro->w, pmb->pscalars->r, pmb->pfield->bcc, pmb->phydro->u,
                      pmb->pscalars->s);
  if (athk_upload(c, ATHK_U, a->lb, pmb->phydro->u.data())) return 1;
  if (NSCALARS > 0 && athk_upload(c, ATHK_S, a->lb, pmb->pscalars->s.data())) return 1;
  return 0;
}
}  // namespace

void Mesh::EnrollUserExplicitSourceFunction(SrcTermFunc my_func) {
  UserSourceTerm_ = my_func;
  static std::vector<UserBCSlot *> slots;
  for (int r = 0; r < ngpu; ++r) {
    UserBCSlot *sl = new UserBCSlot{this, r};
    slots.push_back(sl);
    AthkCheck(athk_enroll_user_source(ctx[r], HostSrcAdapter, sl), "Mesh::EnrollUserExplicitSourceFunction");
  }
}

void Mesh::EnrollUserBoundaryFunction(BoundaryFace dir, BValFunc my_bc) {
  std::stringstream msg;
  if (dir < 0 || dir > 5) {
    msg << "### FATAL ERROR in EnrollUserBoundaryFunction" << std::endl
        << "dirName = " << dir << " not valid" << std::endl;
    ATHENA_ERROR(msg);
  }
  if (mesh_bcs[dir] != BoundaryFlag::user) {
    msg << "### FATAL ERROR in EnrollUserBoundaryFunction" << std::endl
        << "The boundary condition flag must be set to the string 'user' in the "
        << " <mesh> block in the input file to use user-enrolled BCs" << std::endl;
    ATHENA_ERROR(msg);
  }
  BoundaryFunction_[dir] = my_bc;
  static std::vector<UserBCSlot *> slots;   // live as long as the process (one Mesh per run)
  for (int r = 0; r < ngpu; ++r) {
    UserBCSlot *sl = new UserBCSlot{this, r};
    slots.push_back(sl);
    AthkCheck(athk_enroll_user_boundary(ctx[r], dir, HostBValAdapter, sl), "Mesh::EnrollUserBoundaryFunction");
  }
}

Mesh::~Mesh() {
  for (MeshBlock *p : blocks_) delete p;
  delete tree;
  for (athk_ctx *a : ctx) if (a) athk_sync(a);
  for (std::size_t q = 0; q < p2p_allocs_.size(); ++q)
    if (p2p_allocs_[q]) athk_p2p_free(ctx[q / 3], p2p_allocs_[q]);
  for (athk_ctx *a : ctx) if (a) athk_destroy(a);
}

// SEND_HYD/_FLD/_SCLR, RECV, SETB for every GPU: pack the messages for neighbours on other GPUs
// (peer stores), then fill the ghost zones
void Mesh::ExchangeGhosts() {
  const int what = 1 | (MAGNETIC_FIELDS_ENABLED ? 2 : 0) | (NSCALARS > 0 ? 4 : 0);
  if (multilevel) {
    // SEND_HYD: blocks next to a coarser one restrict their boundary cells (phase 1); SETB_HYD: same
    // level (neighbour table), from finer / from coarser (phase 2); PROLONG (time_integrator.cpp:
    // 2133-2160, bvals_refine.cpp:95): restrict the ghost-ghost zones, W(U) on the coarse buffer,
    // prolongate the primitives, U(W) on the new ghost zones (phases 3-6)
    const char *who = "Mesh::ExchangeGhosts";
    RunRefineList(1);
    if (ngpu > 1) for (athk_ctx *a : ctx) AthkCheck(athk_remote_pack_ghosts(a, 0), who);
    for (athk_ctx *a : ctx) AthkCheck(athk_exchange_ghosts(a, what & ~2, 0), who);
    // SETB_FLD on a refined mesh (refinement.cpp, SetupFieldRefinement): face, edge, corner neighbours in
    // turn (same level and from finer), then from coarser into the coarse buffer
    if (MAGNETIC_FIELDS_ENABLED) for (int list : {8, 9, 10}) RunRefineList(list);
    RunRefineList(2);
    if (MAGNETIC_FIELDS_ENABLED) for (int list : {11, 64, 65}) RunRefineList(list);
    // (list 7 = the physical boundary functions on the coarse primitives, between W(U) and prolongation;
    //  12, 13 = shared and internal faces of the prolongated field, before U(W, bcc) of list 6)
    for (int list : {3, 4, 7}) RunRefineList(list);
    ApplyUserBoundariesOnCoarseLevel();   // (host functions, like the fine level's: Mesh::EnrollUserBoundaryFunction)
    RunRefineList(5);
    if (MAGNETIC_FIELDS_ENABLED) for (int list : {12, 13}) RunRefineList(list);
    RunRefineList(6);
    if (ngpu > 1) SyncAll();   // nobody overwrites what another GPU's boxes still read
    return;
  }
  if (ngpu > 1) for (athk_ctx *a : ctx) AthkCheck(athk_remote_pack_ghosts(a, 0), "Mesh::ExchangeGhosts");
  for (athk_ctx *a : ctx) AthkCheck(athk_exchange_ghosts(a, what, 0), "Mesh::ExchangeGhosts");
}

// Mesh::Initialize for a fresh run (mesh.cpp:1622-1967): pgen on the host, upload, first
// boundary exchange, W(U) on interior + ghosts, first time step
void Mesh::Initialize(ParameterInput *pin) {
  for (int b = 0; b < nblocal; ++b) {
    MeshBlock *pmb = blocks_[b];
    if (!res_flag) pmb->ProblemGenerator(pin);   // restart: the arrays were filled by LoadRestart
    athk_ctx *a = ctx_of(pmb);
    const int l = pmb->lid;
    AthkCheck(athk_upload(a, ATHK_U, l, pmb->phydro->u.data()), "Mesh::Initialize");
    if (MAGNETIC_FIELDS_ENABLED) {
      AthkCheck(athk_upload(a, ATHK_B_X1F, l, pmb->pfield->b.x1f.data()), "Mesh::Initialize");
      AthkCheck(athk_upload(a, ATHK_B_X2F, l, pmb->pfield->b.x2f.data()), "Mesh::Initialize");
      AthkCheck(athk_upload(a, ATHK_B_X3F, l, pmb->pfield->b.x3f.data()), "Mesh::Initialize");
    }
    if (NSCALARS > 0) AthkCheck(athk_upload(a, ATHK_S, l, pmb->pscalars->s.data()), "Mesh::Initialize");
    if (RELATIVISTIC_DYNAMICS && !res_flag) {
      // what the pgen left in w / w1 seeds the fixed-point W(U).  The first W(U) (below) writes Hydro::w
      // (mesh.cpp:1833), the next one Hydro::w1 (time_integrator.cpp:2221); the device always writes the
      // array ATHK_W1 names and then swaps the two names
      AthkCheck(athk_upload(a, ATHK_W1, l, pmb->phydro->w.data()), "Mesh::Initialize");
      AthkCheck(athk_upload(a, ATHK_W, l, pmb->phydro->w1.data()), "Mesh::Initialize");
    }
  }
  for (athk_ctx *a : ctx) AthkCheck(athk_sync(a), "Mesh::Initialize");
  ExchangeGhosts();
  for (athk_ctx *a : ctx) AthkCheck(athk_cons2prim(a, 0), "Mesh::Initialize");
  for (athk_ctx *a : ctx) AthkCheck(athk_set_boundary_time(a, time, 0.0), "Mesh::Initialize");
  bc_time_ = time; bc_dt_ = 0.0;
  for (athk_ctx *a : ctx) AthkCheck(athk_apply_physical_boundaries(a), "Mesh::Initialize");   // mesh.cpp:1886-1891
  NewTimeStep();
}

// Restart: the header after <par_end> (restart.cpp:94-117), the per-block id list and the raw
// register dumps.  The Mesh has already been built from the parameters the file carries; the
// file must describe the same block list.
void Mesh::LoadRestart(const std::string &raw, std::size_t off) {
  std::stringstream msg;
  auto take = [&](void *dst, std::size_t n) {
    if (off + n > raw.size()) {
      msg << "### FATAL ERROR in Mesh constructor" << std::endl << "The restart file is broken." << std::endl;
      ATHENA_ERROR(msg);
    }
    std::memcpy(dst, raw.data() + off, n);
    off += n;
  };
  int nb_file, root_file;
  RegionSize ms_file;
  std::uint64_t datasize;
  take(&nb_file, sizeof(int));
  take(&root_file, sizeof(int));
  take(&ms_file, sizeof(RegionSize));
  take(&time, sizeof(Real));
  take(&dt, sizeof(Real));
  take(&ncycle, sizeof(int));
  take(&datasize, sizeof(std::uint64_t));
  if (nb_file != nbtotal || root_file != root_level || ms_file.nx1 != mesh_size.nx1 || ms_file.nx2 != mesh_size.nx2 ||
      ms_file.nx3 != mesh_size.nx3) {
    msg << "### FATAL ERROR in Mesh constructor" << std::endl
        << "The restart file holds " << nb_file << " MeshBlocks of a " << ms_file.nx1 << "x" << ms_file.nx2 << "x"
        << ms_file.nx3 << " mesh; the parameters describe " << nbtotal << " blocks of a " << mesh_size.nx1 << "x"
        << mesh_size.nx2 << "x" << mesh_size.nx3 << " mesh" << std::endl;
    ATHENA_ERROR(msg);
  }
  // the MeshBlocks, their coordinates and the device geometry were built from the parsed <mesh>
  // block: a restart whose -i / command line moved the domain would run on other coordinates than
  // the file's (the reference builds its blocks from the file's RegionSize)
  if (ms_file.x1min != mesh_size.x1min || ms_file.x1max != mesh_size.x1max || ms_file.x2min != mesh_size.x2min ||
      ms_file.x2max != mesh_size.x2max || ms_file.x3min != mesh_size.x3min || ms_file.x3max != mesh_size.x3max ||
      ms_file.x1rat != mesh_size.x1rat || ms_file.x2rat != mesh_size.x2rat || ms_file.x3rat != mesh_size.x3rat) {
    std::stringstream msg;
    msg << "### FATAL ERROR in Mesh::LoadRestart" << std::endl
        << "the <mesh> limits / ratios in the parameters differ from the restart file's RegionSize" << std::endl;
    ATHENA_ERROR(msg);
  }
  mesh_size = ms_file;
  for (int n = 0; n < nbtotal; ++n) {
    LogicalLocation l;
    double cost;
    take(&l, sizeof(LogicalLocation));
    take(&cost, sizeof(double));
    if (l.lx1 != loclist[n].lx1 || l.lx2 != loclist[n].lx2 || l.lx3 != loclist[n].lx3 || l.level != loclist[n].level) {
      msg << "### FATAL ERROR in Mesh constructor" << std::endl
          << "MeshBlock " << n << " of the restart file is not at the location the block list gives it" << std::endl;
      ATHENA_ERROR(msg);
    }
  }
  MeshBlock *p0 = blocks_[0];
  std::uint64_t expect = p0->phydro->u.GetSizeInBytes();
  if (MAGNETIC_FIELDS_ENABLED)
    expect += p0->pfield->b.x1f.GetSizeInBytes() + p0->pfield->b.x2f.GetSizeInBytes() +
              p0->pfield->b.x3f.GetSizeInBytes();
  if (NSCALARS > 0) expect += p0->pscalars->s.GetSizeInBytes();
  if (datasize != expect) {   // meshblock.cpp:411-417
    msg << "### FATAL ERROR in MeshBlock constructor" << std::endl
        << "The restart file is broken or input parameters are inconsistent (data size per MeshBlock "
        << datasize << " != " << expect << ": check -b, --nghost, --nscalars)." << std::endl;
    ATHENA_ERROR(msg);
  }
  for (int n = 0; n < nbtotal; ++n) {
    MeshBlock *pmb = blocks_[n];
    take(pmb->phydro->u.data(), pmb->phydro->u.GetSizeInBytes());
    if (MAGNETIC_FIELDS_ENABLED) {
      take(pmb->pfield->b.x1f.data(), pmb->pfield->b.x1f.GetSizeInBytes());
      take(pmb->pfield->b.x2f.data(), pmb->pfield->b.x2f.GetSizeInBytes());
      take(pmb->pfield->b.x3f.data(), pmb->pfield->b.x3f.GetSizeInBytes());
    }
    if (NSCALARS > 0) take(pmb->pscalars->s.data(), pmb->pscalars->s.GetSizeInBytes());
  }
  res_flag = true;
}

void Mesh::NewTimeStep() {   // mesh.cpp:1236-1282
  std::vector<Real> bdt(nblocal);
  Real dtmin;
  for (int r = 0; r < ngpu; ++r)
    AthkCheck(athk_new_block_dt(ctx[r], cfl_number, bdt.data() + nslist[r], &dtmin), "Mesh::NewTimeStep");
  for (int b = 0; b < nblocal; ++b) blocks_[b]->new_block_dt_ = bdt[b];
  dt = static_cast<Real>(2.0) * dt;
  for (int b = 0; b < nblocal; ++b) dt = std::min(dt, blocks_[b]->new_block_dt_);
  if (time < tlim && (tlim - time) < dt) dt = tlim - time;
}

void Mesh::DownloadAll() {
  for (int b = 0; b < nblocal; ++b) {
    MeshBlock *pmb = blocks_[b];
    athk_ctx *athk = ctx_of(pmb);   // (shadows the member: this block's GPU)
    const int l = pmb->lid;
    AthkCheck(athk_download(athk, ATHK_U, l, pmb->phydro->u.data()), "Mesh::DownloadAll");
    AthkCheck(athk_download(athk, ATHK_W, l, pmb->phydro->w.data()), "Mesh::DownloadAll");
    if (MAGNETIC_FIELDS_ENABLED) {
      AthkCheck(athk_download(athk, ATHK_B_X1F, l, pmb->pfield->b.x1f.data()), "Mesh::DownloadAll");
      AthkCheck(athk_download(athk, ATHK_B_X2F, l, pmb->pfield->b.x2f.data()), "Mesh::DownloadAll");
      AthkCheck(athk_download(athk, ATHK_B_X3F, l, pmb->pfield->b.x3f.data()), "Mesh::DownloadAll");
      AthkCheck(athk_download(athk, ATHK_BCC, l, pmb->pfield->bcc.data()), "Mesh::DownloadAll");
    }
    if (NSCALARS > 0) {
      AthkCheck(athk_download(athk, ATHK_S, l, pmb->pscalars->s.data()), "Mesh::DownloadAll");
      AthkCheck(athk_download(athk, ATHK_R, l, pmb->pscalars->r.data()), "Mesh::DownloadAll");
    }
  }
}

void Mesh::UserWorkAfterLoop(ParameterInput *pin) {
  if (pgen && pgen->user_work_after_loop) pgen->user_work_after_loop(this, pin);
}

void Mesh::OutputCycleDiagnostics() {   // mesh.cpp:2304-2333
  if (ncycle_out != 0 && ncycle % ncycle_out == 0) {
    char buf[160];
    std::snprintf(buf, sizeof(buf), "cycle=%d time=%.16e dt=%.16e", ncycle, time, dt);
    std::cout << buf << std::endl;
  }
}
