// refinement.cpp -- static mesh refinement on the host side: the block tree, the neighbour search
// across levels, and the index boxes of the cell-centred boundary exchange between levels, handed to
// the library as box operations (include/athk.h, athk_refine_*; the arithmetic runs on the device).
//
// Reference (file:line under /root/reference/src):
//   MeshBlockTree                      mesh/meshblock_tree.cpp:48-470 (CreateRootGrid, AddMeshBlock, Refine
//                                      with the 2:1 balance, GetMeshBlockList, FindNeighbor)
//   <refinementN> regions              mesh/mesh.cpp:340-483
//   BoundaryBase::SearchAndSetNeighbors bvals/bvals_base.cpp:301-738 (order: faces, edges, corners)
//   Load / Set boundary buffers         bvals/cc/bvals_cc.cpp:309-560
//   ProlongateBoundaries                bvals/bvals_refine.cpp:95-330, 553-660
//   flux correction                     bvals/cc/flux_correction_cc.cpp:77-275
#include <algorithm>
#include <cmath>

#include "athena.hpp"

//----------------------------------------------------------------------------------------
// MeshBlockTree
MeshBlockTree::MeshBlockTree(Mesh *pm) : pmesh_(pm), proot_(this) {
  loc_ = LogicalLocation{0, 0, 0, 0};
  for (auto &p : pleaf_) p = nullptr;
}
MeshBlockTree::MeshBlockTree(MeshBlockTree *parent, int ox1, int ox2, int ox3)
    : pmesh_(parent->pmesh_), proot_(parent->proot_) {
  loc_.lx1 = (parent->loc_.lx1 << 1) + ox1;
  loc_.lx2 = (parent->loc_.lx2 << 1) + ox2;
  loc_.lx3 = (parent->loc_.lx3 << 1) + ox3;
  loc_.level = parent->loc_.level + 1;
  for (auto &p : pleaf_) p = nullptr;
}
MeshBlockTree::~MeshBlockTree() {
  for (auto &p : pleaf_) delete p;
}

void MeshBlockTree::CreateRootGrid() {   // meshblock_tree.cpp:87-110 (3-D: 8 leaves)
  if (loc_.level == pmesh_->root_level) return;
  has_leaves_ = true;
  const std::int64_t levfac = 1LL << (pmesh_->root_level - loc_.level - 1);
  for (int n = 0; n < 8; ++n) {
    const int i = n & 1, j = (n >> 1) & 1, k = (n >> 2) & 1;
    if ((loc_.lx3 * 2 + k) * levfac < pmesh_->nrbx3 && (loc_.lx2 * 2 + j) * levfac < pmesh_->nrbx2 &&
        (loc_.lx1 * 2 + i) * levfac < pmesh_->nrbx1) {
      pleaf_[n] = new MeshBlockTree(this, i, j, k);
      pleaf_[n]->CreateRootGrid();
    }
  }
}

void MeshBlockTree::AddMeshBlock(LogicalLocation rloc) {   // :116-132
  if (loc_.level == rloc.level) return;
  if (!has_leaves_) Refine();
  const int sh = rloc.level - loc_.level - 1;
  const int mx = ((rloc.lx1 >> sh) & 1LL) == 1LL, my = ((rloc.lx2 >> sh) & 1LL) == 1LL,
            mz = ((rloc.lx3 >> sh) & 1LL) == 1LL;
  pleaf_[mx + (my << 1) + (mz << 2)]->AddMeshBlock(rloc);
}

void MeshBlockTree::Refine() {   // :166-254: the neighbours are refined up to this level first
  if (has_leaves_) return;
  has_leaves_ = true;
  for (int n = 0; n < 8; ++n) pleaf_[n] = new MeshBlockTree(this, n & 1, (n >> 1) & 1, (n >> 2) & 1);
  const int sh = loc_.level - pmesh_->root_level;
  const std::int64_t nmax[3] = {(std::int64_t)pmesh_->nrbx1 << sh, (std::int64_t)pmesh_->nrbx2 << sh,
                                (std::int64_t)pmesh_->nrbx3 << sh};
  LogicalLocation nloc;
  nloc.level = loc_.level;
  auto wrap = [&](std::int64_t l, int d, bool *ok) {
    *ok = true;
    if (l < 0) {
      if (pmesh_->mesh_bcs[2 * d] != BoundaryFlag::periodic) *ok = false;
      else l = nmax[d] - 1;
    }
    if (l >= nmax[d]) {
      if (pmesh_->mesh_bcs[2 * d + 1] != BoundaryFlag::periodic) *ok = false;
      else l = 0;
    }
    return l;
  };
  for (int oz = -1; oz <= 1; ++oz) {
    bool ok3, ok2, ok1;
    nloc.lx3 = wrap(loc_.lx3 + oz, 2, &ok3);
    if (!ok3) continue;
    for (int oy = -1; oy <= 1; ++oy) {
      nloc.lx2 = wrap(loc_.lx2 + oy, 1, &ok2);
      if (!ok2) continue;
      for (int ox = -1; ox <= 1; ++ox) {
        if (ox == 0 && oy == 0 && oz == 0) continue;
        nloc.lx1 = wrap(loc_.lx1 + ox, 0, &ok1);
        if (!ok1) continue;
        proot_->AddMeshBlock(nloc);
      }
    }
  }
  gid_ = -1;
}

void MeshBlockTree::GetMeshBlockList(std::vector<LogicalLocation> *list) {   // :336-350, Z-order
  if (!has_leaves_) {
    gid_ = static_cast<int>(list->size());
    list->push_back(loc_);
    return;
  }
  for (int n = 0; n < 8; ++n)
    if (pleaf_[n] != nullptr) pleaf_[n]->GetMeshBlockList(list);
}

MeshBlockTree *MeshBlockTree::FindNeighbor(LogicalLocation myloc, int ox1, int ox2, int ox3) {   // :361-470
  std::int64_t l[3] = {myloc.lx1 + ox1, myloc.lx2 + ox2, myloc.lx3 + ox3};
  const int ll = myloc.level;
  const int nrb[3] = {pmesh_->nrbx1, pmesh_->nrbx2, pmesh_->nrbx3};
  for (int d = 0; d < 3; ++d) {
    const std::int64_t nmax = (std::int64_t)nrb[d] << (ll - pmesh_->root_level);
    if (l[d] < 0) {
      if (pmesh_->mesh_bcs[2 * d] == BoundaryFlag::periodic) l[d] = nmax - 1;
      else return nullptr;
    }
    if (l[d] >= nmax) {
      if (pmesh_->mesh_bcs[2 * d + 1] == BoundaryFlag::periodic) l[d] = 0;
      else return nullptr;
    }
  }
  MeshBlockTree *bt = proot_;
  if (ll < 1) return proot_;
  for (int level = 0; level < ll; ++level) {
    if (!bt->has_leaves_) {
      if (level == ll - 1) return bt;   // a coarser leaf
      std::stringstream msg;
      msg << "### FATAL ERROR in FindNeighbor" << std::endl
          << "Neighbor search failed. The Block Tree is broken." << std::endl;
      ATHENA_ERROR(msg);
    }
    const int sh = ll - level - 1;
    bt = bt->GetLeaf(((l[0] >> sh) & 1LL) == 1LL, ((l[1] >> sh) & 1LL) == 1LL, ((l[2] >> sh) & 1LL) == 1LL);
    if (bt == nullptr) {
      std::stringstream msg;
      msg << "### FATAL ERROR in FindNeighbor" << std::endl
          << "Neighbor search failed. The Block Tree is broken." << std::endl;
      ATHENA_ERROR(msg);
    }
  }
  return bt;   // a leaf on the same level, or the parent of finer leaves
}

//----------------------------------------------------------------------------------------
// <refinementN> blocks -> tree -> Z-ordered location list (mesh.cpp:340-492)
static Real MeshGenX(std::int64_t index, std::int64_t nrange) {   // mesh.hpp:457-473
  const std::int64_t noffset = index - (nrange) / 2;
  const std::int64_t noffset_ceil = index - (nrange + 1) / 2;
  return static_cast<Real>(noffset + noffset_ceil) / (2.0 * nrange);
}
static Real UniformGen(Real x, Real xmin, Real xmax) { return static_cast<Real>(0.5) * (xmin + xmax) + (x * xmax - x * xmin); }

void Mesh::BuildRefinedTree(ParameterInput *pin) {
  std::stringstream msg;
  tree = new MeshBlockTree(this);
  tree->CreateRootGrid();
  current_level = root_level;
  std::vector<std::string> names;
  for (const ParameterInput::Block &b : pin->blocks())
    if (b.name.compare(0, 10, "refinement") == 0) names.push_back(b.name);
  for (const std::string &bn : names) {
    RegionSize ref;
    ref.x1min = pin->GetReal(bn, "x1min"); ref.x1max = pin->GetReal(bn, "x1max");
    ref.x2min = pin->GetReal(bn, "x2min"); ref.x2max = pin->GetReal(bn, "x2max");
    ref.x3min = pin->GetReal(bn, "x3min"); ref.x3max = pin->GetReal(bn, "x3max");
    const int ref_lev = pin->GetInteger(bn, "level");
    const int lrlev = ref_lev + root_level;
    if (lrlev > current_level) current_level = lrlev;
    if (ref_lev < 1) {
      msg << "### FATAL ERROR in Mesh constructor" << std::endl
          << "Refinement level must be larger than 0 (root level = 0)" << std::endl;
      ATHENA_ERROR(msg);
    }
    if (ref.x1min > ref.x1max || ref.x2min > ref.x2max || ref.x3min > ref.x3max) {
      msg << "### FATAL ERROR in Mesh constructor" << std::endl << "Invalid refinement region is specified." << std::endl;
      ATHENA_ERROR(msg);
    }
    if (ref.x1min < mesh_size.x1min || ref.x1max > mesh_size.x1max || ref.x2min < mesh_size.x2min ||
        ref.x2max > mesh_size.x2max || ref.x3min < mesh_size.x3min || ref.x3max > mesh_size.x3max) {
      msg << "### FATAL ERROR in Mesh constructor" << std::endl
          << "Refinement region must be smaller than the whole mesh." << std::endl;
      ATHENA_ERROR(msg);
    }
    // logical range of the region on its level
    std::int64_t lmin[3], lmax[3];
    const int nrb[3] = {nrbx1, nrbx2, nrbx3};
    const Real rmin[3] = {ref.x1min, ref.x2min, ref.x3min}, rmax[3] = {ref.x1max, ref.x2max, ref.x3max};
    const Real mmin[3] = {mesh_size.x1min, mesh_size.x2min, mesh_size.x3min},
               mmax[3] = {mesh_size.x1max, mesh_size.x2max, mesh_size.x3max};
    for (int d = 0; d < 3; ++d) {
      const std::int64_t lxmax = nrb[d] * (1LL << ref_lev);
      for (lmin[d] = 0; lmin[d] < lxmax; lmin[d]++)
        if (UniformGen(MeshGenX(lmin[d] + 1, lxmax), mmin[d], mmax[d]) > rmin[d]) break;
      for (lmax[d] = lmin[d]; lmax[d] < lxmax; lmax[d]++)
        if (UniformGen(MeshGenX(lmax[d] + 1, lxmax), mmin[d], mmax[d]) >= rmax[d]) break;
      if (lmin[d] % 2 == 1) lmin[d]--;
      if (lmax[d] % 2 == 0) lmax[d]++;
    }
    for (std::int64_t k = lmin[2]; k < lmax[2]; k += 2)
      for (std::int64_t j = lmin[1]; j < lmax[1]; j += 2)
        for (std::int64_t i = lmin[0]; i < lmax[0]; i += 2) tree->AddMeshBlock(LogicalLocation{i, j, k, lrlev});
  }
  loclist.clear();
  tree->GetMeshBlockList(&loclist);
}

//----------------------------------------------------------------------------------------
// BoundaryBase::SearchAndSetNeighbors for one MeshBlock (3-D, multilevel: nf1 = nf2 = 2)
void Mesh::SearchAndSetNeighbors(MeshBlock *pmb) {
  const LogicalLocation &loc = pmb->loc;
  std::vector<NeighborBlock> &nbs = pmb->neighbor;
  nbs.clear();
  for (auto &a : pmb->nblevel) for (auto &b : a) for (int &v : b) v = -1;
  pmb->nblevel[1][1][1] = loc.level;
  const int myox1 = ((loc.lx1 & 1LL) == 1LL) * 2 - 1, myox2 = ((loc.lx2 & 1LL) == 1LL) * 2 - 1,
            myox3 = ((loc.lx3 & 1LL) == 1LL) * 2 - 1;
  auto add = [&](MeshBlockTree *t, int o1, int o2, int o3, int type, int f1, int f2) {
    NeighborBlock nb;
    nb.gid = t->gid_; nb.lid = t->gid_; nb.level = t->loc_.level;
    nb.ox1 = o1; nb.ox2 = o2; nb.ox3 = o3; nb.fi1 = f1; nb.fi2 = f2; nb.type = type;
    nb.fid = -1;
    if (type == 0) nb.fid = (o1 != 0) ? (o1 + 1) / 2 : ((o2 != 0) ? 2 + (o2 + 1) / 2 : 4 + (o3 + 1) / 2);
    nbs.push_back(nb);
  };
  // faces: x1, x2, x3
  for (int dir = 0; dir < 3; ++dir)
    for (int n = -1; n <= 1; n += 2) {
      const int o1 = dir == 0 ? n : 0, o2 = dir == 1 ? n : 0, o3 = dir == 2 ? n : 0;
      MeshBlockTree *nt = tree->FindNeighbor(loc, o1, o2, o3);
      if (nt == nullptr) continue;
      if (nt->has_leaves_) {   // finer
        const int fface = 1 - (n + 1) / 2;
        pmb->nblevel[o3 + 1][o2 + 1][o1 + 1] = nt->loc_.level + 1;
        for (int f2 = 0; f2 < 2; ++f2)
          for (int f1 = 0; f1 < 2; ++f1) {
            MeshBlockTree *nf = dir == 0 ? nt->GetLeaf(fface, f1, f2)
                                         : (dir == 1 ? nt->GetLeaf(f1, fface, f2) : nt->GetLeaf(f1, f2, fface));
            add(nf, o1, o2, o3, 0, f1, f2);
          }
      } else {
        pmb->nblevel[o3 + 1][o2 + 1][o1 + 1] = nt->loc_.level;
        add(nt, o1, o2, o3, 0, 0, 0);
      }
    }
  // edges: x1x2, x1x3, x2x3
  for (int e = 0; e < 3; ++e)
    for (int m = -1; m <= 1; m += 2)
      for (int n = -1; n <= 1; n += 2) {
        const int o1 = (e == 2) ? 0 : n, o2 = (e == 0) ? m : (e == 2 ? n : 0), o3 = (e == 0) ? 0 : m;
        MeshBlockTree *nt = tree->FindNeighbor(loc, o1, o2, o3);
        if (nt == nullptr) continue;
        if (nt->has_leaves_) {
          const int ff1 = 1 - (n + 1) / 2, ff2 = 1 - (m + 1) / 2;
          pmb->nblevel[o3 + 1][o2 + 1][o1 + 1] = nt->loc_.level + 1;
          for (int f1 = 0; f1 < 2; ++f1) {
            MeshBlockTree *nf = e == 0 ? nt->GetLeaf(ff1, ff2, f1)
                                       : (e == 1 ? nt->GetLeaf(ff1, f1, ff2) : nt->GetLeaf(f1, ff1, ff2));
            add(nf, o1, o2, o3, 1, f1, 0);
          }
        } else {
          const int nlevel = nt->loc_.level;
          pmb->nblevel[o3 + 1][o2 + 1][o1 + 1] = nlevel;
          const bool mine = (e == 0) ? (myox1 == n && myox2 == m)
                                     : (e == 1 ? (myox1 == n && myox3 == m) : (myox2 == n && myox3 == m));
          if (nlevel >= loc.level || mine) add(nt, o1, o2, o3, 1, 0, 0);
        }
      }
  // corners
  for (int l = -1; l <= 1; l += 2)
    for (int m = -1; m <= 1; m += 2)
      for (int n = -1; n <= 1; n += 2) {
        MeshBlockTree *nt = tree->FindNeighbor(loc, n, m, l);
        if (nt == nullptr) continue;
        if (nt->has_leaves_) nt = nt->GetLeaf(1 - (n + 1) / 2, 1 - (m + 1) / 2, 1 - (l + 1) / 2);
        const int nlevel = nt->loc_.level;
        pmb->nblevel[l + 1][m + 1][n + 1] = nlevel;
        if (nlevel >= loc.level || (myox1 == n && myox2 == m && myox3 == l)) add(nt, n, m, l, 2, 0, 0);
      }
}

//----------------------------------------------------------------------------------------
namespace {
struct Box { int si, ei, sj, ej, sk, ek; long long count() const { return (long long)(ei - si + 1) * (ej - sj + 1) * (ek - sk + 1); } };

athk_box_op MakeOp(int kind, int dst, int src, int dir, const Box &d, int so1, int so2, int so3) {
  athk_box_op o;
  o.kind = kind; o.dst = dst; o.src = src; o.dir = dir;
  o.lo[0] = d.si; o.hi[0] = d.ei; o.lo[1] = d.sj; o.hi[1] = d.ej; o.lo[2] = d.sk; o.hi[2] = d.ek;
  o.so[0] = so1; o.so[1] = so2; o.so[2] = so3;
  o.pad = 0;
  return o;
}
}  // namespace

// coarse twin of the block's x?v (Coordinates with coarse_flag, coordinates.cpp:33-37,108-160 +
// cartesian.cpp:27-29): nx/2 + 2*NGHOST cells starting NGHOST coarse cells below the block
static void CoarseCellCentres(const MeshBlock *pmb, const Mesh *pm, int d, std::vector<Real> *xv) {
  const int ng = NGHOST;
  const int nx = (d == 0) ? pmb->block_size.nx1 : (d == 1 ? pmb->block_size.nx2 : pmb->block_size.nx3);
  const std::int64_t lx = (d == 0) ? pmb->loc.lx1 : (d == 1 ? pmb->loc.lx2 : pmb->loc.lx3);
  const int mnx = (d == 0) ? pm->mesh_size.nx1 : (d == 1 ? pm->mesh_size.nx2 : pm->mesh_size.nx3);
  const Real mmin = (d == 0) ? pm->mesh_size.x1min : (d == 1 ? pm->mesh_size.x2min : pm->mesh_size.x3min);
  const Real mmax = (d == 0) ? pm->mesh_size.x1max : (d == 1 ? pm->mesh_size.x2max : pm->mesh_size.x3max);
  const Real bmin = (d == 0) ? pmb->block_size.x1min : (d == 1 ? pmb->block_size.x2min : pmb->block_size.x3min);
  const Real bmax = (d == 0) ? pmb->block_size.x1max : (d == 1 ? pmb->block_size.x2max : pmb->block_size.x3max);
  const int il = ng, iu = ng + nx / 2 - 1, nc = nx / 2 + 2 * ng;
  const std::int64_t nrootmesh = (std::int64_t)mnx * (1LL << (pmb->loc.level - pm->root_level));
  std::vector<Real> xf(nc + 1);
  for (int i = il - ng; i <= iu + ng + 1; ++i) {
    const std::int64_t noffset = static_cast<std::int64_t>((i - il) * 2 + lx * nx);
    xf[i] = UniformGen(MeshGenX(noffset, nrootmesh), mmin, mmax);
  }
  xf[il] = bmin;
  xf[iu + 1] = bmax;
  xv->assign(nc, 0.0);
  for (int i = il - ng; i <= iu + ng; ++i) (*xv)[i] = 0.5 * (xf[i + 1] + xf[i]);
}

// Everything the device needs for the exchange between levels: coarse buffers, coordinates, the
// neighbour table (same level: block index; other level: ATHK_NBR_REFINED) and the boxes per phase.
void Mesh::SetupRefinement() {
  const char *who = "Mesh::SetupRefinement";
  AthkCheck(athk_refine_enable(athk), who);
  const int ng = NGHOST, cng = (NGHOST + 1) / 2 + 1;
  std::vector<int> nbr((std::size_t)nbtotal * 27, -1);
  for (int n = 0; n < nbtotal; ++n) {
    MeshBlock *pmb = blocks_[n];
    SearchAndSetNeighbors(pmb);
    std::vector<Real> cx[3];
    for (int d = 0; d < 3; ++d) CoarseCellCentres(pmb, this, d, &cx[d]);
    AthkCheck(athk_refine_set_block_coords(athk, pmb->lid, pmb->pcoord->x1v.data(), pmb->pcoord->x2v.data(),
                                           pmb->pcoord->x3v.data(), cx[0].data(), cx[1].data(), cx[2].data()),
              who);
    for (const NeighborBlock &nb : pmb->neighbor) {
      int &e = nbr[(std::size_t)n * 27 + (nb.ox3 + 1) * 9 + (nb.ox2 + 1) * 3 + (nb.ox1 + 1)];
      e = (nb.level == pmb->loc.level) ? nb.lid : ATHK_NBR_REFINED;
    }
  }
  AthkCheck(athk_set_neighbors(athk, nbr.data()), who);

  std::vector<athk_box_op> ph[ATHK_REFINE_NPHASES];
  for (int n = 0; n < nbtotal; ++n) {
    MeshBlock *pmb = blocks_[n];
    const int t = pmb->lid, L = pmb->loc.level;
    const int is = pmb->is, ie = pmb->ie, js = pmb->js, je = pmb->je, ks = pmb->ks, ke = pmb->ke;
    const int nx1 = pmb->block_size.nx1, nx2 = pmb->block_size.nx2, nx3 = pmb->block_size.nx3;
    const int cis = ng, cie = cis + nx1 / 2 - 1, cjs = ng, cje = cjs + nx2 / 2 - 1, cks = ng, cke = cks + nx3 / 2 - 1;
    const std::int64_t lx1 = pmb->loc.lx1, lx2 = pmb->loc.lx2, lx3 = pmb->loc.lx3;
    // LoadBoundaryBufferToCoarser's box on a block for neighbour offset (o1,o2,o3): bvals_cc.cpp:309-331
    auto to_coarser_box = [&](int o1, int o2, int o3) {
      const int cn = ng - 1;
      Box b;
      b.si = (o1 > 0) ? (cie - cn) : cis; b.ei = (o1 < 0) ? (cis + cn) : cie;
      b.sj = (o2 > 0) ? (cje - cn) : cjs; b.ej = (o2 < 0) ? (cjs + cn) : cje;
      b.sk = (o3 > 0) ? (cke - cn) : cks; b.ek = (o3 < 0) ? (cks + cn) : cke;
      return b;
    };
    for (const NeighborBlock &nb : pmb->neighbor) {
      const int o1 = nb.ox1, o2 = nb.ox2, o3 = nb.ox3;
      if (nb.level < L) {
        // ---- phase 1: what this block restricts for its coarser neighbour (also its own stencil)
        ph[1].push_back(MakeOp(1, t, t, 0, to_coarser_box(o1, o2, o3), 0, 0, 0));
        // ---- phase 2: SetBoundaryFromCoarser (bvals_cc.cpp:424-480) <- LoadBoundaryBufferToFiner of the
        // coarser block (:335-378), which sees this block at (-o) with fi = this block's parities
        Box d;
        if (o1 == 0) { d.si = cis; d.ei = cie; if ((lx1 & 1LL) == 0LL) d.ei += cng; else d.si -= cng; }
        else if (o1 > 0) { d.si = cie + 1; d.ei = cie + cng; }
        else { d.si = cis - cng; d.ei = cis - 1; }
        if (o2 == 0) { d.sj = cjs; d.ej = cje; if ((lx2 & 1LL) == 0LL) d.ej += cng; else d.sj -= cng; }
        else if (o2 > 0) { d.sj = cje + 1; d.ej = cje + cng; }
        else { d.sj = cjs - cng; d.ej = cjs - 1; }
        if (o3 == 0) { d.sk = cks; d.ek = cke; if ((lx3 & 1LL) == 0LL) d.ek += cng; else d.sk -= cng; }
        else if (o3 > 0) { d.sk = cke + 1; d.ek = cke + cng; }
        else { d.sk = cks - cng; d.ek = cks - 1; }
        {
          const int p1 = -o1, p2 = -o2, p3 = -o3;   // as the coarser block sees this one
          const int par1 = (int)(lx1 & 1LL), par2 = (int)(lx2 & 1LL), par3 = (int)(lx3 & 1LL);
          int fi1 = 0, fi2 = 0;   // SearchAndSetNeighbors' (f1, f2) of this block in the coarser one's list
          const int nz = (p1 != 0) + (p2 != 0) + (p3 != 0);
          if (nz == 1) {
            if (p1 != 0) { fi1 = par2; fi2 = par3; }
            else if (p2 != 0) { fi1 = par1; fi2 = par3; }
            else { fi1 = par1; fi2 = par2; }
          } else if (nz == 2) {
            if (p3 == 0) fi1 = par3;
            else if (p2 == 0) fi1 = par2;
            else fi1 = par1;
          }
          const int cn = cng - 1;
          Box s;
          s.si = (p1 > 0) ? (ie - cn) : is; s.ei = (p1 < 0) ? (is + cn) : ie;
          s.sj = (p2 > 0) ? (je - cn) : js; s.ej = (p2 < 0) ? (js + cn) : je;
          s.sk = (p3 > 0) ? (ke - cn) : ks; s.ek = (p3 < 0) ? (ks + cn) : ke;
          if (p1 == 0) { if (fi1 == 1) s.si += nx1 / 2 - cng; else s.ei -= nx1 / 2 - cng; }
          if (p2 == 0) {
            if (p1 != 0) { if (fi1 == 1) s.sj += nx2 / 2 - cng; else s.ej -= nx2 / 2 - cng; }
            else { if (fi2 == 1) s.sj += nx2 / 2 - cng; else s.ej -= nx2 / 2 - cng; }
          }
          if (p3 == 0) {
            if (p1 != 0 && p2 != 0) { if (fi1 == 1) s.sk += nx3 / 2 - cng; else s.ek -= nx3 / 2 - cng; }
            else { if (fi2 == 1) s.sk += nx3 / 2 - cng; else s.ek -= nx3 / 2 - cng; }
          }
          if (s.count() != d.count() || (s.ei - s.si) != (d.ei - d.si) || (s.ej - s.sj) != (d.ej - d.sj)) {
            std::stringstream msg;
            msg << "### FATAL ERROR in " << who << std::endl << "to-finer boxes of blocks " << nb.gid << " -> " << n
                << " disagree" << std::endl;
            ATHENA_ERROR(msg);
          }
          ph[2].push_back(MakeOp(3, t, nb.lid, 0, d, s.si, s.sj, s.sk));
        }
        // ---- phases 3-6: ProlongateBoundaries for this coarser neighbour (bvals_refine.cpp:151-260)
        const int nis = std::max(o1 - 1, -1), nie = std::min(o1 + 1, 1), njs = std::max(o2 - 1, -1),
                  nje = std::min(o2 + 1, 1), nks = std::max(o3 - 1, -1), nke = std::min(o3 + 1, 1);
        for (int nk = nks; nk <= nke; ++nk)
          for (int nj = njs; nj <= nje; ++nj)
            for (int ni = nis; ni <= nie; ++ni) {
              const int ntype = std::abs(ni) + std::abs(nj) + std::abs(nk);
              if (ntype == 0 || pmb->nblevel[nk + 1][nj + 1][ni + 1] != L) continue;
              Box r;   // RestrictGhostCellsOnSameLevel, :263-300
              if (ni == 0) { r.si = cis; r.ei = cie; if (o1 == 1) r.si = cie; else if (o1 == -1) r.ei = cis; }
              else if (ni == 1) { r.si = cie + 1; r.ei = cie + 1; }
              else { r.si = cis - 1; r.ei = cis - 1; }
              if (nj == 0) { r.sj = cjs; r.ej = cje; if (o2 == 1) r.sj = cje; else if (o2 == -1) r.ej = cjs; }
              else if (nj == 1) { r.sj = cje + 1; r.ej = cje + 1; }
              else { r.sj = cjs - 1; r.ej = cjs - 1; }
              if (nk == 0) { r.sk = cks; r.ek = cke; if (o3 == 1) r.sk = cke; else if (o3 == -1) r.ek = cks; }
              else if (nk == 1) { r.sk = cke + 1; r.ek = cke + 1; }
              else { r.sk = cks - 1; r.ek = cks - 1; }
              ph[3].push_back(MakeOp(1, t, t, 0, r, 0, 0, 0));
            }
        const int cn = cng - 1;
        Box p;
        if (o1 == 0) { p.si = cis; p.ei = cie; if ((lx1 & 1LL) == 0LL) p.ei += cn; else p.si -= cn; }
        else if (o1 > 0) { p.si = cie + 1; p.ei = cie + cn; }
        else { p.si = cis - cn; p.ei = cis - 1; }
        if (o2 == 0) { p.sj = cjs; p.ej = cje; if ((lx2 & 1LL) == 0LL) p.ej += cn; else p.sj -= cn; }
        else if (o2 > 0) { p.sj = cje + 1; p.ej = cje + cn; }
        else { p.sj = cjs - cn; p.ej = cjs - 1; }
        if (o3 == 0) { p.sk = cks; p.ek = cke; if ((lx3 & 1LL) == 0LL) p.ek += cn; else p.sk -= cn; }
        else if (o3 > 0) { p.sk = cke + 1; p.ek = cke + cn; }
        else { p.sk = cks - cn; p.ek = cks - 1; }
        // ApplyPhysicalBoundariesOnCoarseLevel: W(U) over the box plus one cell where a neighbour
        // exists (:400-443; every face has one on a periodic mesh)
        int f1m = 1, f1p = 1, f2m = 1, f2p = 1, f3m = 1, f3p = 1;
        if (o1 == 0) { f1m = pmb->nblevel[1][1][0] != -1; f1p = pmb->nblevel[1][1][2] != -1; }
        if (o2 == 0) { f2m = pmb->nblevel[1][0][1] != -1; f2p = pmb->nblevel[1][2][1] != -1; }
        if (o3 == 0) { f3m = pmb->nblevel[0][1][1] != -1; f3p = pmb->nblevel[2][1][1] != -1; }
        Box c2p = {p.si - f1m, p.ei + f1p, p.sj - f2m, p.ej + f2p, p.sk - f3m, p.ek + f3p};
        ph[4].push_back(MakeOp(4, t, t, 0, c2p, 0, 0, 0));
        // ... then the physical boundary functions on the coarse primitives, one ghost layer, on the
        // faces this neighbour's box touches (:445-530); reflecting / outflow here (phase 7, run
        // between phases 4 and 5)
        {
          const int off[3] = {o1, o2, o3};
          const int clo[3] = {cis, cjs, cks}, chi[3] = {cie, cje, cke};
          for (int d = 0; d < 3; ++d) {
            if (off[d] != 0) continue;
            for (int side = 0; side < 2; ++side) {
              const BoundaryFlag bf = pmb->block_bcs[2 * d + side];
              if (bf != BoundaryFlag::reflect && bf != BoundaryFlag::outflow) continue;
              Box bb = p;
              int so[3] = {0, 0, 0};
              const int ghost = side ? chi[d] + 1 : clo[d] - 1;
              so[d] = side ? chi[d] : clo[d];   // (ngh = 1: mirror image and outflow source coincide)
              if (d == 0) bb.si = bb.ei = ghost;
              if (d == 1) bb.sj = bb.ej = ghost;
              if (d == 2) bb.sk = bb.ek = ghost;
              ph[7].push_back(MakeOp(8, t, bf == BoundaryFlag::reflect ? 1 : 2, d, bb, so[0], so[1], so[2]));
            }
          }
        }
        ph[5].push_back(MakeOp(5, t, t, 0, p, 0, 0, 0));
        Box f = {(p.si - cis) * 2 + is, (p.ei - cis) * 2 + is + 1, (p.sj - cjs) * 2 + js, (p.ej - cjs) * 2 + js + 1,
                 (p.sk - cks) * 2 + ks, (p.ek - cks) * 2 + ks + 1};
        ph[6].push_back(MakeOp(6, t, t, 0, f, 0, 0, 0));
      } else if (nb.level > L) {
        // ---- phase 2: SetBoundaryFromFiner (bvals_cc.cpp:485-560) <- the finer block's restricted box
        Box d;
        if (o1 == 0) { d.si = is; d.ei = ie; if (nb.fi1 == 1) d.si += nx1 / 2; else d.ei -= nx1 / 2; }
        else if (o1 > 0) { d.si = ie + 1; d.ei = ie + ng; }
        else { d.si = is - ng; d.ei = is - 1; }
        if (o2 == 0) {
          d.sj = js; d.ej = je;
          if (o1 != 0) { if (nb.fi1 == 1) d.sj += nx2 / 2; else d.ej -= nx2 / 2; }
          else { if (nb.fi2 == 1) d.sj += nx2 / 2; else d.ej -= nx2 / 2; }
        } else if (o2 > 0) { d.sj = je + 1; d.ej = je + ng; }
        else { d.sj = js - ng; d.ej = js - 1; }
        if (o3 == 0) {
          d.sk = ks; d.ek = ke;
          if (o1 != 0 && o2 != 0) { if (nb.fi1 == 1) d.sk += nx3 / 2; else d.ek -= nx3 / 2; }
          else { if (nb.fi2 == 1) d.sk += nx3 / 2; else d.ek -= nx3 / 2; }
        } else if (o3 > 0) { d.sk = ke + 1; d.ek = ke + ng; }
        else { d.sk = ks - ng; d.ek = ks - 1; }
        const Box s = to_coarser_box(-o1, -o2, -o3);   // (block sizes are equal on every level)
        if (s.count() != d.count() || (s.ei - s.si) != (d.ei - d.si) || (s.ej - s.sj) != (d.ej - d.sj)) {
          std::stringstream msg;
          msg << "### FATAL ERROR in " << who << std::endl << "to-coarser boxes of blocks " << nb.gid << " -> " << n
              << " disagree" << std::endl;
          ATHENA_ERROR(msg);
        }
        ph[2].push_back(MakeOp(2, t, nb.lid, 0, d, s.si, s.sj, s.sk));
        // ---- phase 0: flux correction from this finer face neighbour (flux_correction_cc.cpp:222-275)
        if (nb.type == 0) {
          const int dir = nb.fid / 2, outer = nb.fid & 1;
          Box fb = {is, ie, js, je, ks, ke};
          int so[3] = {is, js, ks};
          if (dir == 0) {
            fb.si = fb.ei = is + (ie - is) * outer + outer;
            if (nb.fi1 == 0) fb.ej -= nx2 / 2; else fb.sj += nx2 / 2;
            if (nb.fi2 == 0) fb.ek -= nx3 / 2; else fb.sk += nx3 / 2;
            so[0] = outer ? is : ie + 1;   // the finer block's face towards this one
          } else if (dir == 1) {
            fb.sj = fb.ej = js + (je - js) * outer + outer;
            if (nb.fi1 == 0) fb.ei -= nx1 / 2; else fb.si += nx1 / 2;
            if (nb.fi2 == 0) fb.ek -= nx3 / 2; else fb.sk += nx3 / 2;
            so[1] = outer ? js : je + 1;
          } else {
            fb.sk = fb.ek = ks + (ke - ks) * outer + outer;
            if (nb.fi1 == 0) fb.ei -= nx1 / 2; else fb.si += nx1 / 2;
            if (nb.fi2 == 0) fb.ej -= nx2 / 2; else fb.sj += nx2 / 2;
            so[2] = outer ? ks : ke + 1;
          }
          ph[0].push_back(MakeOp(7, t, nb.lid, dir, fb, so[0], so[1], so[2]));
        }
      }
    }
  }
  for (int p = 0; p < ATHK_REFINE_NPHASES; ++p)
    AthkCheck(athk_refine_set_ops(athk, p, ph[p].data(), (int)ph[p].size()), who);
}
