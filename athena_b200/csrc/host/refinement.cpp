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
//   face-centred field: Load / Set      bvals/fc/bvals_fc.cpp:344-976 (same level on a refined mesh, to / from
//                                       coarser and finer), RestrictGhostCellsOnSameLevel / ProlongateGhostCells
//                                       bvals/bvals_refine.cpp:318-356, 588-637
//   EMF correction across levels        bvals/fc/flux_correction_fc.cpp:315-570, 944-1190, 1214-1541;
//                                       CountFineEdges bvals/fc/bvals_fc.cpp:1129-1197
#include <algorithm>
#include <cmath>
#include <cstdlib>

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
    nb.gid = t->gid_; nb.level = t->loc_.level;
    nb.rank = ranklist[t->gid_]; nb.lid = t->gid_ - nslist[nb.rank];
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
  o.peer = 0;
  return o;
}
// a box whose source block is the neighbour nb of a block on GPU `rank`
athk_box_op FromNeighbor(athk_box_op o, const NeighborBlock &nb, int rank) {
  o.peer = (nb.rank == rank) ? 0 : nb.rank + 1;
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
  for (athk_ctx *a : ctx) AthkCheck(athk_refine_enable(a), who);
  for (athk_ctx *a : ctx) AthkCheck(athk_refine_set_peers(a, ngpu, ctx.data()), who);
  const int ng = NGHOST, cng = (NGHOST + 1) / 2 + 1;
  for (int n = 0; n < nbtotal; ++n) {
    MeshBlock *pmb = blocks_[n];
    SearchAndSetNeighbors(pmb);
    std::vector<Real> cx[3];
    for (int d = 0; d < 3; ++d) CoarseCellCentres(pmb, this, d, &cx[d]);
    AthkCheck(athk_refine_set_block_coords(ctx_of(pmb), pmb->lid, pmb->pcoord->x1v.data(), pmb->pcoord->x2v.data(),
                                           pmb->pcoord->x3v.data(), cx[0].data(), cx[1].data(), cx[2].data()),
              who);
  }
  // neighbour tables (same level: block index or remote link; other level: ATHK_NBR_REFINED), links and
  // peer-to-peer buffers of the same-level messages between GPUs (multigpu.cpp)
  SetupRemoteLinks();

  // lists[rank][list]: the boxes of the MeshBlocks GPU `rank` owns
  std::vector<std::vector<std::vector<athk_box_op>>> lists(ngpu, std::vector<std::vector<athk_box_op>>(ATHK_REFINE_NPHASES));
  for (int n = 0; n < nbtotal; ++n) {
    MeshBlock *pmb = blocks_[n];
    std::vector<std::vector<athk_box_op>> &ph = lists[pmb->rank];
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
          ph[2].push_back(FromNeighbor(MakeOp(3, t, nb.lid, 0, d, s.si, s.sj, s.sk), nb, pmb->rank));
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
              if (bf == BoundaryFlag::user) {   // DispatchBoundaryFunctions(..., cis, cie, sj, ej, sk, ek, 1, ...)
                CoarseUserBC ub = {n, 2 * d + side, p.si, p.ei, p.sj, p.ej, p.sk, p.ek};
                if (d == 0) { ub.il = clo[0]; ub.iu = chi[0]; }
                if (d == 1) { ub.jl = clo[1]; ub.ju = chi[1]; }
                if (d == 2) { ub.kl = clo[2]; ub.ku = chi[2]; }
                coarse_user_bcs_.push_back(ub);
                continue;
              }
              if (bf != BoundaryFlag::reflect && bf != BoundaryFlag::outflow) continue;
              Box bb = p;
              int so[3] = {0, 0, 0};
              const int ghost = side ? chi[d] + 1 : clo[d] - 1;
              so[d] = side ? chi[d] : clo[d];   // (ngh = 1: mirror image and outflow source coincide)
              if (d == 0) bb.si = bb.ei = ghost;
              if (d == 1) bb.sj = bb.ej = ghost;
              if (d == 2) bb.sk = bb.ek = ghost;
              ph[7].push_back(MakeOp(8, t, bf == BoundaryFlag::reflect ? 1 : 2, d, bb, so[0], so[1], so[2]));
              if (MAGNETIC_FIELDS_ENABLED) {   // ... and on coarse_b_ (bvals/fc/reflect_fc.cpp, outflow_fc.cpp, ngh = 1)
                const bool refl = (bf == BoundaryFlag::reflect);
                for (int c = 0; c < 3; ++c) {
                  int lo[3] = {p.si, p.sj, p.sk}, hi[3] = {p.ei, p.ej, p.ek}, fso[3] = {0, 0, 0};
                  bool negate = false;
                  if (c == d) {   // the normal component: faces clo-1 / chi+2 from clo+1 / chi (reflected) or clo / chi+1
                    lo[d] = hi[d] = side ? chi[d] + 2 : clo[d] - 1;
                    fso[d] = refl ? (side ? chi[d] : clo[d] + 1) : (side ? chi[d] + 1 : clo[d]);
                    negate = refl;
                  } else {        // transverse components: the ghost cell layer from the first active one
                    lo[d] = hi[d] = side ? chi[d] + 1 : clo[d] - 1;
                    fso[d] = side ? chi[d] : clo[d];
                    hi[c] += 1;
                  }
                  athk_box_op o = MakeOp(18, t, negate ? 1 : 2, c, Box{lo[0], hi[0], lo[1], hi[1], lo[2], hi[2]}, fso[0], fso[1], fso[2]);
                  o.pad = d;
                  ph[7].push_back(o);
                }
              }
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
        ph[2].push_back(FromNeighbor(MakeOp(2, t, nb.lid, 0, d, s.si, s.sj, s.sk), nb, pmb->rank));
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
          ph[0].push_back(FromNeighbor(MakeOp(7, t, nb.lid, dir, fb, so[0], so[1], so[2]), nb, pmb->rank));
        }
      }
    }
  }
  if (NSCALARS > 0) for (int rk = 0; rk < ngpu; ++rk) {
    std::vector<std::vector<athk_box_op>> &ph = lists[rk];
    // passive scalars travel with the hydro variables (MeshRefinement::pvars_cc_: (s, coarse_s_), then
    // (r, coarse_r_) for the prolongation, bvals_refine.cpp:560-575): every cell-centred box gets a twin
    // acting on them (pad = 1); W(U) / U(W) boxes handle them in place.  The flux correction of
    // s_flux is its own list (66): the scalar fluxes are formed after the hydro ones.
    for (int list : {1, 2, 3, 5, 7}) {
      const std::size_t n0 = ph[list].size();
      for (std::size_t q = 0; q < n0; ++q) {
        athk_box_op o = ph[list][q];
        if (o.kind == 1 || o.kind == 2 || o.kind == 3 || o.kind == 5 || o.kind == 8) { o.pad = 1; ph[list].push_back(o); }
      }
    }
    for (athk_box_op o : ph[0]) { o.pad = 1; ph[66].push_back(o); }
  }
  if (MAGNETIC_FIELDS_ENABLED) SetupFieldRefinement(&lists);
  for (int rk = 0; rk < ngpu; ++rk)
    for (int p = 0; p < ATHK_REFINE_NPHASES; ++p)
      AthkCheck(athk_refine_set_ops(ctx[rk], p, lists[rk][p].data(), (int)lists[rk][p].size()), who);
}

// The user-enrolled boundary functions on coarse_prim_ / coarse_b_ with the coarse coordinates, one ghost
// layer (bvals_refine.cpp:445-530 -> bvals.cpp:673-814).  The functions run on the host, as on the fine
// level: the block's coarse arrays come to the host mirrors, every enrolled face touched by one of the
// block's coarser neighbours is served in the reference's order, and the arrays go back.
void Mesh::ApplyUserBoundariesOnCoarseLevel() {
  const char *who = "Mesh::ApplyUserBoundariesOnCoarseLevel";
  std::size_t q = 0;
  while (q < coarse_user_bcs_.size()) {
    const int gid = coarse_user_bcs_[q].gid;
    MeshBlock *pmb = blocks_[gid];
    athk_ctx *a = ctx_of(pmb);
    const int ng = NGHOST;
    const int cn1 = pmb->block_size.nx1 / 2 + 2 * ng, cn2 = pmb->block_size.nx2 / 2 + 2 * ng, cn3 = pmb->block_size.nx3 / 2 + 2 * ng;
    if (pmb->pcoarsec == nullptr) {
      pmb->pcoarsec = new Coordinates(pmb, this, true);
      pmb->coarse_prim_.NewAthenaArray(5, cn3, cn2, cn1);   // (the device keeps five slots for every EOS)
      if (MAGNETIC_FIELDS_ENABLED) {
        pmb->coarse_b_.x1f.NewAthenaArray(cn3, cn2, cn1 + 1);
        pmb->coarse_b_.x2f.NewAthenaArray(cn3, cn2 + 1, cn1);
        pmb->coarse_b_.x3f.NewAthenaArray(cn3 + 1, cn2, cn1);
      }
    }
    double *b1 = MAGNETIC_FIELDS_ENABLED ? pmb->coarse_b_.x1f.data() : nullptr;
    double *b2 = MAGNETIC_FIELDS_ENABLED ? pmb->coarse_b_.x2f.data() : nullptr;
    double *b3 = MAGNETIC_FIELDS_ENABLED ? pmb->coarse_b_.x3f.data() : nullptr;
    AthkCheck(athk_refine_coarse_transfer(a, pmb->lid, 0, pmb->coarse_prim_.data(), b1, b2, b3), who);
    for (; q < coarse_user_bcs_.size() && coarse_user_bcs_[q].gid == gid; ++q) {
      const CoarseUserBC &t = coarse_user_bcs_[q];
      if (BoundaryFunction_[t.face] == nullptr) continue;
      BoundaryFunction_[t.face](pmb, pmb->pcoarsec, pmb->coarse_prim_, pmb->coarse_b_, bc_time_, bc_dt_, t.il, t.iu, t.jl,
                                t.ju, t.kl, t.ku, 1);
    }
    AthkCheck(athk_refine_coarse_transfer(a, pmb->lid, 1, pmb->coarse_prim_.data(), b1, b2, b3), who);
  }
}

void Mesh::SyncAll() { for (athk_ctx *a : ctx) AthkCheck(athk_sync(a), "Mesh::SyncAll"); }

// Boxes may read blocks of another GPU: with several GPUs every list starts after all GPUs have
// finished what was enqueued before it (host-side; refined meshes are not the throughput path)
void Mesh::RunRefineList(int list) {
  if (ngpu > 1) SyncAll();
  for (athk_ctx *a : ctx) AthkCheck(athk_refine_run(a, list, 0), "Mesh::RunRefineList");
}

//----------------------------------------------------------------------------------------
// Face-centred field and EMF correction on a refined mesh.  Lists (see Mesh::ExchangeGhosts and
// TimeIntegratorTaskList::DoTaskListOneStage for the order they run in):
//   1  (shared with the cell-centred boxes) RestrictField for coarser neighbours
//   8, 9, 10  SetBoundarySameLevel / SetBoundaryFromFiner of face, edge, corner neighbours: the reference
//      sets neighbour after neighbour in that order and edge / corner neighbours overwrite shared faces
//   11, 64, 65 SetBoundaryFromCoarser of face, edge, corner neighbours (in that order: two coarser blocks
//      both hand over the face they share, and their copies of it may differ in the last bits; after 8-10:
//      it overwrites boundary faces of the coarse buffer that the restriction for OTHER neighbours
//      produced and lists 8-10 still read)
//   3  (shared) RestrictGhostCellsOnSameLevel, 12 ProlongateSharedField, 13 ProlongateInternalField
//   14 ClearCoarseFluxBoundary, 16 + slot SetFluxBoundaryFromFiner of the slot-th face / edge neighbour
//      (accumulation order = neighbour order), 15 AverageFluxBoundary where finer EMFs were summed
void Mesh::SetupFieldRefinement(std::vector<std::vector<std::vector<athk_box_op>>> *lists) {
  const char *who = "Mesh::SetupFieldRefinement";
  const int ng = NGHOST, cng = (NGHOST + 1) / 2 + 1;
  for (int n = 0; n < nbtotal; ++n) {
    MeshBlock *pmb = blocks_[n];
    std::vector<std::vector<athk_box_op>> &ph = (*lists)[pmb->rank];
    const int t = pmb->lid, L = pmb->loc.level;
    const int s_[3] = {pmb->is, pmb->js, pmb->ks}, e_[3] = {pmb->ie, pmb->je, pmb->ke};
    const int nx[3] = {pmb->block_size.nx1, pmb->block_size.nx2, pmb->block_size.nx3};
    const int cs[3] = {ng, ng, ng}, ce[3] = {ng + nx[0] / 2 - 1, ng + nx[1] / 2 - 1, ng + nx[2] / 2 - 1};
    const std::int64_t lx[3] = {pmb->loc.lx1, pmb->loc.lx2, pmb->loc.lx3};
    auto mk = [](const int lo[3], const int hi[3]) { return Box{lo[0], hi[0], lo[1], hi[1], lo[2], hi[2]}; };
    // which of (fi1, fi2) halves direction d of a neighbour at offset o (bvals_fc.cpp:820-860)
    auto fi_of = [](const int o[3], int d, int fi1, int fi2) {
      if (d == 0) return fi1;
      if (d == 1) return (o[0] != 0) ? fi1 : fi2;
      return (o[0] != 0 && o[1] != 0) ? fi1 : fi2;
    };
    // LoadBoundaryBufferToCoarser (bvals_fc.cpp:403-470): coarse-index box of component c the block restricts
    // for a coarser neighbour at offset p
    auto to_coarser_fc = [&](int c, const int p[3], bool face) {
      int lo[3], hi[3];
      for (int d = 0; d < 3; ++d) {
        if (d == c) {
          if (p[d] == 0) { lo[d] = cs[d]; hi[d] = ce[d] + 1; }
          else if (p[d] > 0) { lo[d] = ce[d] - ng + 1; hi[d] = ce[d]; }
          else { lo[d] = cs[d] + 1; hi[d] = cs[d] + ng; }
          if (!face) { if (p[d] > 0) hi[d]++; else if (p[d] < 0) lo[d]--; }
        } else {
          if (p[d] == 0) { lo[d] = cs[d]; hi[d] = ce[d]; }
          else if (p[d] > 0) { lo[d] = ce[d] - ng + 1; hi[d] = ce[d]; }
          else { lo[d] = cs[d]; hi[d] = cs[d] + ng - 1; }
        }
      }
      return mk(lo, hi);
    };
    for (const NeighborBlock &nb : pmb->neighbor) {
      const int o[3] = {nb.ox1, nb.ox2, nb.ox3};
      const int po[3] = {-nb.ox1, -nb.ox2, -nb.ox3};
      const bool face = (nb.type == 0);
      const int list_set = 8 + nb.type;
      if (nb.level == L) {
        // ---- SetBoundarySameLevel, multilevel form (bvals_fc.cpp:583-680)
        for (int c = 0; c < 3; ++c) {
          int lo[3], hi[3];
          for (int d = 0; d < 3; ++d) {
            if (d == c) {
              if (o[d] == 0) { lo[d] = s_[d]; hi[d] = e_[d] + 1; }
              else if (o[d] > 0) { lo[d] = e_[d] + 2; hi[d] = e_[d] + ng + 1; }
              else { lo[d] = s_[d] - ng; hi[d] = s_[d] - 1; }
              if (!face) { if (o[d] > 0) lo[d]--; else if (o[d] < 0) hi[d]++; }
            } else {
              if (o[d] == 0) { lo[d] = s_[d]; hi[d] = e_[d]; }
              else if (o[d] > 0) { lo[d] = e_[d] + 1; hi[d] = e_[d] + ng; }
              else { lo[d] = s_[d] - ng; hi[d] = s_[d] - 1; }
            }
          }
          ph[list_set].push_back(FromNeighbor(MakeOp(9, t, nb.lid, c, mk(lo, hi), lo[0] - o[0] * nx[0],
                                                     lo[1] - o[1] * nx[1], lo[2] - o[2] * nx[2]), nb, pmb->rank));
        }
      } else if (nb.level > L) {
        // ---- SetBoundaryFromFiner (bvals_fc.cpp:813-976) <- the finer block's restricted box
        for (int c = 0; c < 3; ++c) {
          int lo[3], hi[3];
          for (int d = 0; d < 3; ++d) {
            const int fi = fi_of(o, d, nb.fi1, nb.fi2);
            if (d == c) {
              if (o[d] == 0) { lo[d] = s_[d]; hi[d] = e_[d] + 1; if (fi == 1) lo[d] += nx[d] / 2; else hi[d] -= nx[d] / 2; }
              else if (o[d] > 0) { lo[d] = e_[d] + 2; hi[d] = e_[d] + ng + 1; }
              else { lo[d] = s_[d] - ng; hi[d] = s_[d] - 1; }
              if (!face) { if (o[d] > 0) lo[d]--; else if (o[d] < 0) hi[d]++; }
            } else {
              if (o[d] == 0) { lo[d] = s_[d]; hi[d] = e_[d]; if (fi == 1) lo[d] += nx[d] / 2; else hi[d] -= nx[d] / 2; }
              else if (o[d] > 0) { lo[d] = e_[d] + 1; hi[d] = e_[d] + ng; }
              else { lo[d] = s_[d] - ng; hi[d] = s_[d] - 1; }
            }
          }
          const Box d = mk(lo, hi), sb = to_coarser_fc(c, po, face);   // (block sizes are equal on every level)
          if (sb.count() != d.count() || (sb.ei - sb.si) != (d.ei - d.si) || (sb.ej - sb.sj) != (d.ej - d.sj)) {
            std::stringstream msg;
            msg << "### FATAL ERROR in " << who << std::endl << "to-coarser field boxes of blocks " << nb.gid << " -> " << n
                << " disagree" << std::endl;
            ATHENA_ERROR(msg);
          }
          ph[list_set].push_back(FromNeighbor(MakeOp(11, t, nb.lid, c, d, sb.si, sb.sj, sb.sk), nb, pmb->rank));
        }
      } else {
        // ---- list 1: RestrictField for this coarser neighbour (LoadBoundaryBufferToCoarser)
        for (int c = 0; c < 3; ++c) ph[1].push_back(MakeOp(10, t, t, c, to_coarser_fc(c, o, face), 0, 0, 0));
        // ---- lists 11 / 64 / 65: SetBoundaryFromCoarser (bvals_fc.cpp:690-806) <- LoadBoundaryBufferToFiner of the
        // coarser block (:476-576), which sees this block at -o with fi = this block's parities
        int fi1 = 0, fi2 = 0;
        {
          const int par[3] = {(int)(lx[0] & 1LL), (int)(lx[1] & 1LL), (int)(lx[2] & 1LL)};
          const int nz = (po[0] != 0) + (po[1] != 0) + (po[2] != 0);
          if (nz == 1) {
            if (po[0] != 0) { fi1 = par[1]; fi2 = par[2]; }
            else if (po[1] != 0) { fi1 = par[0]; fi2 = par[2]; }
            else { fi1 = par[0]; fi2 = par[1]; }
          } else if (nz == 2) {
            if (po[2] == 0) fi1 = par[2];
            else if (po[1] == 0) fi1 = par[1];
            else fi1 = par[0];
          }
        }
        const int cn = cng - 1;
        for (int c = 0; c < 3; ++c) {
          int lo[3], hi[3], slo[3], shi[3];
          for (int d = 0; d < 3; ++d) {
            const bool even = (lx[d] & 1LL) == 0LL;
            const int fi = fi_of(po, d, fi1, fi2);
            if (d == c) {
              if (o[d] == 0) { lo[d] = cs[d]; hi[d] = ce[d] + 1; if (even) hi[d] += cng; else lo[d] -= cng; }
              else if (o[d] > 0) { lo[d] = ce[d] + 1; hi[d] = ce[d] + 1 + cng; }
              else { lo[d] = cs[d] - cng; hi[d] = cs[d]; }
              if (po[d] == 0) {
                if (fi == 1) { slo[d] = s_[d] + nx[d] / 2 - cng; shi[d] = e_[d] + 1; }
                else { slo[d] = s_[d]; shi[d] = e_[d] + 1 - nx[d] / 2 + cng; }
              } else if (po[d] > 0) { slo[d] = e_[d] + 1 - cng; shi[d] = e_[d] + 1; }
              else { slo[d] = s_[d]; shi[d] = s_[d] + cng; }
            } else {
              if (o[d] == 0) { lo[d] = cs[d]; hi[d] = ce[d]; if (even) hi[d] += cng; else lo[d] -= cng; }
              else if (o[d] > 0) { lo[d] = ce[d] + 1; hi[d] = ce[d] + cng; }
              else { lo[d] = cs[d] - cng; hi[d] = cs[d] - 1; }
              if (po[d] == 0) {
                slo[d] = s_[d]; shi[d] = e_[d];
                if (fi == 1) slo[d] += nx[d] / 2 - cng; else shi[d] -= nx[d] / 2 - cng;
              } else if (po[d] > 0) { slo[d] = e_[d] - cn; shi[d] = e_[d]; }
              else { slo[d] = s_[d]; shi[d] = s_[d] + cn; }
            }
          }
          const Box d = mk(lo, hi), sb = mk(slo, shi);
          if (sb.count() != d.count() || (sb.ei - sb.si) != (d.ei - d.si) || (sb.ej - sb.sj) != (d.ej - d.sj)) {
            std::stringstream msg;
            msg << "### FATAL ERROR in " << who << std::endl << "to-finer field boxes of blocks " << nb.gid << " -> " << n
                << " disagree (component " << c << ")" << std::endl;
            ATHENA_ERROR(msg);
          }
          ph[nb.type == 0 ? 11 : 63 + nb.type].push_back(FromNeighbor(MakeOp(12, t, nb.lid, c, d, sb.si, sb.sj, sb.sk), nb, pmb->rank));
        }
        // ---- list 3: RestrictGhostCellsOnSameLevel, field part (bvals_refine.cpp:318-356)
        const int nlo[3] = {std::max(o[0] - 1, -1), std::max(o[1] - 1, -1), std::max(o[2] - 1, -1)};
        const int nhi[3] = {std::min(o[0] + 1, 1), std::min(o[1] + 1, 1), std::min(o[2] + 1, 1)};
        for (int nk = nlo[2]; nk <= nhi[2]; ++nk)
          for (int nj = nlo[1]; nj <= nhi[1]; ++nj)
            for (int ni = nlo[0]; ni <= nhi[0]; ++ni) {
              const int nn[3] = {ni, nj, nk};
              if (std::abs(ni) + std::abs(nj) + std::abs(nk) == 0 || pmb->nblevel[nk + 1][nj + 1][ni + 1] != L) continue;
              int rlo[3], rhi[3];
              for (int d = 0; d < 3; ++d) {
                if (nn[d] == 0) { rlo[d] = cs[d]; rhi[d] = ce[d]; if (o[d] == 1) rlo[d] = ce[d]; else if (o[d] == -1) rhi[d] = cs[d]; }
                else if (nn[d] == 1) { rlo[d] = rhi[d] = ce[d] + 1; }
                else { rlo[d] = rhi[d] = cs[d] - 1; }
              }
              for (int c = 0; c < 3; ++c) {
                int lo[3] = {rlo[0], rlo[1], rlo[2]}, hi[3] = {rhi[0], rhi[1], rhi[2]};
                hi[c] = rhi[c] + 1;
                // the face on the block's boundary is not restricted when the block beyond it is coarser
                int m[3] = {ni + 1, nj + 1, nk + 1};
                if (lo[c] == cs[c]) { m[c] = nn[c]; if (pmb->nblevel[m[2]][m[1]][m[0]] < L) lo[c]++; }
                m[c] = nn[c] + 1;
                if (hi[c] == ce[c] + 1) { m[c] = nn[c] + 2; if (pmb->nblevel[m[2]][m[1]][m[0]] < L) hi[c]--; }
                if (lo[c] > hi[c]) continue;
                ph[3].push_back(MakeOp(10, t, t, c, mk(lo, hi), 0, 0, 0));
              }
            }
        // ---- lists 12, 13: ProlongateGhostCells, field part (bvals_refine.cpp:588-637)
        int plo[3], phi[3];
        for (int d = 0; d < 3; ++d) {
          if (o[d] == 0) { plo[d] = cs[d]; phi[d] = ce[d]; if ((lx[d] & 1LL) == 0LL) phi[d] += cn; else plo[d] -= cn; }
          else if (o[d] > 0) { plo[d] = ce[d] + 1; phi[d] = ce[d] + cn; }
          else { plo[d] = cs[d] - cn; phi[d] = cs[d] - 1; }
        }
        for (int c = 0; c < 3; ++c) {
          int lo[3] = {plo[0], plo[1], plo[2]}, hi[3] = {phi[0], phi[1], phi[2]};
          hi[c] = phi[c] + 1;
          int m[3] = {o[0] + 1, o[1] + 1, o[2] + 1};
          m[c] = o[c];       // the block below this neighbour in direction c
          if (o[c] >= 0 && pmb->nblevel[m[2]][m[1]][m[0]] >= L) lo[c]++;
          m[c] = o[c] + 2;   // ... and above it
          if (o[c] <= 0 && pmb->nblevel[m[2]][m[1]][m[0]] >= L) hi[c]--;
          if (lo[c] > hi[c]) continue;
          ph[12].push_back(MakeOp(13, t, t, c, mk(lo, hi), 0, 0, 0));
        }
        ph[13].push_back(MakeOp(14, t, t, 0, mk(plo, phi), 0, 0, 0));
      }
    }
    // ---- EMF correction: CountFineEdges (bvals_fc.cpp:1129-1197)
    bool edge_flag[12];
    int nedge_fine[12];
    {
      int eid = 0;
      for (int pl = 0; pl < 3; ++pl)   // x1x2 (ox2 outer, ox1 inner), x1x3 (ox3, ox1), x2x3 (ox3, ox2)
        for (int oa = -1; oa <= 1; oa += 2)
          for (int ob = -1; ob <= 1; ob += 2) {
            const int as = std::max(oa - 1, -1), ae = std::min(oa + 1, 1), bs = std::max(ob - 1, -1), be = std::min(ob + 1, 1);
            int nf = 0, fl = L;
            for (int na = as; na <= ae; ++na)
              for (int nbb = bs; nbb <= be; ++nbb) {
                const int lev = (pl == 0) ? pmb->nblevel[1][na + 1][nbb + 1]
                                          : (pl == 1 ? pmb->nblevel[na + 1][1][nbb + 1] : pmb->nblevel[na + 1][nbb + 1][1]);
                if (lev > fl) { fl++; nf = 0; }
                if (lev == fl) nf++;
              }
            edge_flag[eid] = (fl == L);
            nedge_fine[eid++] = nf;
          }
    }
    const int is = s_[0], ie = e_[0], js = s_[1], je = e_[1], ks = s_[2], ke = e_[2];
    // ---- list 14 ClearCoarseFluxBoundary (:1214-1350) and list 15 AverageFluxBoundary (:1355-1541), the
    // parts that concern EMFs summed from finer blocks (the same-level average is athk_recv_flux_correction's)
    for (int f = 0; f < 6; ++f) {
      const int dir = f / 2, side = f & 1;
      const int lev = (dir == 0) ? pmb->nblevel[1][1][2 * side] : (dir == 1 ? pmb->nblevel[1][2 * side][1] : pmb->nblevel[2 * side][1][1]);
      if (lev <= L) continue;
      if (dir == 0) {
        const int i = side ? ie + 1 : is;
        ph[14].push_back(MakeOp(15, t, t, 1, Box{i, i, js, je, ks + 1, ke}, 0, 0, 0));
        ph[14].push_back(MakeOp(15, t, t, 2, Box{i, i, js + 1, je, ks, ke}, 0, 0, 0));
        const int km = ks + nx[2] / 2, jm = js + nx[1] / 2;
        ph[15].push_back(MakeOp(17, t, 2, 1, Box{i, i, js, je, km, km}, 0, 0, 0));
        ph[15].push_back(MakeOp(17, t, 2, 2, Box{i, i, jm, jm, ks, ke}, 0, 0, 0));
      } else if (dir == 1) {
        const int j = side ? je + 1 : js;
        ph[14].push_back(MakeOp(15, t, t, 0, Box{is, ie, j, j, ks + 1, ke}, 0, 0, 0));
        ph[14].push_back(MakeOp(15, t, t, 2, Box{is + 1, ie, j, j, ks, ke}, 0, 0, 0));
        const int km = ks + nx[2] / 2, im = is + nx[0] / 2;
        ph[15].push_back(MakeOp(17, t, 2, 0, Box{is, ie, j, j, km, km}, 0, 0, 0));
        ph[15].push_back(MakeOp(17, t, 2, 2, Box{im, im, j, j, ks, ke}, 0, 0, 0));
      } else {
        const int k = side ? ke + 1 : ks;
        ph[14].push_back(MakeOp(15, t, t, 0, Box{is, ie, js + 1, je, k, k}, 0, 0, 0));
        ph[14].push_back(MakeOp(15, t, t, 1, Box{is + 1, ie, js, je, k, k}, 0, 0, 0));
        const int jm = js + nx[1] / 2, im = is + nx[0] / 2;
        ph[15].push_back(MakeOp(17, t, 2, 0, Box{is, ie, jm, jm, k, k}, 0, 0, 0));
        ph[15].push_back(MakeOp(17, t, 2, 1, Box{im, im, js, je, k, k}, 0, 0, 0));
      }
    }
    for (int eid = 0; eid < 12; ++eid) {
      if (edge_flag[eid]) continue;
      Box b;
      int comp;
      if (eid < 4) { const int i = (eid & 1) ? ie + 1 : is, j = (eid & 2) ? je + 1 : js; b = Box{i, i, j, j, ks, ke}; comp = 2; }
      else if (eid < 8) { const int i = (eid & 1) ? ie + 1 : is, k = (eid & 2) ? ke + 1 : ks; b = Box{i, i, js, je, k, k}; comp = 1; }
      else { const int j = (eid & 1) ? je + 1 : js, k = (eid & 2) ? ke + 1 : ks; b = Box{is, ie, j, j, k, k}; comp = 0; }
      ph[14].push_back(MakeOp(15, t, t, comp, b, 0, 0, 0));
      if (nedge_fine[eid] != 1) ph[15].push_back(MakeOp(17, t, nedge_fine[eid], comp, b, 0, 0, 0));
    }
    // ---- lists 16 + slot: SetFluxBoundaryFromFiner (:944-1165) of the restricted surface EMFs the finer
    // neighbour loads (LoadFluxBoundaryBufferToCoarser :315-570)
    for (const NeighborBlock &nb : pmb->neighbor) {
      if (nb.level != L + 1 || nb.type > 1) continue;
      const int h1 = nx[0] / 2, h2 = nx[1] / 2, h3 = nx[2] / 2;
      if (nb.type == 0) {
        const int dir = nb.fid / 2, outer = nb.fid & 1;
        const int list = 16 + nb.fid * 4 + nb.fi2 * 2 + nb.fi1;
        int il = is, iu = ie, jl = js, ju = je, kl = ks, ku = ke;
        if (dir == 0) {
          const int i = outer ? ie + 1 : is, fi = outer ? is : ie + 1;
          if (nb.fi1 == 0) ju = js + h2 - 1; else jl = js + h2;
          if (nb.fi2 == 0) ku = ks + h3 - 1; else kl = ks + h3;
          ph[list].push_back(FromNeighbor(MakeOp(16, t, nb.lid, 1, Box{i, i, jl, ju, kl, ku + 1}, fi, js, ks), nb, pmb->rank));
          ph[list].push_back(FromNeighbor(MakeOp(16, t, nb.lid, 2, Box{i, i, jl, ju + 1, kl, ku}, fi, js, ks), nb, pmb->rank));
        } else if (dir == 1) {
          const int j = outer ? je + 1 : js, fj = outer ? js : je + 1;
          if (nb.fi1 == 0) iu = is + h1 - 1; else il = is + h1;
          if (nb.fi2 == 0) ku = ks + h3 - 1; else kl = ks + h3;
          ph[list].push_back(FromNeighbor(MakeOp(16, t, nb.lid, 0, Box{il, iu, j, j, kl, ku + 1}, is, fj, ks), nb, pmb->rank));
          ph[list].push_back(FromNeighbor(MakeOp(16, t, nb.lid, 2, Box{il, iu + 1, j, j, kl, ku}, is, fj, ks), nb, pmb->rank));
        } else {
          const int k = outer ? ke + 1 : ks, fk = outer ? ks : ke + 1;
          if (nb.fi1 == 0) iu = is + h1 - 1; else il = is + h1;
          if (nb.fi2 == 0) ju = js + h2 - 1; else jl = js + h2;
          ph[list].push_back(FromNeighbor(MakeOp(16, t, nb.lid, 0, Box{il, iu, jl, ju + 1, k, k}, is, js, fk), nb, pmb->rank));
          ph[list].push_back(FromNeighbor(MakeOp(16, t, nb.lid, 1, Box{il, iu + 1, jl, ju, k, k}, is, js, fk), nb, pmb->rank));
        }
      } else {
        int eid;
        if (nb.ox3 == 0) eid = ((nb.ox1 + 1) >> 1) | ((nb.ox2 + 1) & 2);
        else if (nb.ox2 == 0) eid = 4 + (((nb.ox1 + 1) >> 1) | ((nb.ox3 + 1) & 2));
        else eid = 8 + (((nb.ox2 + 1) >> 1) | ((nb.ox3 + 1) & 2));
        const int list = 16 + 24 + eid * 2 + nb.fi1;
        if (eid < 4) {
          const int i = (eid & 1) ? ie + 1 : is, j = (eid & 2) ? je + 1 : js;
          const int fi = (eid & 1) ? is : ie + 1, fj = (eid & 2) ? js : je + 1;
          int kl = ks, ku = ke;
          if (nb.fi1 == 0) ku = ks + h3 - 1; else kl = ks + h3;
          ph[list].push_back(FromNeighbor(MakeOp(16, t, nb.lid, 2, Box{i, i, j, j, kl, ku}, fi, fj, ks), nb, pmb->rank));
        } else if (eid < 8) {
          const int i = (eid & 1) ? ie + 1 : is, k = (eid & 2) ? ke + 1 : ks;
          const int fi = (eid & 1) ? is : ie + 1, fk = (eid & 2) ? ks : ke + 1;
          int jl = js, ju = je;
          if (nb.fi1 == 0) ju = js + h2 - 1; else jl = js + h2;
          ph[list].push_back(FromNeighbor(MakeOp(16, t, nb.lid, 1, Box{i, i, jl, ju, k, k}, fi, js, fk), nb, pmb->rank));
        } else {
          const int j = (eid & 1) ? je + 1 : js, k = (eid & 2) ? ke + 1 : ks;
          const int fj = (eid & 1) ? js : je + 1, fk = (eid & 2) ? ks : ke + 1;
          int il = is, iu = ie;
          if (nb.fi1 == 0) iu = is + h1 - 1; else il = is + h1;
          ph[list].push_back(FromNeighbor(MakeOp(16, t, nb.lid, 0, Box{il, iu, j, j, k, k}, is, fj, fk), nb, pmb->rank));
        }
      }
    }
  }
}
