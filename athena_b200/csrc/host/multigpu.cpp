// multigpu.cpp -- Mesh::SetupRemoteLinks: neighbour tables of every GPU's MeshBlocks, the links to
// neighbours that live on another GPU, and the peer-to-peer buffers their messages travel through.
//
// Reference: BoundaryBase::SearchAndSetNeighbors (bvals/bvals_base.cpp:301-372) finds the neighbour
// of every (block, offset) and records its rank; same-rank neighbours are served by
// CopyVariableBufferSameProcess (bvals_var.cpp:108-119), other ranks by MPI messages
// (bvals_var.cpp:212-262, flux_correction_fc.cpp:623-654).  Here every "rank" is a GPU of this
// process: a neighbour on the same GPU is a local block index in the table, a neighbour on another
// GPU is a remote link whose ghost / EMF messages the pack kernels store straight into the owner's
// receive buffer over NVLink (include/athk.h, athk_remote_set_p2p).  All links between two GPUs are
// coalesced into one region per quantity, ordered canonically by (receiver's block, offset as the
// receiver sees it) so that sender and receiver agree on the layout without talking.
#include <algorithm>
#include <map>

#include "athena.hpp"

namespace {
std::uint64_t Morton(const LogicalLocation &l) {
  std::uint64_t k = 0;
  for (int bit = 0; bit < 21; ++bit) {
    k |= ((std::uint64_t)((l.lx1 >> bit) & 1)) << (3 * bit);
    k |= ((std::uint64_t)((l.lx2 >> bit) & 1)) << (3 * bit + 1);
    k |= ((std::uint64_t)((l.lx3 >> bit) & 1)) << (3 * bit + 2);
  }
  return k;
}
int Slot(int o1, int o2, int o3) { return (o3 + 1) * 9 + (o2 + 1) * 3 + (o1 + 1); }

struct Link {
  int lb, o1, o2, o3;      // local block (on its GPU) and neighbour offset
  int my_gid, nb_gid, peer;
  long long gcount, ecount;
  long long gs, gr, es, er;   // element offsets in this GPU's send / receive layouts
};
}  // namespace

void Mesh::SetupRemoteLinks() {
  const char *who = "Mesh::SetupRemoteLinks";
  std::map<std::uint64_t, int> gid_of;
  for (int n = 0; n < nbtotal; ++n) gid_of[Morton(loclist[n])] = n;
  std::vector<std::vector<Link>> links(ngpu);
  for (int r = 0; r < ngpu; ++r) {
    std::vector<int> nbr((std::size_t)nblist[r] * 27, -1);
    for (int lb = 0; lb < nblist[r]; ++lb) {
      const int n = nslist[r] + lb;
      const LogicalLocation &l = loclist[n];
      if (multilevel) {
        // refined mesh: the neighbours SearchAndSetNeighbors found (refinement.cpp); other levels are
        // served by the box lists and only marked in the table
        const MeshBlock *pmb = blocks_[n];
        for (const NeighborBlock &nb : pmb->neighbor) {
          int &entry = nbr[(std::size_t)lb * 27 + Slot(nb.ox1, nb.ox2, nb.ox3)];
          if (nb.level != pmb->loc.level) {
            entry = ATHK_NBR_REFINED;
          } else if (nb.rank == r) {
            entry = nb.lid;
          } else {
            entry = -(2 + (int)links[r].size());
            Link k = {lb, nb.ox1, nb.ox2, nb.ox3, n, nb.gid, nb.rank, 0, 0, 0, 0, 0, 0};
            k.gcount = athk_remote_ghost_count(ctx[r], nb.ox1, nb.ox2, nb.ox3);
            k.ecount = athk_remote_emf_count(ctx[r], nb.ox1, nb.ox2, nb.ox3);
            links[r].push_back(k);
          }
        }
        continue;
      }
      for (int o3 = -1; o3 <= 1; ++o3)
        for (int o2 = -1; o2 <= 1; ++o2)
          for (int o1 = -1; o1 <= 1; ++o1) {
            if (o1 == 0 && o2 == 0 && o3 == 0) continue;
            // no neighbour across a non-periodic mesh boundary (entry stays -1)
            const std::int64_t q1 = l.lx1 + o1, q2 = l.lx2 + o2, q3 = l.lx3 + o3;
            if ((q1 < 0 || q1 >= nrbx1) && mesh_bcs[0] != BoundaryFlag::periodic) continue;
            if ((q2 < 0 || q2 >= nrbx2) && mesh_bcs[2] != BoundaryFlag::periodic) continue;
            if ((q3 < 0 || q3 >= nrbx3) && mesh_bcs[4] != BoundaryFlag::periodic) continue;
            LogicalLocation q{(q1 + nrbx1) % nrbx1, (q2 + nrbx2) % nrbx2, (q3 + nrbx3) % nrbx3, l.level};
            const int g = gid_of[Morton(q)];
            int &entry = nbr[(std::size_t)lb * 27 + Slot(o1, o2, o3)];
            if (ranklist[g] == r) {
              entry = g - nslist[r];
            } else {
              entry = -(2 + (int)links[r].size());
              Link k = {lb, o1, o2, o3, n, g, ranklist[g], 0, 0, 0, 0, 0, 0};
              k.gcount = athk_remote_ghost_count(ctx[r], o1, o2, o3);
              k.ecount = athk_remote_emf_count(ctx[r], o1, o2, o3);
              links[r].push_back(k);
            }
          }
    }
    AthkCheck(athk_set_neighbors(ctx[r], nbr.data()), who);
  }
  if (ngpu == 1) return;

  // per GPU and peer: where the region for that peer starts in the send / receive layouts
  struct Seg { long long gs0, gr0, es0, er0; };
  std::vector<std::map<int, Seg>> seg(ngpu);
  std::vector<long long> g_r(ngpu, 0), e_r(ngpu, 0);
  std::vector<std::vector<int>> peers(ngpu);
  for (int r = 0; r < ngpu; ++r) {
    std::vector<Link> &L = links[r];
    for (const Link &k : L)
      if (std::find(peers[r].begin(), peers[r].end(), k.peer) == peers[r].end()) peers[r].push_back(k.peer);
    std::sort(peers[r].begin(), peers[r].end());
    long long gs = 0, gr = 0, es = 0, er = 0;
    for (int p : peers[r]) {
      std::vector<int> mine;
      for (int q = 0; q < (int)L.size(); ++q) if (L[q].peer == p) mine.push_back(q);
      // sender side: the receiver's block is the neighbour, the offset as it sees it is -o
      std::vector<int> so = mine, ro = mine;
      std::sort(so.begin(), so.end(), [&](int a, int b) {
        return std::make_pair(L[a].nb_gid, Slot(-L[a].o1, -L[a].o2, -L[a].o3)) <
               std::make_pair(L[b].nb_gid, Slot(-L[b].o1, -L[b].o2, -L[b].o3));
      });
      std::sort(ro.begin(), ro.end(), [&](int a, int b) {
        return std::make_pair(L[a].my_gid, Slot(L[a].o1, L[a].o2, L[a].o3)) <
               std::make_pair(L[b].my_gid, Slot(L[b].o1, L[b].o2, L[b].o3));
      });
      seg[r][p] = Seg{gs, gr, es, er};
      for (int q : so) { L[q].gs = gs; gs += L[q].gcount; L[q].es = es; es += L[q].ecount; }
      for (int q : ro) { L[q].gr = gr; gr += L[q].gcount; L[q].er = er; er += L[q].ecount; }
    }
    g_r[r] = gr;
    e_r[r] = er;
    const int n = (int)L.size();
    std::vector<int> lb(n), o1(n), o2(n), o3(n);
    std::vector<long long> vgs(n), vgr(n), ves(n), ver(n);
    for (int q = 0; q < n; ++q) {
      lb[q] = L[q].lb; o1[q] = L[q].o1; o2[q] = L[q].o2; o3[q] = L[q].o3;
      vgs[q] = L[q].gs; vgr[q] = L[q].gr; ves[q] = L[q].es; ver[q] = L[q].er;
    }
    AthkCheck(athk_remote_configure(ctx[r], n, lb.data(), o1.data(), o2.data(), o3.data(), vgs.data(), vgr.data(),
                                    ves.data(), ver.data()), who);
  }
  // receive buffers (two halves each) and flag words of every GPU; peer access both ways
  std::vector<double *> gbuf(ngpu, nullptr), ebuf(ngpu, nullptr);
  std::vector<unsigned long long *> fbuf(ngpu, nullptr);
  p2p_allocs_.assign((std::size_t)3 * ngpu, nullptr);
  const int dev0 = 0;
  (void)dev0;
  for (int r = 0; r < ngpu; ++r) {
    void *p = nullptr;
    AthkCheck(athk_p2p_alloc(ctx[r], 16 * std::max<long long>(g_r[r], 8), &p, nullptr), who);
    gbuf[r] = static_cast<double *>(p); p2p_allocs_[3 * r] = p;
    AthkCheck(athk_p2p_alloc(ctx[r], 16 * std::max<long long>(e_r[r], 8), &p, nullptr), who);
    ebuf[r] = static_cast<double *>(p); p2p_allocs_[3 * r + 1] = p;
    AthkCheck(athk_p2p_alloc(ctx[r], 16 * std::max<long long>((long long)peers[r].size(), 8), &p, nullptr), who);
    fbuf[r] = static_cast<unsigned long long *>(p); p2p_allocs_[3 * r + 2] = p;
  }
  for (int r = 0; r < ngpu; ++r) {
    if (multilevel) {   // boxes between levels read any GPU's blocks directly
      for (int p = 0; p < ngpu; ++p) if (p != r) AthkCheck(athk_p2p_enable_peer(ctx[r], gpu_device0_ + p), who);
    } else {
      for (int p : peers[r]) AthkCheck(athk_p2p_enable_peer(ctx[r], gpu_device0_ + p), who);
    }
  }
  for (int r = 0; r < ngpu; ++r) {
    const std::vector<Link> &L = links[r];
    const int n = (int)L.size();
    if (n == 0) continue;
    std::vector<double *> gdst((std::size_t)2 * n), edst((std::size_t)2 * n);
    for (int q = 0; q < n; ++q) {
      const int p = L[q].peer;
      const long long goff = seg[p][r].gr0 + (L[q].gs - seg[r][p].gs0);
      const long long eoff = seg[p][r].er0 + (L[q].es - seg[r][p].es0);
      for (int par = 0; par < 2; ++par) {
        gdst[(std::size_t)par * n + q] = gbuf[p] + par * g_r[p] + goff;
        edst[(std::size_t)par * n + q] = ebuf[p] + par * e_r[p] + eoff;
      }
    }
    std::vector<unsigned long long *> pflags(peers[r].size());
    for (std::size_t k = 0; k < peers[r].size(); ++k) {
      const int p = peers[r][k];
      const std::size_t me = std::find(peers[p].begin(), peers[p].end(), r) - peers[p].begin();
      pflags[k] = fbuf[p] + 2 * me;
    }
    AthkCheck(athk_remote_set_p2p(ctx[r], (int)peers[r].size(), gdst.data(), edst.data(), gbuf[r], g_r[r], ebuf[r],
                                  e_r[r], pflags.data(), fbuf[r]), who);
  }
}
