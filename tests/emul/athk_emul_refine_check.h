// TEST INFRASTRUCTURE ONLY (host harness build, -DATHK_EMULATE; never part of libathk.so).
// Boxes of one refinement list run concurrently on the device, so two boxes of a list may only overlap
// where they write identical values.  With ATHK_REFINE_ORDER_CHECK set, athk_refine_run executes the
// list box by box in forward and in reverse order from the same state and fails when the results differ.
#ifndef ATHK_EMUL_REFINE_CHECK_H_
#define ATHK_EMUL_REFINE_CHECK_H_
#include <cstdlib>
#include <cstring>
#include <vector>

static int athk_emul_refine_order_check(athk_ctx *c, int phase, int reg, BoxOpPoint f) {
  static const bool on = getenv("ATHK_REFINE_ORDER_CHECK") != nullptr;
  if (!on || c->rf.nops[phase] < 2) return 0;
  const Grid &g = c->g;
  const CoarseGrid &cg = c->rf.cg;
  const bool mhd = c->cfg.mhd != 0;
  struct Arr { Real *p; size_t n; };
  std::vector<Arr> arrs = {{c->u[reg], (size_t)g.nb * 5 * g.sn}, {c->w, (size_t)g.nb * 5 * g.sn},
                           {c->rf.cu, (size_t)g.nb * 5 * cg.csn}, {c->rf.cw, (size_t)g.nb * 5 * cg.csn}};
  for (int d = 0; d < 3; ++d) if (c->flux[d]) arrs.push_back({c->flux[d], (size_t)g.nb * 5 * g.sn});
  if (mhd) {
    arrs.push_back({c->b[reg], (size_t)g.nb * 3 * g.sn});
    arrs.push_back({c->bcc, (size_t)g.nb * 3 * g.sn});
    arrs.push_back({c->e, (size_t)g.nb * 3 * g.sn});
    arrs.push_back({c->rf.cb, (size_t)g.nb * 3 * cg.cfn});
    arrs.push_back({c->rf.cbcc, (size_t)g.nb * 3 * cg.csn});
  }
  std::vector<std::vector<Real>> before, fwd;
  for (const Arr &a : arrs) before.emplace_back(a.p, a.p + a.n);
  auto run = [&](bool reverse) {
    const int n = c->rf.nops[phase];
    for (int q = 0; q < n; ++q) {
      f.ops = c->rf.ops[phase] + (reverse ? n - 1 - q : q);
      launch_links(c->stream, 1, c->rf.maxvol[phase], f);
    }
  };
  run(false);
  for (const Arr &a : arrs) fwd.emplace_back(a.p, a.p + a.n);
  for (size_t q = 0; q < arrs.size(); ++q) memcpy(arrs[q].p, before[q].data(), arrs[q].n * sizeof(Real));
  run(true);
  int rc = 0;
  for (size_t q = 0; q < arrs.size() && !rc; ++q)
    if (memcmp(arrs[q].p, fwd[q].data(), arrs[q].n * sizeof(Real)) != 0)
      rc = fail(ATHK_ERR_STATE, "refine_run: the boxes of list %d disagree about an element of array %d (order-dependent "
                                "result; the device would race)", phase, (int)q);
  for (size_t q = 0; q < arrs.size(); ++q) memcpy(arrs[q].p, before[q].data(), arrs[q].n * sizeof(Real));
  return rc;
}
#endif  // ATHK_EMUL_REFINE_CHECK_H_
