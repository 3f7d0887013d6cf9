// athk_api.cu -- context, memory and the C ABI (include/athk.h) of libathk.so.
// Product build: nvcc -gencode arch=compute_100a,code=sm_100a --fmad=false (see athk_physics.h).
#include "athk_ctx.h"
#include "athk_sweep_launch_impl.h"

ATHK_INSTANTIATE_RS(athk::RS_HLLC, false)
ATHK_INSTANTIATE_RS(athk::RS_HLLD, true)
ATHK_EXTERN_RS(athk::RS_ROE, false)
ATHK_EXTERN_RS(athk::RS_ROE, true)
ATHK_EXTERN_RS(athk::RS_HLLE, false)
ATHK_EXTERN_RS(athk::RS_HLLE, true)
ATHK_EXTERN_RS(athk::RS_LLF, false)
ATHK_EXTERN_RS(athk::RS_LLF, true)
ATHK_EXTERN_RS(athk::RS_HLLE_ISO, false)   // athk_rs_iso.cu
ATHK_EXTERN_RS(athk::RS_HLLD_ISO, true)
ATHK_EXTERN_RS(athk::RS_SR, false)         // athk_rs_sr.cu
#ifndef ATHK_EMULATE
ATHK_EXTERN_RS_FAST(athk::RS_HLLC, false)   // athk_fast.cu (--fmad=true)
ATHK_EXTERN_RS_FAST(athk::RS_HLLD, true)
#endif

static thread_local std::string g_last_error;

// peer-to-peer transport hooks (defined at the end of this file)
int athk_p2p_before_ghost_fill(athk_ctx *c);
int athk_p2p_before_emf_apply(athk_ctx *c);
int athk_p2p_after_pack(athk_ctx *c, int which);
static int ghost_pack_early(athk_ctx *c, int reg);
static int ghost_pack_finish(athk_ctx *c, int reg);

#undef fail
int athk_fail(int code, const char *fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}
#define fail athk_fail


static size_t cc_elems(const athk_ctx *c, int nvar) { return (size_t)c->g.nb * nvar * c->g.sn; }

template <class T>
static int dev_alloc(T **p, size_t n, cudaStream_t st) {
  *p = nullptr;
  if (n == 0) return 0;
  CU(cudaMalloc((void **)p, n * sizeof(T)));
  CU(cudaMemsetAsync(*p, 0, n * sizeof(T), st));
  return 0;
}

extern "C" const char *athk_last_error(void) { return g_last_error.c_str(); }

extern "C" int athk_create(const athk_config *cfg, athk_ctx **out) {
  if (!cfg || !out) return fail(ATHK_ERR_ARG, "athk_create: null argument");
  *out = nullptr;
  if (cfg->nx1 < 4 || cfg->nx2 < 4 || cfg->nx3 < 4)
    return fail(ATHK_ERR_ARG, "athk_create: only 3-D MeshBlocks with nx>=4 are supported "
                              "(got %d x %d x %d)", cfg->nx1, cfg->nx2, cfg->nx3);
  if (cfg->nghost < 2 || cfg->nghost > 4)
    return fail(ATHK_ERR_ARG, "athk_create: nghost must be 2..4 (got %d)", cfg->nghost);
  if (cfg->nblocks < 1) return fail(ATHK_ERR_ARG, "athk_create: nblocks must be >= 1");
  if (cfg->nscalars < 0 || cfg->nscalars > 8)
    return fail(ATHK_ERR_ARG, "athk_create: nscalars must be 0..8");
  // configure.py:376-389: hllc is hydro-only, hlld MHD-only; roe / hlle / llf exist for both
  if (cfg->rsolver < ATHK_RS_HLLC || cfg->rsolver > ATHK_RS_LLF)
    return fail(ATHK_ERR_ARG, "athk_create: unknown rsolver %d", cfg->rsolver);
  if (cfg->mhd && cfg->rsolver == ATHK_RS_HLLC)
    return fail(ATHK_ERR_ARG, "athk_create: rsolver HLLC is hydro-only (use HLLD, Roe, HLLE or LLF with mhd=1)");
  if (!cfg->mhd && cfg->rsolver == ATHK_RS_HLLD)
    return fail(ATHK_ERR_ARG, "athk_create: rsolver HLLD needs mhd=1 (use HLLC, Roe, HLLE or LLF)");
  if (cfg->eos != ATHK_EOS_ISOTHERMAL && !(cfg->gamma > 1.0)) return fail(ATHK_ERR_ARG, "athk_create: gamma must be > 1");
  if (cfg->arith != ATHK_ARITH_STRICT && cfg->arith != ATHK_ARITH_FAST)
    return fail(ATHK_ERR_ARG, "athk_create: arith must be ATHK_ARITH_STRICT (0) or ATHK_ARITH_FAST (1)");
  if (cfg->eos != ATHK_EOS_ADIABATIC && cfg->eos != ATHK_EOS_ISOTHERMAL && cfg->eos != ATHK_EOS_ADIABATIC_SR)
    return fail(ATHK_ERR_ARG, "athk_create: unknown eos %d", cfg->eos);
  if (cfg->eos == ATHK_EOS_ADIABATIC_SR) {   // configure.py -s
    if (cfg->mhd) return fail(ATHK_ERR_ARG, "athk_create: special relativity is implemented for hydro only (mhd = 0)");
    if (cfg->rsolver != ATHK_RS_HLLE && cfg->rsolver != ATHK_RS_HLLC && cfg->rsolver != ATHK_RS_LLF)
      return fail(ATHK_ERR_ARG, "athk_create: special relativity is implemented with rsolver HLLC, HLLE, LLF "
                                "(hllc_rel.cpp, hlle_rel.cpp, llf_rel.cpp)");
    if (cfg->arith == ATHK_ARITH_FAST) return fail(ATHK_ERR_ARG, "athk_create: no fast-arithmetic set for special relativity");
  }
  if (cfg->eos == ATHK_EOS_ISOTHERMAL) {
    if (!(cfg->iso_sound_speed > 0.0)) return fail(ATHK_ERR_ARG, "athk_create: isothermal EOS needs iso_sound_speed > 0");
    // hlle.cpp / hlld_iso.cpp have their own kernels; LLF, HLLE (MHD) and Roe (hydro) branch on the EOS
    // inside the adiabatic kernels' flux functions (NON_BAROTROPIC_EOS == 0 branches of llf.cpp, llf_mhd.cpp,
    // hlle_mhd.cpp, roe.cpp, roe_mhd.cpp)
    const bool ok = cfg->mhd ? (cfg->rsolver == ATHK_RS_HLLD || cfg->rsolver == ATHK_RS_HLLE || cfg->rsolver == ATHK_RS_LLF ||
                                cfg->rsolver == ATHK_RS_ROE)
                             : (cfg->rsolver == ATHK_RS_HLLE || cfg->rsolver == ATHK_RS_LLF || cfg->rsolver == ATHK_RS_ROE);
    if (!ok)
      return fail(ATHK_ERR_ARG, "athk_create: the isothermal EOS is implemented with rsolver HLLE, LLF, Roe (hydro) and "
                                "HLLD, HLLE, LLF, Roe (mhd)");
    if (cfg->arith == ATHK_ARITH_FAST) return fail(ATHK_ERR_ARG, "athk_create: no fast-arithmetic set for the isothermal EOS");
  }
  if (cfg->arith == ATHK_ARITH_FAST && cfg->rsolver != ATHK_RS_HLLC && cfg->rsolver != ATHK_RS_HLLD)
    return fail(ATHK_ERR_ARG, "athk_create: the fast-arithmetic kernel set exists for HLLC / HLLD only");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1)
    return fail(ATHK_ERR_NODEVICE, "athk_create: no CUDA device visible; libathk has no CPU path");
  if (cfg->device < 0 || cfg->device >= ndev)
    return fail(ATHK_ERR_ARG, "athk_create: device %d out of range (%d visible)", cfg->device, ndev);
  CU(cudaSetDevice(cfg->device));

  athk_ctx *c = new athk_ctx();
  c->cfg = *cfg;
  Grid &g = c->g;
  g.nx1 = cfg->nx1; g.nx2 = cfg->nx2; g.nx3 = cfg->nx3; g.ng = cfg->nghost;
  g.nc1 = g.nx1 + 2 * g.ng; g.nc2 = g.nx2 + 2 * g.ng; g.nc3 = g.nx3 + 2 * g.ng;
  g.is = g.ng; g.ie = g.is + g.nx1 - 1;
  g.js = g.ng; g.je = g.js + g.nx2 - 1;
  g.ks = g.ng; g.ke = g.ks + g.nx3 - 1;
  g.P1 = ((g.nc1 + 1 + 3) / 4) * 4;
  g.P2 = g.nc2 + 1;
  g.P3 = g.nc3 + 1;
  g.nb = cfg->nblocks;
  g.sj = g.P1;
  g.sk = (idx_t)g.P1 * g.P2;
  g.sn = g.sk * g.P3;
  const Real fmin_default = sqrt(1024 * (Real)FLT_MIN);   // eos.hpp:161, adiabatic_mhd.cpp:29-31
  c->eos = make_eos(cfg->gamma, cfg->dfloor > 0 ? cfg->dfloor : fmin_default,
                    cfg->pfloor > 0 ? cfg->pfloor : fmin_default,
                    cfg->sfloor > 0 ? cfg->sfloor : fmin_default);
  c->eos.iso = (cfg->eos == ATHK_EOS_ISOTHERMAL) ? 1 : 0;
  c->eos.iso_cs = cfg->iso_sound_speed;
  c->eos.sr = (cfg->eos == ATHK_EOS_ADIABATIC_SR) ? 1 : 0;
  c->eos.gamma_max = cfg->gamma_max > 0.0 ? cfg->gamma_max : 1000.0;
  c->eos.sr_solver = cfg->rsolver;
  c->sf.nx1 = g.nx1; c->sf.nx2 = g.nx2; c->sf.nx3 = g.nx3;
  c->launches = 0;
  c->dt_requested = c->dt_valid = false;
  {   // planes per thread of the k-marching stencil kernels (development knobs: ATHK_KT_CORNER, ATHK_KT_CT)
    const char *e1 = getenv("ATHK_KT_CORNER"), *e2 = getenv("ATHK_KT_CT");
    c->kt_corner = e1 ? atoi(e1) : 13;
    c->kt_ct = e2 ? atoi(e2) : 13;
  }
  c->prof_on = false;
  c->coords_set = c->nbr_set = false;
  c->nremote = 0;
  memset(&c->rl, 0, sizeof(c->rl));
  c->rl_int = nullptr; c->rl_off = nullptr;
  c->max_ghost_count = c->max_emf_count = 0;
  CU(cudaStreamCreate(&c->stream));
  cudaStream_t st = c->stream;
  const int ns = cfg->nscalars;
  int rc = 0;
  for (int r = 0; r < 3 && !rc; ++r) rc = dev_alloc(&c->u[r], r < 2 ? cc_elems(c, 5) : 0, st);
  if (!rc) rc = dev_alloc(&c->w, cc_elems(c, 5), st);
  c->w_alt = nullptr;
  if (!rc && c->eos.sr) rc = dev_alloc(&c->w_alt, cc_elems(c, 5), st);
  for (int d = 0; d < 3 && !rc; ++d) rc = dev_alloc(&c->flux[d], cc_elems(c, 5), st);
  const size_t mh = cfg->mhd ? 1 : 0;
  if (!rc) rc = dev_alloc(&c->bcc, mh * cc_elems(c, 3), st);
  for (int r = 0; r < 3 && !rc; ++r) rc = dev_alloc(&c->b[r], (r < 2 ? mh : 0) * cc_elems(c, 3), st);
  if (!rc) rc = dev_alloc(&c->ef, mh * cc_elems(c, 6), st);
  if (!rc) rc = dev_alloc(&c->wct, mh * cc_elems(c, 3), st);
  if (!rc) rc = dev_alloc(&c->e, mh * cc_elems(c, 3), st);
  if (!rc) rc = dev_alloc(&c->emfbuf, mh * (size_t)g.nb * c->sf.per_block(), st);
  for (int r = 0; r < 3 && !rc; ++r) rc = dev_alloc(&c->s[r], (r < 2 && ns) ? cc_elems(c, ns) : 0, st);
  if (!rc) rc = dev_alloc(&c->r, ns ? cc_elems(c, ns) : 0, st);
  for (int d = 0; d < 3 && !rc; ++d) rc = dev_alloc(&c->sflux[d], ns ? cc_elems(c, ns) : 0, st);
  int PG = g.P1 > g.P2 ? g.P1 : g.P2;
  if (g.P3 > PG) PG = g.P3;
  c->geo.PG = PG;
  if (!rc) rc = dev_alloc(&c->geo_store, (size_t)9 * g.nb * PG, st);
  if (!rc) rc = dev_alloc(&c->nbr, (size_t)g.nb * 27, st);
  if (!rc) rc = dev_alloc(&c->dt_blocks, (size_t)g.nb, st);
  if (rc) {
    athk_destroy(c);
    return rc;
  }
  for (int d = 0; d < 3; ++d) {
    c->geo.dxf[d] = c->geo_store + (size_t)(0 + d) * g.nb * PG;
    c->geo.cp[d] = c->geo_store + (size_t)(3 + d) * g.nb * PG;
    c->geo.cm[d] = c->geo_store + (size_t)(6 + d) * g.nb * PG;
  }
  {   // slab geometry table for the remote-link kernels
    SlabTable *ht = new SlabTable();
    for (int o3 = -1; o3 <= 1; ++o3)
      for (int o2 = -1; o2 <= 1; ++o2)
        for (int o1 = -1; o1 <= 1; ++o1) {
          const int os = nbr_slot(o1, o2, o3);
          for (int q = 0; q < 4; ++q) ht->slab[os][q] = make_slab(g, q - 1, o1, o2, o3);
          for (int sct = 0; sct < 6; ++sct)
            ht->sec[os][sct] = ghost_section_off(g, cfg->mhd != 0, cfg->nscalars, o1, o2, o3, sct);
        }
    c->slab_tab = nullptr;
    if (cudaMalloc((void **)&c->slab_tab, sizeof(SlabTable)) == cudaSuccess)
      cudaMemcpy(c->slab_tab, ht, sizeof(SlabTable), cudaMemcpyHostToDevice);
    delete ht;
    c->rl.tab = c->slab_tab;
  }
  c->h_geo.assign((size_t)9 * g.nb * PG, 0.0);
  c->h_nbr.assign((size_t)g.nb * 27, -1);
  c->h_bcflags.assign((size_t)g.nb * 6, 0);
  c->bcflags = nullptr;
  c->bc_dirty = false;
  c->stage = nullptr;
  c->stage_count = 0;
  for (int f = 0; f < 6; ++f) c->bc_any[f] = 0;
  CU(cudaStreamSynchronize(st));
  *out = c;
  return ATHK_OK;
}

extern "C" void athk_destroy(athk_ctx *c) {
  if (!c) return;
  cudaSetDevice(c->cfg.device);
  for (int r = 0; r < 3; ++r) { cudaFree(c->u[r]); cudaFree(c->b[r]); cudaFree(c->s[r]); }
  for (int d = 0; d < 3; ++d) { cudaFree(c->flux[d]); cudaFree(c->sflux[d]); }
  cudaFree(c->w); cudaFree(c->w_alt); cudaFree(c->bcc); cudaFree(c->ef); cudaFree(c->wct); cudaFree(c->e);
  cudaFree(c->bcflags);
  cudaFree(c->d_floors);
  cudaFree(c->stage);
  cudaFree(c->r); cudaFree(c->emfbuf); cudaFree(c->geo_store); cudaFree(c->nbr);
  cudaFree(c->rf.cu); cudaFree(c->rf.cw); cudaFree(c->rf.xv); cudaFree(c->rf.cxv);
  cudaFree(c->rf.cb); cudaFree(c->rf.cbcc); cudaFree(c->rf.cs); cudaFree(c->rf.cr);
  for (int p = 0; p < ATHK_REFINE_NPHASES; ++p) cudaFree(c->rf.ops[p]);
  cudaFree(c->dt_blocks);
  cudaFree(c->rl_int); cudaFree(c->rl_off); cudaFree(c->slab_tab);
  cudaFree(c->p2p.gdst); cudaFree(c->p2p.edst); cudaFree(c->p2p.peer_flags); cudaFree(c->p2p.timeouts);
#ifndef ATHK_EMULATE
  if (c->p2p.side) { cudaStreamDestroy(c->p2p.side); cudaEventDestroy(c->p2p.ev_ready); cudaEventDestroy(c->p2p.ev_packed); }
#endif
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

extern "C" int athk_sync(athk_ctx *c) {
  if (!c) return fail(ATHK_ERR_ARG, "null context");
  cudaSetDevice(c->cfg.device);
  CU(cudaStreamSynchronize(c->stream));
  return ATHK_OK;
}

extern "C" long long athk_launch_count(const athk_ctx *c) { return c ? c->launches : 0; }

extern "C" int athk_profile(athk_ctx *c, int enable, char *report, int report_len) {
  if (!c) return fail(ATHK_ERR_ARG, "null context");
  cudaSetDevice(c->cfg.device);
#ifndef ATHK_EMULATE
  if (report && report_len > 0) {
    CU(cudaStreamSynchronize(c->stream));
    std::vector<std::string> names;
    std::vector<double> ms;
    std::vector<int> cnt;
    for (size_t q = 0; q < c->prof.size(); ++q) {
      float t = 0.f;
      if (cudaEventElapsedTime(&t, c->prof[q].a, c->prof[q].b) != cudaSuccess) t = 0.f;
      size_t k = 0;
      for (; k < names.size(); ++k) if (names[k] == c->prof[q].name) break;
      if (k == names.size()) { names.push_back(c->prof[q].name); ms.push_back(0.0); cnt.push_back(0); }
      ms[k] += t;
      cnt[k]++;
      cudaEventDestroy(c->prof[q].a);
      cudaEventDestroy(c->prof[q].b);
    }
    c->prof.clear();
    std::string out = "{";
    char line[256];
    for (size_t k = 0; k < names.size(); ++k) {
      snprintf(line, sizeof(line), "%s\"%s\": {\"ms\": %.6f, \"launches\": %d}", k ? ", " : "",
               names[k].c_str(), ms[k], cnt[k]);
      out += line;
    }
    out += "}";
    snprintf(report, report_len, "%s", out.c_str());
  }
#else
  if (report && report_len > 0) snprintf(report, report_len, "{}");
#endif
  c->prof_on = enable != 0;
  return ATHK_OK;
}
extern "C" void *athk_stream(athk_ctx *c) { return c ? (void *)c->stream : nullptr; }

//----------------------------------------------------------------------------------------
extern "C" int athk_set_block_coords(athk_ctx *c, int lb, const double *x1f, const double *x1v,
                                     const double *dx1f, const double *x2f, const double *x2v,
                                     const double *dx2f, const double *x3f, const double *x3v,
                                     const double *dx3f) {
  if (!c) return fail(ATHK_ERR_ARG, "null context");
  cudaSetDevice(c->cfg.device);
  if (lb < 0 || lb >= c->g.nb) return fail(ATHK_ERR_ARG, "set_block_coords: block %d out of range", lb);
  const double *xf[3] = {x1f, x2f, x3f}, *xv[3] = {x1v, x2v, x3v}, *dxf[3] = {dx1f, dx2f, dx3f};
  const int nc[3] = {c->g.nc1, c->g.nc2, c->g.nc3};
  const int PG = c->geo.PG;
  const size_t per = (size_t)c->g.nb * PG;
  for (int d = 0; d < 3; ++d) {
    if (!xf[d] || !xv[d] || !dxf[d]) return fail(ATHK_ERR_ARG, "set_block_coords: null array");
    Real *hd = &c->h_geo[(0 + d) * per + (size_t)lb * PG];
    Real *hp = &c->h_geo[(3 + d) * per + (size_t)lb * PG];
    Real *hm = &c->h_geo[(6 + d) * per + (size_t)lb * PG];
    for (int n = 0; n < nc[d]; ++n) {
      hd[n] = dxf[d][n];
      // plm.cpp:152-153 -- the geometric factors are ~0.5 but not exactly; keep the same ops
      hp[n] = (xf[d][n + 1] - xv[d][n]) / dxf[d][n];
      hm[n] = (xv[d][n] - xf[d][n]) / dxf[d][n];
    }
  }
  for (int a = 0; a < 9; ++a)
    CU(cudaMemcpyAsync(c->geo_store + a * per + (size_t)lb * PG, &c->h_geo[a * per + (size_t)lb * PG],
                       PG * sizeof(Real), cudaMemcpyHostToDevice, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  c->coords_set = true;
  return ATHK_OK;
}

extern "C" int athk_set_neighbors(athk_ctx *c, const int *nbr) {
  if (!c || !nbr) return fail(ATHK_ERR_ARG, "set_neighbors: null argument");
  cudaSetDevice(c->cfg.device);
  const int n = c->g.nb * 27;
  for (int q = 0; q < n; ++q) {
    if (nbr[q] >= c->g.nb) return fail(ATHK_ERR_ARG, "set_neighbors: entry %d = %d out of range", q, nbr[q]);
    c->h_nbr[q] = nbr[q];
  }
  CU(cudaMemcpyAsync(c->nbr, c->h_nbr.data(), n * sizeof(int), cudaMemcpyHostToDevice, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  c->nbr_set = true;
  return ATHK_OK;
}

//----------------------------------------------------------------------------------------
// array table: device base, component offset, nvar, extra extent per direction
struct ArrDesc {
  Real *base;
  int ncomp_total;   // components in the device array per block
  int comp0;         // first component
  int nvar;          // components exposed
  int a1, a2, a3;    // host dims = (nc3+a3, nc2+a2, nc1+a1)
};

static bool describe(athk_ctx *c, int array, ArrDesc *d) {
  const int ns = c->cfg.nscalars;
  switch (array) {
    // (isothermal EOS: the host's arrays have NHYDRO = 4 components; the device's fifth stays zero)
    case ATHK_U: case ATHK_U1: case ATHK_U2:
      *d = {c->u[array - ATHK_U], 5, 0, c->eos.iso ? 4 : 5, 0, 0, 0}; break;
    case ATHK_W: *d = {c->w, 5, 0, c->eos.iso ? 4 : 5, 0, 0, 0}; break;
    case ATHK_W1: if (!c->w_alt) return false; *d = {c->w_alt, 5, 0, 5, 0, 0, 0}; break;
    case ATHK_BCC: *d = {c->bcc, 3, 0, 3, 0, 0, 0}; break;
    case ATHK_B_X1F: case ATHK_B_X2F: case ATHK_B_X3F: {
      const int q = array - ATHK_B_X1F;
      *d = {c->b[0], 3, q, 1, q == 0, q == 1, q == 2}; break; }
    case ATHK_B1_X1F: case ATHK_B1_X2F: case ATHK_B1_X3F: {
      const int q = array - ATHK_B1_X1F;
      *d = {c->b[1], 3, q, 1, q == 0, q == 1, q == 2}; break; }
    case ATHK_B2_X1F: case ATHK_B2_X2F: case ATHK_B2_X3F: {
      const int q = array - ATHK_B2_X1F;
      *d = {c->b[2], 3, q, 1, q == 0, q == 1, q == 2}; break; }
    case ATHK_FLUX_X1: case ATHK_FLUX_X2: case ATHK_FLUX_X3: {
      const int q = array - ATHK_FLUX_X1;
      *d = {c->flux[q], 5, 0, 5, q == 0, q == 1, q == 2}; break; }
    case ATHK_E_X1E: *d = {c->e, 3, 0, 1, 0, 1, 1}; break;
    case ATHK_E_X2E: *d = {c->e, 3, 1, 1, 1, 0, 1}; break;
    case ATHK_E_X3E: *d = {c->e, 3, 2, 1, 1, 1, 0}; break;
    case ATHK_E3_X1F: *d = {c->ef, 6, 0, 1, 1, 0, 0}; break;
    case ATHK_E2_X1F: *d = {c->ef, 6, 1, 1, 1, 0, 0}; break;
    case ATHK_E1_X2F: *d = {c->ef, 6, 2, 1, 0, 1, 0}; break;
    case ATHK_E3_X2F: *d = {c->ef, 6, 3, 1, 0, 1, 0}; break;
    case ATHK_E2_X3F: *d = {c->ef, 6, 4, 1, 0, 0, 1}; break;
    case ATHK_E1_X3F: *d = {c->ef, 6, 5, 1, 0, 0, 1}; break;
    case ATHK_WGHT_X1F: *d = {c->wct, 3, 0, 1, 1, 0, 0}; break;
    case ATHK_WGHT_X2F: *d = {c->wct, 3, 1, 1, 0, 1, 0}; break;
    case ATHK_WGHT_X3F: *d = {c->wct, 3, 2, 1, 0, 0, 1}; break;
    case ATHK_S: case ATHK_S1: case ATHK_S2:
      *d = {c->s[array - ATHK_S], ns, 0, ns, 0, 0, 0}; break;
    case ATHK_R: *d = {c->r, ns, 0, ns, 0, 0, 0}; break;
    case ATHK_SFLUX_X1: case ATHK_SFLUX_X2: case ATHK_SFLUX_X3: {
      const int q = array - ATHK_SFLUX_X1;
      *d = {c->sflux[q], ns, 0, ns, q == 0, q == 1, q == 2}; break; }
    default: return false;
  }
  return d->base != nullptr && d->nvar > 0;
}

// Host <-> device copies of one array of one MeshBlock.  The host side is AthenaArray's dense
// layout, the device side the padded box: a strided cudaMemcpy3D moves 544-byte rows and reaches
// only ~10 GB/s over PCIe, so the copy goes through a dense device staging buffer instead -- one
// contiguous cudaMemcpyAsync at full PCIe rate plus a (microseconds) scatter / gather kernel.
// Everything is ordered on the context's stream, which also serialises the reuse of the buffer.
#ifndef ATHK_EMULATE
__global__ void __launch_bounds__(256) k_stage_copy(Real *dev, Real *dense, int nvar, int n1, int n2, int n3,
                                                    idx_t sn, idx_t sk, idx_t sj, int to_device) {
  const long long total = (long long)nvar * n3 * n2 * n1;
  for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < total;
       q += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(q % n1);
    long long t = q / n1;
    const int j = (int)(t % n2);
    t /= n2;
    const int k = (int)(t % n3);
    const int n = (int)(t / n3);
    Real *d = dev + (idx_t)n * sn + (idx_t)k * sk + (idx_t)j * sj + i;
    if (to_device) *d = dense[q];
    else dense[q] = *d;
  }
}
#endif

static int copy_array(athk_ctx *c, int array, int lb, double *host, bool to_device,
                      bool sync_after = true) {
  if (!c || !host) return fail(ATHK_ERR_ARG, "upload/download: null argument");
  cudaSetDevice(c->cfg.device);
  if (lb < 0 || lb >= c->g.nb) return fail(ATHK_ERR_ARG, "upload/download: block %d out of range", lb);
  ArrDesc d;
  if (!describe(c, array, &d))
    return fail(ATHK_ERR_ARG, "upload/download: array %d not allocated in this configuration", array);
  const Grid &g = c->g;
  const int n1 = g.nc1 + d.a1, n2 = g.nc2 + d.a2, n3 = g.nc3 + d.a3;
  Real *dev0 = d.base + ((idx_t)lb * d.ncomp_total + d.comp0) * g.sn;
#ifndef ATHK_EMULATE
  const size_t count = (size_t)d.nvar * n3 * n2 * n1;
  if (count > c->stage_count) {
    CU(cudaStreamSynchronize(c->stream));
    cudaFree(c->stage);
    c->stage = nullptr;
    c->stage_count = 0;
    CU(cudaMalloc((void **)&c->stage, count * sizeof(Real)));
    c->stage_count = count;
  }
  const int blocks = (int)((count + 255) / 256 < 148 * 16 ? (count + 255) / 256 : 148 * 16);
  if (to_device) {
    CU(cudaMemcpyAsync(c->stage, host, count * sizeof(Real), cudaMemcpyHostToDevice, c->stream));
    k_stage_copy<<<blocks, 256, 0, c->stream>>>(dev0, c->stage, d.nvar, n1, n2, n3, g.sn, g.sk, g.sj, 1);
    CU(cudaGetLastError());
  } else {
    k_stage_copy<<<blocks, 256, 0, c->stream>>>(dev0, c->stage, d.nvar, n1, n2, n3, g.sn, g.sk, g.sj, 0);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(host, c->stage, count * sizeof(Real), cudaMemcpyDeviceToHost, c->stream));
  }
#else
  for (int n = 0; n < d.nvar; ++n) {
    Real *dev = dev0 + (idx_t)n * g.sn;
    double *h = host + (size_t)n * n3 * n2 * n1;
    for (int k = 0; k < n3; ++k)
      for (int j = 0; j < n2; ++j) {
        Real *dr = dev + gidx(g, k, j, 0);
        double *hr = h + ((size_t)k * n2 + j) * n1;
        if (to_device) memcpy(dr, hr, n1 * sizeof(Real));
        else memcpy(hr, dr, n1 * sizeof(Real));
      }
  }
#endif
  if (!to_device && sync_after) CU(cudaStreamSynchronize(c->stream));
  return ATHK_OK;
}

// all MeshBlocks of an array through ONE contiguous PCIe copy and one scatter / gather kernel
#ifndef ATHK_EMULATE
__global__ void __launch_bounds__(256) k_stage_copy_all(Real *dev, Real *dense, int nb, int nvar, int n1, int n2, int n3,
                                                        int o1, int o2, int o3, idx_t sblk, idx_t sn, idx_t sk, idx_t sj,
                                                        int to_device) {
  const long long per = (long long)nvar * n3 * n2 * n1;
  const long long total = per * nb;
  for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < total;
       q += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(q % n1);
    long long t = q / n1;
    const int j = (int)(t % n2);
    t /= n2;
    const int k = (int)(t % n3);
    t /= n3;
    const int n = (int)(t % nvar);
    const int b = (int)(t / nvar);
    Real *d = dev + (idx_t)b * sblk + (idx_t)n * sn + (idx_t)(k + o3) * sk + (idx_t)(j + o2) * sj + (i + o1);
    if (to_device) *d = dense[q];
    else dense[q] = *d;
  }
}
#endif

static int copy_all(athk_ctx *c, int array, double *host, int active_only, bool to_device) {
  if (!c || !host) return fail(ATHK_ERR_ARG, "upload_all/download_all: null argument");
  cudaSetDevice(c->cfg.device);
  ArrDesc d;
  if (!describe(c, array, &d))
    return fail(ATHK_ERR_ARG, "upload_all/download_all: array %d not allocated in this configuration", array);
  const Grid &g = c->g;
  const int n1 = (active_only ? g.nx1 : g.nc1) + d.a1, n2 = (active_only ? g.nx2 : g.nc2) + d.a2,
            n3 = (active_only ? g.nx3 : g.nc3) + d.a3;
  const int o = active_only ? g.ng : 0;
  Real *dev0 = d.base + (idx_t)d.comp0 * g.sn;
  const idx_t sblk = (idx_t)d.ncomp_total * g.sn;
#ifndef ATHK_EMULATE
  const size_t count = (size_t)g.nb * d.nvar * n3 * n2 * n1;
  if (count > c->stage_count) {
    CU(cudaStreamSynchronize(c->stream));
    cudaFree(c->stage);
    c->stage = nullptr;
    c->stage_count = 0;
    CU(cudaMalloc((void **)&c->stage, count * sizeof(Real)));
    c->stage_count = count;
  }
  const int blocks = 148 * 16;
  if (to_device) {
    CU(cudaMemcpyAsync(c->stage, host, count * sizeof(Real), cudaMemcpyHostToDevice, c->stream));
    k_stage_copy_all<<<blocks, 256, 0, c->stream>>>(dev0, c->stage, g.nb, d.nvar, n1, n2, n3, o, o, o, sblk, g.sn, g.sk, g.sj, 1);
    CU(cudaGetLastError());
  } else {
    k_stage_copy_all<<<blocks, 256, 0, c->stream>>>(dev0, c->stage, g.nb, d.nvar, n1, n2, n3, o, o, o, sblk, g.sn, g.sk, g.sj, 0);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(host, c->stage, count * sizeof(Real), cudaMemcpyDeviceToHost, c->stream));
  }
  c->launches++;
#else
  double *h = host;
  for (int b = 0; b < g.nb; ++b)
    for (int n = 0; n < d.nvar; ++n)
      for (int k = 0; k < n3; ++k)
        for (int j = 0; j < n2; ++j) {
          Real *dr = dev0 + (idx_t)b * sblk + (idx_t)n * g.sn + gidx(g, k + o, j + o, o);
          if (to_device) memcpy(dr, h, n1 * sizeof(Real));
          else memcpy(h, dr, n1 * sizeof(Real));
          h += n1;
        }
#endif
  return ATHK_OK;
}

extern "C" int athk_upload_all(athk_ctx *c, int array, const double *host, int active_only) {
  return copy_all(c, array, const_cast<double *>(host), active_only, true);
}
extern "C" int athk_download_all_async(athk_ctx *c, int array, double *host, int active_only) {
  return copy_all(c, array, host, active_only, false);
}

extern "C" int athk_download_async(athk_ctx *c, int array, int lb, double *host) {
  return copy_array(c, array, lb, host, false, false);
}

extern "C" int athk_upload(athk_ctx *c, int array, int lb, const double *host) {
  return copy_array(c, array, lb, const_cast<double *>(host), true);
}
extern "C" int athk_download(athk_ctx *c, int array, int lb, double *host) {
  return copy_array(c, array, lb, host, false);
}

extern "C" int athk_device_array(athk_ctx *c, int array, void **dptr, long long *stride_block,
                                 long long *stride_var, long long *stride_k, long long *stride_j) {
  if (!c) return fail(ATHK_ERR_ARG, "null context");
  ArrDesc d;
  if (!describe(c, array, &d)) return fail(ATHK_ERR_ARG, "device_array: array %d not allocated", array);
  if (dptr) *dptr = d.base + (idx_t)d.comp0 * c->g.sn;
  if (stride_block) *stride_block = (long long)d.ncomp_total * c->g.sn;
  if (stride_var) *stride_var = c->g.sn;
  if (stride_k) *stride_k = c->g.sk;
  if (stride_j) *stride_j = c->g.sj;
  return ATHK_OK;
}

//----------------------------------------------------------------------------------------
// (a process may hold one context per GPU -- the C++ host's <mesh> num_gpus -- so every entry
//  point makes its context's device current first)
#define NEED_READY(c)                                                                      \
  do {                                                                                     \
    if (!(c)) return fail(ATHK_ERR_ARG, "null context");                                   \
    cudaSetDevice((c)->cfg.device);                                                        \
    if (!(c)->coords_set || !(c)->nbr_set)                                                 \
      return fail(ATHK_ERR_STATE, "%s: call athk_set_block_coords and athk_set_neighbors " \
                                  "first", __func__);                                      \
  } while (0)

#ifndef ATHK_EMULATE
void athk_prof_begin(athk_ctx *c, const char *name) {
  if (!c->prof_on) return;
  athk_ctx::ProfRec r;
  r.name = name;
  cudaEventCreate(&r.a);
  cudaEventCreate(&r.b);
  cudaEventRecord(r.a, c->stream);
  c->prof.push_back(r);
}
void athk_prof_end(athk_ctx *c) {
  if (c->prof_on && !c->prof.empty()) cudaEventRecord(c->prof.back().b, c->stream);
}
#else
void athk_prof_begin(athk_ctx *, const char *) {}
void athk_prof_end(athk_ctx *) {}
#endif


// streaming (HBM-bound) point kernels: compiled for 8 CTAs of 128 threads per SM (<= 64 registers)
#ifndef ATHK_STREAM_MINB
#define ATHK_STREAM_MINB 8
#endif
// the k-marching stencil kernels hold a whole plane's operands in registers (corner E: ~45 doubles)
#ifndef ATHK_CORNER_MINB
#define ATHK_CORNER_MINB 4
#endif
#ifndef ATHK_CT_MINB
#define ATHK_CT_MINB 8
#endif

static bool bad_reg(int r) { return r < 0 || r > 2; }

#ifndef ATHK_EMULATE
__global__ void k_fill(Real *p, int n, Real v) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q < n) p[q] = v;
}
#endif

static int fill_dt_blocks(athk_ctx *c) {
#ifndef ATHK_EMULATE
  k_fill<<<(c->g.nb + 255) / 256, 256, 0, c->stream>>>(c->dt_blocks, c->g.nb, DBL_MAX);
  CU(cudaGetLastError());
  c->launches++;
#else
  for (int b = 0; b < c->g.nb; ++b) c->dt_blocks[b] = DBL_MAX;
#endif
  return ATHK_OK;
}

extern "C" int athk_request_block_dt(athk_ctx *c) {
  if (!c) return fail(ATHK_ERR_ARG, "null context");
  c->dt_requested = true;
  return ATHK_OK;
}

// part 0: whole box; 1: active zone only (needs no ghost data: can run under the inter-GPU ghost
// transfer); 2: ghost shell only.
#ifndef ATHK_C2PDT_MINB
#define ATHK_C2PDT_MINB 6   // CTAs per SM of the MHD W(U) + block-dt kernel (development knob)
#endif
static int cons2prim_part(athk_ctx *c, int reg, int part);
extern "C" int athk_cons2prim(athk_ctx *c, int reg) { return cons2prim_part(c, reg, 0); }

static int cons2prim_part(athk_ctx *c, int reg, int part) {
  NEED_READY(c);
  if (bad_reg(reg) || !c->u[reg]) return fail(ATHK_ERR_ARG, "cons2prim: register %d not allocated", reg);
  const int ns = c->cfg.nscalars;
  if (ns && !c->s[reg]) return fail(ATHK_ERR_ARG, "cons2prim: scalar register %d not allocated", reg);
  const Grid &g = c->g;
  const Range3 rall = {0, g.nc3 - 1, 0, g.nc2 - 1, 0, g.nc1 - 1};
  const Range3 rin = {g.ks, g.ke, g.js, g.je, g.is, g.ie};
  Real *dt_out = nullptr;
  const bool fold = c->dt_requested && part != 2;   // the block dt is a property of the active zone
  if (part != 2) c->dt_valid = false;
  if (c->eos.sr && part != 2) {   // ph->w.SwapAthenaArray(ph->w1): write the array of two calls ago
    Real *t = c->w; c->w = c->w_alt; c->w_alt = t;
  }
  if (fold) {
    int rc = fill_dt_blocks(c);
    if (rc) return rc;
    dt_out = c->dt_blocks;
  }
  if (c->cfg.mhd) {
    Cons2PrimPoint<true> f = {g, c->eos, c->u[reg], c->b[reg], c->w, c->bcc, c->nbr, dt_out, c->geo,
                              ns ? c->s[reg] : nullptr, c->r, ns};
    if (part == 2) LAUNCH_NAMED(c, "athk_cons2prim", launch_shell(c->stream, g.nb, rall, rin, f));
    else if (fold) LAUNCH_NAMED(c, "athk_cons2prim_dt",
                                launch_points_min<ATHK_C2PDT_MINB>(c->stream, g.nb, part == 1 ? rin : rall, f, dt_out));
    else LAUNCH_NAMED(c, "athk_cons2prim",
                      launch_points_occ<ATHK_STREAM_MINB>(c->stream, g.nb, part == 1 ? rin : rall, f));
  } else if (c->eos.sr) {
    // (the fixed-point iteration wants ~100 registers: 4 CTAs per SM instead of the streaming kernels' 8 / 6)
    Cons2PrimPoint<false, true> f = {g, c->eos, c->u[reg], nullptr, c->w, nullptr, c->nbr, dt_out, c->geo,
                               ns ? c->s[reg] : nullptr, c->r, ns};
    if (part == 2) LAUNCH_NAMED(c, "athk_cons2prim", launch_shell(c->stream, g.nb, rall, rin, f));
    else if (fold) LAUNCH_NAMED(c, "athk_cons2prim_dt",
                                launch_points_min<4>(c->stream, g.nb, part == 1 ? rin : rall, f, dt_out));
    else LAUNCH_NAMED(c, "athk_cons2prim",
                      launch_points_occ<4>(c->stream, g.nb, part == 1 ? rin : rall, f));
  } else {
    Cons2PrimPoint<false> f = {g, c->eos, c->u[reg], nullptr, c->w, nullptr, c->nbr, dt_out, c->geo,
                               ns ? c->s[reg] : nullptr, c->r, ns};
    if (part == 2) LAUNCH_NAMED(c, "athk_cons2prim", launch_shell(c->stream, g.nb, rall, rin, f));
    else if (fold) LAUNCH_NAMED(c, "athk_cons2prim_dt",
                                launch_points_min<6>(c->stream, g.nb, part == 1 ? rin : rall, f, dt_out));
    else LAUNCH_NAMED(c, "athk_cons2prim",
                      launch_points_occ<ATHK_STREAM_MINB>(c->stream, g.nb, part == 1 ? rin : rall, f));
  }
  if (fold) { c->dt_valid = true; c->dt_requested = false; }
  return ATHK_OK;
}

static int check_order(athk_ctx *c, int order, const char *who) {
  if (order < 1 || order > 3)
    return fail(ATHK_ERR_ARG, "%s: xorder %d not implemented (1 donor cell, 2 PLM, 3 PPM)", who, order);
  if (order == 3 && c->cfg.nghost < 3)   // reconstruction.cpp:117-127
    return fail(ATHK_ERR_ARG, "%s: xorder=3 (PPM) needs nghost >= 3 (context has %d)", who, c->cfg.nghost);
  return ATHK_OK;
}

extern "C" int athk_calculate_fluxes(athk_ctx *c, int order, double dt_mesh, int breg) {
  NEED_READY(c);
  if (c->cfg.mhd && (bad_reg(breg) || !c->b[breg]))
    return fail(ATHK_ERR_ARG, "calculate_fluxes: field register %d not allocated", breg);
  int rc = check_order(c, order, "calculate_fluxes");
  if (rc) return rc;
  const bool mhd = c->cfg.mhd != 0;
  if (c->eos.sr) return athk_flux3_rs<RS_SR, false>(c, order, dt_mesh, breg);
  if (c->eos.iso && c->cfg.rsolver == (mhd ? ATHK_RS_HLLD : ATHK_RS_HLLE))
    return mhd ? athk_flux3_rs<RS_HLLD_ISO, true>(c, order, dt_mesh, breg) : athk_flux3_rs<RS_HLLE_ISO, false>(c, order, dt_mesh, breg);
  switch (c->cfg.rsolver) {
    case ATHK_RS_ROE: return mhd ? athk_flux3_rs<RS_ROE, true>(c, order, dt_mesh, breg) : athk_flux3_rs<RS_ROE, false>(c, order, dt_mesh, breg);
    case ATHK_RS_HLLE: return mhd ? athk_flux3_rs<RS_HLLE, true>(c, order, dt_mesh, breg) : athk_flux3_rs<RS_HLLE, false>(c, order, dt_mesh, breg);
    case ATHK_RS_LLF: return mhd ? athk_flux3_rs<RS_LLF, true>(c, order, dt_mesh, breg) : athk_flux3_rs<RS_LLF, false>(c, order, dt_mesh, breg);
    default: return mhd ? athk_flux3_rs<RS_HLLD, true>(c, order, dt_mesh, breg) : athk_flux3_rs<RS_HLLC, false>(c, order, dt_mesh, breg);
  }
}

// part 0: every edge; 1: the block's boundary surface only (what the EMF exchange needs);
// 2: the rest.  1 + 2 == 0; the split lets the inter-GPU EMF transfer run under part 2.
static int corner_e_part(athk_ctx *c, int part) {
  NEED_READY(c);
  if (!c->cfg.mhd) return fail(ATHK_ERR_ARG, "compute_corner_e: context is not MHD");
  const Grid &g = c->g;
  const Range3 r = {g.ks, g.ke + 1, g.js, g.je + 1, g.is, g.ie + 1};
  const Range3 rin = {g.ks + 1, g.ke, g.js + 1, g.je, g.is + 1, g.ie};
  CornerEPoint f = {g, c->w, c->bcc, c->ef, c->wct, c->e};
  if (part == 1) LAUNCH_NAMED(c, "athk_compute_corner_e", launch_shell(c->stream, g.nb, r, rin, f));
  else LAUNCH_NAMED(c, "athk_compute_corner_e",
                    launch_points_kcol<ATHK_CORNER_MINB>(c->stream, g.nb, part == 2 ? rin : r, c->kt_corner, f));
  return ATHK_OK;
}

extern "C" int athk_compute_corner_e(athk_ctx *c) { return corner_e_part(c, 0); }

extern "C" int athk_send_flux_correction(athk_ctx *c) {
  NEED_READY(c);
  if (!c->cfg.mhd) return fail(ATHK_ERR_ARG, "send_flux_correction: context is not MHD");
  const Grid &g = c->g;
  const Range3 r = {g.ks, g.ke + 1, g.js, g.je + 1, g.is, g.ie + 1};   // boundary surface only
  const Range3 rin = {g.ks + 1, g.ke, g.js + 1, g.je, g.is + 1, g.ie};
  EmfPackPoint f = {g, c->sf, c->e, c->emfbuf};
  LAUNCH_NAMED(c, __func__, launch_shell(c->stream, g.nb, r, rin, f));
  return ATHK_OK;
}

extern "C" int athk_recv_flux_correction(athk_ctx *c) {
  NEED_READY(c);
  if (!c->cfg.mhd) return fail(ATHK_ERR_ARG, "recv_flux_correction: context is not MHD");
  const Grid &g = c->g;
  const Range3 r = {g.ks, g.ke + 1, g.js, g.je + 1, g.is, g.ie + 1};
  const Range3 rin = {g.ks + 1, g.ke, g.js + 1, g.je, g.is + 1, g.ie};
  { int rcw = athk_p2p_before_emf_apply(c); if (rcw) return rcw; }
  EmfApplyPoint f = {g, c->sf, c->e, c->emfbuf, c->nbr, c->rl.erbuf, c->rl.ebase};
  LAUNCH_NAMED(c, __func__, launch_shell(c->stream, g.nb, r, rin, f));
  return ATHK_OK;
}

// weighted_ave.cpp branch structure -> kernel mode (see WeightedAvePoint)
static int wa_mode(const double w[5], int *mode) {
  if (w[3] != 0.0 || w[4] != 0.0) return -1;
  if (w[0] == 1.0) {
    if (w[2] != 0.0) *mode = 0;
    else if (w[1] != 0.0) *mode = 1;
    else *mode = -2;   // nothing to do
  } else if (w[0] == 0.0) {
    if (w[2] != 0.0) *mode = 2;
    else if (w[1] != 0.0) *mode = 3;
    else *mode = 7;    // ZeroClear
  } else {
    if (w[2] != 0.0) *mode = 4;
    else if (w[1] != 0.0) *mode = 5;
    else *mode = 6;
  }
  return 0;
}

static int weighted_ave(athk_ctx *c, Real *out, const Real *in1, const Real *in2, int nvar, int fc,
                        const double w[5]) {
  int mode;
  if (wa_mode(w, &mode))
    return fail(ATHK_ERR_ARG, "weighted_ave: wght[3], wght[4] != 0 (fourth and fifth input of MeshBlock::WeightedAve) not implemented");
  if (mode == -2) return ATHK_OK;
  const Grid &g = c->g;
  if ((mode == 0 || mode == 2 || mode == 4) && !in2) return fail(ATHK_ERR_ARG, "weighted_ave: in2 register not allocated");
  if (mode != 6 && mode != 7 && !in1) return fail(ATHK_ERR_ARG, "weighted_ave: in1 register not allocated");
  Range3 r = {g.ks, g.ke + fc, g.js, g.je + fc, g.is, g.ie + fc};
  if (mode == 7) {   // all weights zero: the reference takes its `u_out = wght[1]*u_in1` branch
                     // (weighted_ave.cpp:135-158), i.e. the register is zeroed
    WeightedAvePoint f = {g, nvar, fc, 3, 0.0, 0.0, 0.0, out, in1 ? in1 : out, out};
    LAUNCH(c, r, f);
    return ATHK_OK;
  }
  WeightedAvePoint f = {g, nvar, fc, mode, w[0], w[1], w[2], out, in1 ? in1 : out, in2 ? in2 : out};
  LAUNCH(c, r, f);
  return ATHK_OK;
}

extern "C" int athk_weighted_ave_cc(athk_ctx *c, int what, int out_reg, int in1_reg, int in2_reg,
                                    const double wght[5]) {
  NEED_READY(c);
  if (bad_reg(out_reg) || bad_reg(in1_reg) || bad_reg(in2_reg) || !wght)
    return fail(ATHK_ERR_ARG, "weighted_ave_cc: bad register or null weights");
  Real **regs = (what == 0) ? c->u : c->s;
  const int nvar = (what == 0) ? 5 : c->cfg.nscalars;
  if (what != 0 && what != 1) return fail(ATHK_ERR_ARG, "weighted_ave_cc: what must be 0 (u) or 1 (s)");
  if (what == 1 && nvar == 0) return fail(ATHK_ERR_ARG, "weighted_ave_cc: the context has no passive scalars");
  if (!regs[out_reg] && out_reg == 2) {
    // the third register (u2 / s2 of the 3S* integrators, hydro.cpp:76-80) exists from its first use on
    int rca = dev_alloc(&regs[2], cc_elems(c, nvar), c->stream);
    if (rca) return rca;
  }
  if (!regs[out_reg]) return fail(ATHK_ERR_ARG, "weighted_ave_cc: out register %d not allocated", out_reg);
  return weighted_ave(c, regs[out_reg], regs[in1_reg], regs[in2_reg], nvar, 0, wght);
}

extern "C" int athk_weighted_ave_fc(athk_ctx *c, int out_reg, int in1_reg, int in2_reg,
                                    const double wght[5]) {
  NEED_READY(c);
  if (!c->cfg.mhd) return fail(ATHK_ERR_ARG, "weighted_ave_fc: context is not MHD");
  if (bad_reg(out_reg) || bad_reg(in1_reg) || bad_reg(in2_reg) || !wght)
    return fail(ATHK_ERR_ARG, "weighted_ave_fc: bad register or null weights");
  if (!c->b[out_reg]) return fail(ATHK_ERR_ARG, "weighted_ave_fc: out register %d not allocated", out_reg);
  return weighted_ave(c, c->b[out_reg], c->b[in1_reg], c->b[in2_reg], 3, 1, wght);
}

extern "C" int athk_swap_registers(athk_ctx *c, int what, int a, int b) {
  if (!c) return fail(ATHK_ERR_ARG, "null context");
  if (bad_reg(a) || bad_reg(b)) return fail(ATHK_ERR_ARG, "swap_registers: bad register");
  Real **regs = (what == 0) ? c->u : (what == 1 ? c->b : (what == 2 ? c->s : nullptr));
  if (!regs) return fail(ATHK_ERR_ARG, "swap_registers: what must be 0 (u), 1 (b) or 2 (s)");
  Real *t = regs[a];
  regs[a] = regs[b];
  regs[b] = t;
  return ATHK_OK;
}

extern "C" int athk_zero_register(athk_ctx *c, int what, int reg) {
  if (!c) return fail(ATHK_ERR_ARG, "null context");
  cudaSetDevice(c->cfg.device);
  if (bad_reg(reg)) return fail(ATHK_ERR_ARG, "zero_register: bad register");
  Real *p = (what == 0) ? c->u[reg] : (what == 1 ? c->b[reg] : (what == 2 ? c->s[reg] : nullptr));
  const int nvar = (what == 0) ? 5 : (what == 1 ? 3 : c->cfg.nscalars);
  if (!p) return fail(ATHK_ERR_ARG, "zero_register: register not allocated");
  CU(cudaMemsetAsync(p, 0, cc_elems(c, nvar) * sizeof(Real), c->stream));
  return ATHK_OK;
}

extern "C" int athk_add_flux_divergence(athk_ctx *c, double wght, int reg) {
  NEED_READY(c);
  if (bad_reg(reg) || !c->u[reg]) return fail(ATHK_ERR_ARG, "add_flux_divergence: register %d not allocated", reg);
  const Grid &g = c->g;
  Range3 r = {g.ks, g.ke, g.js, g.je, g.is, g.ie};
  FluxDivPoint f = {g, c->geo, 5, wght, c->flux[0], c->flux[1], c->flux[2], c->u[reg]};
  LAUNCH(c, r, f);
  return ATHK_OK;
}

extern "C" int athk_ct(athk_ctx *c, double wght, int reg) {
  NEED_READY(c);
  if (!c->cfg.mhd) return fail(ATHK_ERR_ARG, "ct: context is not MHD");
  if (bad_reg(reg) || !c->b[reg]) return fail(ATHK_ERR_ARG, "ct: register %d not allocated", reg);
  const Grid &g = c->g;
  Range3 r = {g.ks, g.ke + 1, g.js, g.je + 1, g.is, g.ie + 1};
  CTPoint f = {g, c->geo, wght, c->e, c->b[reg]};
  LAUNCH(c, r, f);
  return ATHK_OK;
}

// mode 0: all neighbours; 1: neighbours on this GPU only; 2: received remote messages only
static int exchange_ghosts_mode(athk_ctx *c, int what_mask, int reg, int mode);
extern "C" int athk_exchange_ghosts(athk_ctx *c, int what_mask, int reg) {
  return exchange_ghosts_mode(c, what_mask, reg, 0);
}

static int exchange_ghosts_mode(athk_ctx *c, int what_mask, int reg, int mode) {
  NEED_READY(c);
  if (mode == 2 && c->rl.n == 0) return ATHK_OK;
  if (bad_reg(reg)) return fail(ATHK_ERR_ARG, "exchange_ghosts: bad register");
  if (mode != 1) { int rcw = athk_p2p_before_ghost_fill(c); if (rcw) return rcw; }
  const Grid &g = c->g;
  const Range3 rin = {g.ks, g.ke, g.js, g.je, g.is, g.ie};   // only the ghost shell is launched
  if (what_mask & 1) {
    if (!c->u[reg]) return fail(ATHK_ERR_ARG, "exchange_ghosts: u register %d not allocated", reg);
    Range3 r = {0, g.nc3 - 1, 0, g.nc2 - 1, 0, g.nc1 - 1};
    if (c->rl.n > 0) {
      GhostFillCC<true> f = {g, 5, c->u[reg], c->nbr, c->cfg.mhd, c->cfg.nscalars, 0, c->rl, mode};
      LAUNCH_NAMED(c, "athk_exchange_ghosts", launch_shell(c->stream, g.nb, r, rin, f));
    } else {
      GhostFillCC<false> f = {g, 5, c->u[reg], c->nbr, c->cfg.mhd, c->cfg.nscalars, 0, c->rl, mode};
      LAUNCH_NAMED(c, "athk_exchange_ghosts", launch_shell(c->stream, g.nb, r, rin, f));
    }
  }
  if (what_mask & 2) {
    if (!c->cfg.mhd || !c->b[reg]) return fail(ATHK_ERR_ARG, "exchange_ghosts: b register %d not allocated", reg);
    Range3 r = {0, g.nc3, 0, g.nc2, 0, g.nc1};
    if (c->rl.n > 0) {
      GhostFillFC<true> f = {g, c->b[reg], c->nbr, c->cfg.nscalars, c->rl, mode};
      LAUNCH_NAMED(c, "athk_exchange_ghosts", launch_shell(c->stream, g.nb, r, rin, f));
    } else {
      GhostFillFC<false> f = {g, c->b[reg], c->nbr, c->cfg.nscalars, c->rl, mode};
      LAUNCH_NAMED(c, "athk_exchange_ghosts", launch_shell(c->stream, g.nb, r, rin, f));
    }
  }
  if (what_mask & 4) {
    if (!c->s[reg]) return fail(ATHK_ERR_ARG, "exchange_ghosts: s register %d not allocated", reg);
    Range3 r = {0, g.nc3 - 1, 0, g.nc2 - 1, 0, g.nc1 - 1};
    if (c->rl.n > 0) {
      GhostFillCC<true> f = {g, c->cfg.nscalars, c->s[reg], c->nbr, c->cfg.mhd, c->cfg.nscalars, 4, c->rl, mode};
      LAUNCH_NAMED(c, "athk_exchange_ghosts", launch_shell(c->stream, g.nb, r, rin, f));
    } else {
      GhostFillCC<false> f = {g, c->cfg.nscalars, c->s[reg], c->nbr, c->cfg.mhd, c->cfg.nscalars, 4, c->rl, mode};
      LAUNCH_NAMED(c, "athk_exchange_ghosts", launch_shell(c->stream, g.nb, r, rin, f));
    }
  }
  return ATHK_OK;
}

//----------------------------------------------------------------------------------------
// Hydro::NewBlockTimeStep: block-wide min reduction.  min() is exact and order independent,
// so a tree reduction gives the reference's serial result bit for bit.
#ifndef ATHK_EMULATE
template <bool MHD>
__global__ void __launch_bounds__(256) k_block_dt(const CellDt<MHD> f, Real *dt_blocks) {
  const Grid &g = f.g;
  const int blk = blockIdx.y;
  const int ncell = g.nx1 * g.nx2 * g.nx3;
  Real m = DBL_MAX;
  for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < ncell; q += gridDim.x * blockDim.x) {
    const int i = q % g.nx1;
    const int t = q / g.nx1;
    const int j = t % g.nx2;
    const int k = t / g.nx2;
    m = min_std(m, f(blk, g.ks + k, g.js + j, g.is + i));
  }
  for (int o = 16; o > 0; o >>= 1) m = min_std(m, __shfl_xor_sync(0xffffffffu, m, o));
  __shared__ Real sm[8];
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 8) {
    m = sm[threadIdx.x];
    for (int o = 4; o > 0; o >>= 1) m = min_std(m, __shfl_xor_sync(0xffu, m, o));
    // positive doubles order like their bit patterns
    if (threadIdx.x == 0)
      atomicMin((unsigned long long *)&dt_blocks[blk], (unsigned long long)__double_as_longlong(m));
  }
}
#endif

extern "C" int athk_new_block_dt(athk_ctx *c, double cfl_number, double *dt_blocks, double *dt_min) {
  NEED_READY(c);
  const Grid &g = c->g;
  std::vector<Real> h(g.nb);
  if (c->dt_valid) {   // already folded by the last W(U) (athk_request_block_dt)
    CU(cudaMemcpyAsync(h.data(), c->dt_blocks, g.nb * sizeof(Real), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    c->dt_valid = false;
  } else {
#ifndef ATHK_EMULATE
  k_fill<<<(g.nb + 255) / 256, 256, 0, c->stream>>>(c->dt_blocks, g.nb, DBL_MAX);
  const int ncell = g.nx1 * g.nx2 * g.nx3;
  int gx = (ncell + 256 * 8 - 1) / (256 * 8);
  if (gx < 1) gx = 1;
  dim3 grid(gx, g.nb);
  if (c->cfg.mhd) {
    CellDt<true> f = {g, c->eos, c->geo, c->w, c->bcc, c->b[0]};
    k_block_dt<true><<<grid, 256, 0, c->stream>>>(f, c->dt_blocks);
  } else {
    CellDt<false> f = {g, c->eos, c->geo, c->w, nullptr, nullptr};
    k_block_dt<false><<<grid, 256, 0, c->stream>>>(f, c->dt_blocks);
  }
  CU(cudaGetLastError());
  c->launches += 2;
  CU(cudaMemcpyAsync(h.data(), c->dt_blocks, g.nb * sizeof(Real), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
#else
  for (int b = 0; b < g.nb; ++b) {
    Real m = DBL_MAX;
    for (int k = g.ks; k <= g.ke; ++k)
      for (int j = g.js; j <= g.je; ++j)
        for (int i = g.is; i <= g.ie; ++i) {
          if (c->cfg.mhd) {
            CellDt<true> f = {g, c->eos, c->geo, c->w, c->bcc, c->b[0]};
            m = min_std(m, f(b, k, j, i));
          } else {
            CellDt<false> f = {g, c->eos, c->geo, c->w, nullptr, nullptr};
            m = min_std(m, f(b, k, j, i));
          }
        }
    h[b] = m;
  }
#endif
  }
  Real mn = DBL_MAX;
  for (int b = 0; b < g.nb; ++b) {
    h[b] *= cfl_number;   // new_blockdt.cpp:191
    if (dt_blocks) dt_blocks[b] = h[b];
    mn = min_std(mn, h[b]);
  }
  if (dt_min) *dt_min = mn;
  return ATHK_OK;
}

//----------------------------------------------------------------------------------------
// Passive scalars (athk_scalars.h)
template <int ORDER>
static int launch_scalar_fluxes(athk_ctx *c) {
  const Grid &g = c->g;
  const int ns = c->cfg.nscalars;
  const MassFlux mf = {{c->flux[0], c->flux[1], c->flux[2]}, (idx_t)5 * g.sn};   // Hydro::flux[d](IDN)
  {
    Range3 r = {g.ks, g.ke, g.js, g.je, g.is, g.ie + 1};
    ScalarFacePoint<0, ORDER> f = {g, c->geo, c->eos, ns, c->r, mf, c->sflux[0]};
    LAUNCH_NAMED(c, "calculate_scalar_fluxes_x1", launch_points(c->stream, g.nb, r, f));
  }
  {
    Range3 r = {g.ks, g.ke, g.js, g.je + 1, g.is, g.ie};
    ScalarFacePoint<1, ORDER> f = {g, c->geo, c->eos, ns, c->r, mf, c->sflux[1]};
    LAUNCH_NAMED(c, "calculate_scalar_fluxes_x2", launch_points(c->stream, g.nb, r, f));
  }
  {
    Range3 r = {g.ks, g.ke + 1, g.js, g.je, g.is, g.ie};
    ScalarFacePoint<2, ORDER> f = {g, c->geo, c->eos, ns, c->r, mf, c->sflux[2]};
    LAUNCH_NAMED(c, "calculate_scalar_fluxes_x3", launch_points(c->stream, g.nb, r, f));
  }
  return ATHK_OK;
}

extern "C" int athk_calculate_scalar_fluxes(athk_ctx *c, int order) {
  NEED_READY(c);
  if (!c->cfg.nscalars) return fail(ATHK_ERR_ARG, "calculate_scalar_fluxes: context has no passive scalars");
  int rc = check_order(c, order, "calculate_scalar_fluxes");
  if (rc) return rc;
  if (order == 1) return launch_scalar_fluxes<1>(c);
  if (order == 2) return launch_scalar_fluxes<2>(c);
  return launch_scalar_fluxes<3>(c);
}

extern "C" int athk_add_scalar_flux_divergence(athk_ctx *c, double wght, int reg) {
  NEED_READY(c);
  if (!c->cfg.nscalars) return fail(ATHK_ERR_ARG, "add_scalar_flux_divergence: context has no passive scalars");
  if (bad_reg(reg) || !c->s[reg]) return fail(ATHK_ERR_ARG, "add_scalar_flux_divergence: register %d not allocated", reg);
  const Grid &g = c->g;
  Range3 r = {g.ks, g.ke, g.js, g.je, g.is, g.ie};
  FluxDivPoint f = {g, c->geo, c->cfg.nscalars, wght, c->sflux[0], c->sflux[1], c->sflux[2], c->s[reg]};
  LAUNCH(c, r, f);
  return ATHK_OK;
}

template <int ORDER>
static int launch_scalar_integrate(athk_ctx *c, const HydroIntegrate &hi) {
  const Grid &g = c->g;
  const int ns = c->cfg.nscalars;
  const MassFlux mf = {{c->sflux[0], c->sflux[1], c->sflux[2]}, (idx_t)ns * g.sn};   // left by the sweeps
  Range3 r = {g.ks, g.ke, g.js, g.je, g.is, g.ie};
  ScalarIntegratePoint<ORDER> f = {g, c->geo, c->eos, ns, c->r, mf, hi};
  LAUNCH_NAMED(c, "scalar_fluxes_integrate", launch_points(c->stream, g.nb, r, f));
  return ATHK_OK;
}
//----------------------------------------------------------------------------------------
// Physical boundary conditions (athk_bcs.h)
extern "C" int athk_set_boundary_flags(athk_ctx *c, int lb, const int flags[6]) {
  if (!c || !flags) return fail(ATHK_ERR_ARG, "set_boundary_flags: null argument");
  if (lb < 0 || lb >= c->g.nb) return fail(ATHK_ERR_ARG, "set_boundary_flags: block %d out of range", lb);
  for (int f = 0; f < 6; ++f) {
    if (flags[f] < ATHK_BC_NONE || flags[f] > ATHK_BC_USER)
      return fail(ATHK_ERR_ARG, "set_boundary_flags: face %d has unknown flag %d (reflecting, outflow, user or none)",
                  f, flags[f]);
    c->h_bcflags[(size_t)lb * 6 + f] = flags[f];
  }
  c->bc_dirty = true;
  return ATHK_OK;
}

template <int DIR>
static int launch_bc_face(athk_ctx *c, int side) {
  const Grid &g = c->g;
  // the DIR coordinate runs over the ghost layer m = 1..ng, the others over the padded box
  Range3 r = {0, g.nc3, 0, g.nc2, 0, g.nc1};
  if (DIR == 0) { r.il = 1; r.iu = g.ng; }
  if (DIR == 1) { r.jl = 1; r.ju = g.ng; }
  if (DIR == 2) { r.kl = 1; r.ku = g.ng; }
  const int mhd = c->cfg.mhd, ns = c->cfg.nscalars;
  PhysBCFillPoint<DIR> f1 = {g, side, c->bcflags, mhd, ns, c->w, mhd ? c->b[0] : nullptr, c->r};
  LAUNCH_NAMED(c, "apply_physical_boundaries", launch_points_occ<ATHK_STREAM_MINB>(c->stream, g.nb, r, f1));
  // user-enrolled functions of this face (bvals.cpp:673-676), block after block
  const int face = 2 * DIR + side;
  for (int lb = 0; lb < g.nb; ++lb) {
    const int *fl = &c->h_bcflags[(size_t)lb * 6];
    if (fl[face] != ATHK_BC_USER) continue;
    if (!c->user_bc[face])   // bvals.cpp:349-356
      return fail(ATHK_ERR_STATE, "A user-defined boundary is specified but the actual boundary function "
                                  "is not enrolled in direction %d (athk_enroll_user_boundary)", face);
    athk_bval_args a;
    a.lb = lb; a.face = face; a.time = c->bc_time; a.dt = c->bc_dt; a.ngh = g.ng;
    a.il = g.is; a.iu = g.ie; a.jl = g.js; a.ju = g.je; a.kl = g.ks; a.ku = g.ke;
    // transverse limits of bvals.cpp:457-470 and of the x2 / x3 passes
    if (DIR == 0) {
      if (fl[2] == ATHK_BC_NONE) a.jl -= g.ng;
      if (fl[3] == ATHK_BC_NONE) a.ju += g.ng;
    } else {
      a.il -= g.ng; a.iu += g.ng;
    }
    if (DIR != 2) {
      if (fl[4] == ATHK_BC_NONE) a.kl -= g.ng;
      if (fl[5] == ATHK_BC_NONE) a.ku += g.ng;
    } else {
      a.jl = g.js - g.ng; a.ju = g.je + g.ng;
    }
    a.w = c->w + (idx_t)lb * 5 * g.sn;
    a.b1 = mhd ? c->b[0] + (idx_t)lb * 3 * g.sn : nullptr;
    a.b2 = mhd ? a.b1 + g.sn : nullptr;
    a.b3 = mhd ? a.b1 + 2 * g.sn : nullptr;
    a.r = ns ? c->r + (idx_t)lb * ns * g.sn : nullptr;
    a.stride_j = g.sj; a.stride_k = g.sk; a.stride_n = g.sn;
    a.nscalars = ns;
    a.stream = (void *)c->stream;
    if (c->user_bc[face](c->user_bc_data[face], &a) != 0)
      return fail(ATHK_ERR_STATE, "user boundary function of face %d failed on block %d", face, lb);
  }
  Range3 rc = {0, g.nc3 - 1, 0, g.nc2 - 1, 0, g.nc1 - 1};
  if (DIR == 0) { rc.il = 1; rc.iu = g.ng; }
  if (DIR == 1) { rc.jl = 1; rc.ju = g.ng; }
  if (DIR == 2) { rc.kl = 1; rc.ku = g.ng; }
  PhysBCConsPoint<DIR> f2 = {g, side, c->bcflags, c->eos, mhd, ns, c->w, mhd ? c->b[0] : nullptr, c->r,
                             c->bcc, c->u[0], ns ? c->s[0] : nullptr};
  LAUNCH_NAMED(c, "apply_physical_boundaries", launch_points_occ<ATHK_STREAM_MINB>(c->stream, g.nb, rc, f2));
  return ATHK_OK;
}

//----------------------------------------------------------------------------------------
// Static mesh refinement, cell-centred variables (athk_refine.h)
extern "C" int athk_refine_enable(athk_ctx *c) {
  if (!c) return fail(ATHK_ERR_ARG, "refine_enable: null context");
  cudaSetDevice(c->cfg.device);
  if (c->rf.on) return ATHK_OK;
  if (c->eos.sr) return fail(ATHK_ERR_ARG, "refine_enable: mesh refinement is not implemented with special relativity");
  const Grid &g = c->g;
  if (g.nx1 % 2 || g.nx2 % 2 || g.nx3 % 2)   // mesh.cpp:340-346
    return fail(ATHK_ERR_ARG, "The size of MeshBlock must be divisible by 2 in order to use SMR or AMR.");
  CoarseGrid &cg = c->rf.cg;
  cg.cng = (g.ng + 1) / 2 + 1;
  cg.cn1 = g.nx1 / 2 + 2 * g.ng; cg.cn2 = g.nx2 / 2 + 2 * g.ng; cg.cn3 = g.nx3 / 2 + 2 * g.ng;
  cg.cis = cg.cjs = cg.cks = g.ng;
  cg.csj = cg.cn1; cg.csk = (idx_t)cg.cn1 * cg.cn2; cg.csn = cg.csk * cg.cn3;
  cg.CPG = std::max(cg.cn1, std::max(cg.cn2, cg.cn3));
  cg.cfj = cg.cn1 + 1; cg.cfk = (idx_t)(cg.cn1 + 1) * (cg.cn2 + 1); cg.cfn = cg.cfk * (cg.cn3 + 1);
  const size_t nc = (size_t)g.nb * 5 * cg.csn;
  int rc = dev_alloc(&c->rf.cu, nc, c->stream);
  if (!rc) rc = dev_alloc(&c->rf.cw, nc, c->stream);
  if (!rc) rc = dev_alloc(&c->rf.xv, (size_t)g.nb * 3 * c->geo.PG, c->stream);
  if (!rc) rc = dev_alloc(&c->rf.cxv, (size_t)g.nb * 3 * cg.CPG, c->stream);
  if (!rc && c->cfg.nscalars) rc = dev_alloc(&c->rf.cs, (size_t)g.nb * c->cfg.nscalars * cg.csn, c->stream);   // coarse_s_
  if (!rc && c->cfg.nscalars) rc = dev_alloc(&c->rf.cr, (size_t)g.nb * c->cfg.nscalars * cg.csn, c->stream);   // coarse_r_
  if (!rc && c->cfg.mhd) rc = dev_alloc(&c->rf.cb, (size_t)g.nb * 3 * cg.cfn, c->stream);     // coarse_b_ (field/field.cpp:60-70)
  if (!rc && c->cfg.mhd) rc = dev_alloc(&c->rf.cbcc, (size_t)g.nb * 3 * cg.csn, c->stream);   // coarse_bcc_
  if (rc) return rc;
  c->rf.on = true;
  return ATHK_OK;
}

extern "C" int athk_refine_set_block_coords(athk_ctx *c, int lb, const double *x1v, const double *x2v,
                                            const double *x3v, const double *cx1v, const double *cx2v,
                                            const double *cx3v) {
  if (!c || !c->rf.on) return fail(ATHK_ERR_STATE, "refine_set_block_coords: call athk_refine_enable first");
  cudaSetDevice(c->cfg.device);
  if (lb < 0 || lb >= c->g.nb) return fail(ATHK_ERR_ARG, "refine_set_block_coords: block %d out of range", lb);
  const double *xv[3] = {x1v, x2v, x3v}, *cxv[3] = {cx1v, cx2v, cx3v};
  const int nc[3] = {c->g.nc1, c->g.nc2, c->g.nc3}, cn[3] = {c->rf.cg.cn1, c->rf.cg.cn2, c->rf.cg.cn3};
  for (int d = 0; d < 3; ++d) {
    if (!xv[d] || !cxv[d]) return fail(ATHK_ERR_ARG, "refine_set_block_coords: null array");
    CU(cudaMemcpyAsync(c->rf.xv + ((size_t)lb * 3 + d) * c->geo.PG, xv[d], nc[d] * sizeof(Real),
                       cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->rf.cxv + ((size_t)lb * 3 + d) * c->rf.cg.CPG, cxv[d], cn[d] * sizeof(Real),
                       cudaMemcpyHostToDevice, c->stream));
  }
  CU(cudaStreamSynchronize(c->stream));
  return ATHK_OK;
}

extern "C" int athk_refine_set_peers(athk_ctx *c, int n, athk_ctx *const *peers) {
  if (!c || !c->rf.on) return fail(ATHK_ERR_STATE, "refine_set_peers: call athk_refine_enable first");
  if (n < 0 || n > ATHK_MAX_PEERS || (n > 0 && !peers)) return fail(ATHK_ERR_ARG, "refine_set_peers: %d contexts (0..%d)", n, ATHK_MAX_PEERS);
  for (int r = 0; r < ATHK_MAX_PEERS; ++r) c->rf.peers[r] = nullptr;
  for (int r = 0; r < n; ++r) {
    athk_ctx *p = peers[r];
    if (!p || p == c) continue;
    if (!p->rf.on) return fail(ATHK_ERR_STATE, "refine_set_peers: context %d has no refinement buffers", r);
    if (p->g.nx1 != c->g.nx1 || p->g.nx2 != c->g.nx2 || p->g.nx3 != c->g.nx3 || p->g.ng != c->g.ng ||
        p->cfg.mhd != c->cfg.mhd || p->cfg.nscalars != c->cfg.nscalars || p->geo.PG != c->geo.PG)
      return fail(ATHK_ERR_ARG, "refine_set_peers: context %d has another block shape or variable set", r);
    c->rf.peers[r] = p;
  }
  c->rf.npeers = n;
  return ATHK_OK;
}

extern "C" int athk_refine_coarse_transfer(athk_ctx *c, int lb, int to_device, double *cprim, double *cb1, double *cb2,
                                           double *cb3) {
  if (!c || !c->rf.on) return fail(ATHK_ERR_STATE, "refine_coarse_transfer: call athk_refine_enable first");
  cudaSetDevice(c->cfg.device);
  if (lb < 0 || lb >= c->g.nb || !cprim) return fail(ATHK_ERR_ARG, "refine_coarse_transfer: block %d or null array", lb);
  const CoarseGrid &cg = c->rf.cg;
  CU(cudaStreamSynchronize(c->stream));
  Real *dw = c->rf.cw + (size_t)lb * 5 * cg.csn;   // dense already: csj = cn1, csk = cn1 * cn2
  if (to_device) CU(cudaMemcpy(dw, cprim, (size_t)5 * cg.csn * sizeof(Real), cudaMemcpyHostToDevice));
  else CU(cudaMemcpy(cprim, dw, (size_t)5 * cg.csn * sizeof(Real), cudaMemcpyDeviceToHost));
  double *hb[3] = {cb1, cb2, cb3};
  if (c->cfg.mhd && (cb1 || cb2 || cb3)) {
    // the device keeps every component in one (cn1+1, cn2+1, cn3+1) box: repack on the host
    std::vector<Real> tmp((size_t)cg.cfn);
    for (int d = 0; d < 3; ++d) {
      if (!hb[d]) continue;
      const int n1 = cg.cn1 + (d == 0), n2 = cg.cn2 + (d == 1), n3 = cg.cn3 + (d == 2);
      Real *db = c->rf.cb + ((size_t)lb * 3 + d) * cg.cfn;
      CU(cudaMemcpy(tmp.data(), db, (size_t)cg.cfn * sizeof(Real), cudaMemcpyDeviceToHost));
      for (int k = 0; k < n3; ++k)
        for (int j = 0; j < n2; ++j)
          for (int i = 0; i < n1; ++i) {
            Real &dev = tmp[(size_t)cfidx(cg, k, j, i)];
            double &host = hb[d][((size_t)k * n2 + j) * n1 + i];
            if (to_device) dev = host; else host = dev;
          }
      if (to_device) CU(cudaMemcpy(db, tmp.data(), (size_t)cg.cfn * sizeof(Real), cudaMemcpyHostToDevice));
    }
  }
  return ATHK_OK;
}

extern "C" int athk_refine_set_ops(athk_ctx *c, int phase, const athk_box_op *ops, int nops) {
  if (!c || !c->rf.on) return fail(ATHK_ERR_STATE, "refine_set_ops: call athk_refine_enable first");
  cudaSetDevice(c->cfg.device);
  if (phase < 0 || phase >= ATHK_REFINE_NPHASES) return fail(ATHK_ERR_ARG, "refine_set_ops: phase %d", phase);
  if (nops < 0 || (nops > 0 && !ops)) return fail(ATHK_ERR_ARG, "refine_set_ops: bad list");
  static_assert(sizeof(athk_box_op) == sizeof(BoxOp), "athk_box_op layout");
  const Grid &g = c->g;
  const CoarseGrid &cg = c->rf.cg;
  long long maxvol = 0;
  for (int n = 0; n < nops; ++n) {   // every box must lie inside the array it addresses
    const athk_box_op &o = ops[n];
    if (o.peer < 0 || o.peer > ATHK_MAX_PEERS || (o.peer > 0 && !c->rf.peers[o.peer - 1]))
      return fail(ATHK_ERR_ARG, "refine_set_ops: box %d names GPU %d, which athk_refine_set_peers did not give", n, o.peer - 1);
    const int src_nb = o.peer ? c->rf.peers[o.peer - 1]->g.nb : g.nb;
    if (o.kind < BOX_RESTRICT || o.kind > BOX_COARSE_BC_FC || o.dst < 0 || o.dst >= g.nb || o.src < 0 ||
        (o.kind != BOX_COARSE_BC && o.kind != BOX_EMF_SCALE && o.kind != BOX_COARSE_BC_FC && o.src >= src_nb))
      return fail(ATHK_ERR_ARG, "refine_set_ops: box %d has kind %d, blocks %d <- %d", n, o.kind, o.dst, o.src);
    if (o.kind < BOX_FC_COPY && o.pad != 0 && (o.pad != 1 || !c->cfg.nscalars))
      return fail(ATHK_ERR_ARG, "refine_set_ops: box %d (kind %d): variable set %d", n, o.kind, o.pad);
    if (o.kind >= BOX_FC_COPY) {
      if (!c->cfg.mhd) return fail(ATHK_ERR_ARG, "refine_set_ops: box %d (kind %d) needs an MHD context", n, o.kind);
      if (o.kind != BOX_FC_PROLONG_INTERNAL && (o.dir < 0 || o.dir > 2))
        return fail(ATHK_ERR_ARG, "refine_set_ops: box %d (kind %d): component %d", n, o.kind, o.dir);
      if (o.kind == BOX_EMF_SCALE && o.src < 1) return fail(ATHK_ERR_ARG, "refine_set_ops: box %d: divisor %d", n, o.src);
      if (o.kind == BOX_COARSE_BC_FC) {
        const int d = o.pad;
        if (d < 0 || d > 2 || o.lo[d] != o.hi[d] || (o.src != 1 && o.src != 2) || o.so[d] < 0 ||
            o.so[d] >= (d == 0 ? cg.cn1 : (d == 1 ? cg.cn2 : cg.cn3)) + (o.dir == d))
          return fail(ATHK_ERR_ARG, "refine_set_ops: box %d is not a one-layer boundary box of the field", n);
      }
    }
    const bool coarse_dst = (o.kind == BOX_RESTRICT || o.kind == BOX_FROM_COARSER || o.kind == BOX_C2P_COARSE ||
                             o.kind == BOX_PROLONG || o.kind == BOX_COARSE_BC || o.kind == BOX_FC_RESTRICT ||
                             o.kind == BOX_FC_FROM_COARSER || o.kind == BOX_FC_PROLONG_SHARED ||
                             o.kind == BOX_FC_PROLONG_INTERNAL || o.kind == BOX_COARSE_BC_FC);
    const bool coarse_face = (o.kind == BOX_FC_RESTRICT || o.kind == BOX_FC_FROM_COARSER || o.kind == BOX_FC_PROLONG_SHARED ||
                              o.kind == BOX_COARSE_BC_FC);
    if (o.kind == BOX_COARSE_BC && (o.dir < 0 || o.dir > 2 || o.lo[o.dir] != o.hi[o.dir] || (o.src != 1 && o.src != 2) ||
                                    o.so[o.dir] < 0 || o.so[o.dir] >= (o.dir == 0 ? cg.cn1 : (o.dir == 1 ? cg.cn2 : cg.cn3))))
      return fail(ATHK_ERR_ARG, "refine_set_ops: box %d is not a one-layer boundary box", n);
    const int lim[3] = {coarse_dst ? cg.cn1 + (coarse_face && o.dir == 0) : g.nc1 + 1,
                        coarse_dst ? cg.cn2 + (coarse_face && o.dir == 1) : g.nc2 + 1,
                        coarse_dst ? cg.cn3 + (coarse_face && o.dir == 2) : g.nc3 + 1};
    long long vol = 1;
    for (int d = 0; d < 3; ++d) {
      // the gradients read one coarse cell around (shared faces: in the two transverse directions)
      const int inset = (o.kind == BOX_PROLONG || (o.kind == BOX_FC_PROLONG_SHARED && d != o.dir)) ? 1 : 0;
      if (o.lo[d] < inset || o.hi[d] >= lim[d] - inset || o.hi[d] < o.lo[d])
        return fail(ATHK_ERR_ARG, "refine_set_ops: box %d (kind %d) leaves its array in direction %d: %d..%d", n,
                    o.kind, d, o.lo[d], o.hi[d]);
      vol *= o.hi[d] - o.lo[d] + 1;
    }
    if (o.kind == BOX_FLUX_FROM_FINER && (o.dir < 0 || o.dir > 2 || o.lo[o.dir] != o.hi[o.dir]))
      return fail(ATHK_ERR_ARG, "refine_set_ops: box %d is not a face box of direction %d", n, o.dir);
    maxvol = std::max(maxvol, vol);
  }
  if (c->rf.ops[phase]) { cudaFree(c->rf.ops[phase]); c->rf.ops[phase] = nullptr; }
  c->rf.nops[phase] = nops;
  c->rf.maxvol[phase] = maxvol;
  if (nops > 0) {
    CU(cudaMalloc((void **)&c->rf.ops[phase], (size_t)nops * sizeof(BoxOp)));
    CU(cudaMemcpyAsync(c->rf.ops[phase], ops, (size_t)nops * sizeof(BoxOp), cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
  }
  return ATHK_OK;
}

#ifdef ATHK_EMULATE
#include "athk_emul_refine_check.h"
#endif
extern "C" int athk_refine_run(athk_ctx *c, int phase, int reg) {
  NEED_READY(c);
  if (!c->rf.on) return fail(ATHK_ERR_STATE, "refine_run: call athk_refine_enable first");
  if (phase < 0 || phase >= ATHK_REFINE_NPHASES) return fail(ATHK_ERR_ARG, "refine_run: phase %d", phase);
  if (bad_reg(reg) || !c->u[reg]) return fail(ATHK_ERR_ARG, "refine_run: register %d not allocated", reg);
  if (c->rf.nops[phase] == 0) return ATHK_OK;
  const bool mhd = c->cfg.mhd != 0;
  const int ns = c->cfg.nscalars;
  if (ns && !c->s[reg]) return fail(ATHK_ERR_ARG, "refine_run: scalar register %d not allocated", reg);
  BoxOpPoint f = {c->g, c->rf.cg, c->geo, c->eos, c->rf.ops[phase], c->u[reg], c->w, c->rf.cu, c->rf.cw,
                  {c->flux[0], c->flux[1], c->flux[2]}, c->rf.xv, c->rf.cxv,
                  ns, ns ? c->s[reg] : nullptr, c->r, c->rf.cs, c->rf.cr, {c->sflux[0], c->sflux[1], c->sflux[2]},
                  mhd ? 1 : 0, mhd ? c->b[reg] : nullptr, c->bcc, c->rf.cb, c->rf.cbcc, c->e, c->emfbuf, c->sf};
  if (mhd && !c->b[reg]) return fail(ATHK_ERR_ARG, "refine_run: field register %d not allocated", reg);
  for (int q = 0; q <= ATHK_MAX_PEERS; ++q) {   // what the boxes read of their source blocks: own arrays, peers' arrays
    const athk_ctx *p = (q == 0) ? c : c->rf.peers[q - 1];
    PeerView &v = f.pv[q];
    if (!p) { v = PeerView{}; continue; }
    v.u = p->u[reg]; v.s = ns ? p->s[reg] : nullptr; v.b = mhd ? p->b[reg] : nullptr;
    v.cu = p->rf.cu; v.cs = p->rf.cs; v.cb = p->rf.cb; v.emfbuf = p->emfbuf;
    for (int d = 0; d < 3; ++d) { v.flx[d] = p->flux[d]; v.sflx[d] = p->sflux[d]; v.dxf[d] = p->geo.dxf[d]; }
  }
#ifdef ATHK_EMULATE
  // test harness only (tests/emul/athk_emul_refine_check.h): with ATHK_REFINE_ORDER_CHECK set, the list is also
  // run box by box in reverse order from the same state; any difference means two boxes of one list
  // disagree about an element, which the concurrent launch on the device would turn into a race
  if (int rcq = athk_emul_refine_order_check(c, phase, reg, f)) return rcq;
#endif
  static const char *serial_env = getenv("ATHK_REFINE_SERIAL");   // debugging: one launch per box of the listed phases ("all" or e.g. "12,13")
  bool serial = false;
  if (serial_env) {
    char key[16];
    snprintf(key, sizeof key, ",%d,", phase);
    std::string sl = std::string(",") + serial_env + ",";
    serial = (sl == ",all,") || sl.find(key) != std::string::npos;
  }
  if (serial) {
    for (int n = 0; n < c->rf.nops[phase]; ++n) {
      f.ops = c->rf.ops[phase] + n;
      LAUNCH_NAMED(c, "athk_refine_run", launch_links(c->stream, 1, c->rf.maxvol[phase], f));
    }
    return ATHK_OK;
  }
  for (int first = 0; first < c->rf.nops[phase]; first += 32768) {   // (grid.y limit)
    f.ops = c->rf.ops[phase] + first;
    const int n = std::min(32768, c->rf.nops[phase] - first);
    LAUNCH_NAMED(c, "athk_refine_run", launch_links(c->stream, n, c->rf.maxvol[phase], f));
  }
  return ATHK_OK;
}

extern "C" int athk_enroll_user_boundary(athk_ctx *c, int face, athk_bval_fn fn, void *user) {
  if (!c) return fail(ATHK_ERR_ARG, "enroll_user_boundary: null context");
  if (face < 0 || face > 5) return fail(ATHK_ERR_ARG, "enroll_user_boundary: face %d (0..5)", face);
  c->user_bc[face] = fn;
  c->user_bc_data[face] = user;
  return ATHK_OK;
}

extern "C" int athk_set_boundary_time(athk_ctx *c, double time, double dt) {
  if (!c) return fail(ATHK_ERR_ARG, "set_boundary_time: null context");
  c->bc_time = time;
  c->bc_dt = dt;
  return ATHK_OK;
}

extern "C" int athk_apply_physical_boundaries(athk_ctx *c) {
  NEED_READY(c);
  if (c->bc_dirty) {
    for (int f = 0; f < 6; ++f) c->bc_any[f] = 0;
    for (int b = 0; b < c->g.nb; ++b)
      for (int f = 0; f < 6; ++f)
        if (c->h_bcflags[(size_t)b * 6 + f] != ATHK_BC_NONE) {
          if (c->h_nbr[(size_t)b * 27 + nbr_slot(f / 2 == 0 ? (f % 2 ? 1 : -1) : 0, f / 2 == 1 ? (f % 2 ? 1 : -1) : 0,
                                                 f / 2 == 2 ? (f % 2 ? 1 : -1) : 0)] != -1)
            return fail(ATHK_ERR_STATE, "apply_physical_boundaries: block %d face %d is flagged as a physical "
                                        "boundary but has a neighbour in the table", b, f);
          c->bc_any[f] = 1;
        }
    if (!c->bcflags) CU(cudaMalloc((void **)&c->bcflags, c->h_bcflags.size() * sizeof(int)));
    CU(cudaMemcpyAsync(c->bcflags, c->h_bcflags.data(), c->h_bcflags.size() * sizeof(int),
                       cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    c->bc_dirty = false;
  }
  int rc = ATHK_OK;
  for (int side = 0; side < 2 && !rc; ++side) if (c->bc_any[0 + side]) rc = launch_bc_face<0>(c, side);
  for (int side = 0; side < 2 && !rc; ++side) if (c->bc_any[2 + side]) rc = launch_bc_face<1>(c, side);
  for (int side = 0; side < 2 && !rc; ++side) if (c->bc_any[4 + side]) rc = launch_bc_face<2>(c, side);
  return rc;
}

extern "C" int athk_set_block_floors(athk_ctx *c, int lb, double dfloor, double pfloor, double sfloor) {
  if (!c) return fail(ATHK_ERR_ARG, "set_block_floors: null context");
  cudaSetDevice(c->cfg.device);
  const int nb = c->g.nb;
  if (lb < 0 || lb >= nb) return fail(ATHK_ERR_ARG, "set_block_floors: block %d out of range", lb);
  if (!(dfloor > 0.0) || !(pfloor > 0.0) || !(sfloor >= 0.0))
    return fail(ATHK_ERR_ARG, "set_block_floors: floors must be positive (got %g, %g, %g)", dfloor, pfloor, sfloor);
  if (c->h_floors.empty()) {   // first call: every block starts from the context's floors
    c->h_floors.resize((size_t)nb * 3);
    for (int b = 0; b < nb; ++b) {
      c->h_floors[3 * b] = c->eos.dfloor;
      c->h_floors[3 * b + 1] = c->eos.pfloor;
      c->h_floors[3 * b + 2] = c->eos.sfloor;
    }
    CU(cudaMalloc((void **)&c->d_floors, c->h_floors.size() * sizeof(Real)));
  }
  c->h_floors[3 * lb] = dfloor;
  c->h_floors[3 * lb + 1] = pfloor;
  c->h_floors[3 * lb + 2] = sfloor;
  CU(cudaStreamSynchronize(c->stream));   // no kernel may be reading the table
  CU(cudaMemcpyAsync(c->d_floors, c->h_floors.data(), c->h_floors.size() * sizeof(Real), cudaMemcpyHostToDevice,
                     c->stream));
  CU(cudaStreamSynchronize(c->stream));
  c->eos.blk_floors = c->d_floors;
  return ATHK_OK;
}

extern "C" int athk_set_constant_acceleration(athk_ctx *c, double g1, double g2, double g3) {
  if (!c) return fail(ATHK_ERR_ARG, "set_constant_acceleration: null context");
  if (c->eos.sr && (g1 != 0.0 || g2 != 0.0 || g3 != 0.0))
    return fail(ATHK_ERR_ARG, "set_constant_acceleration: no source terms with special relativity");
  c->gacc[0] = g1; c->gacc[1] = g2; c->gacc[2] = g3;
  return ATHK_OK;
}

extern "C" int athk_enroll_user_source(athk_ctx *c, athk_src_fn fn, void *user) {
  if (!c) return fail(ATHK_ERR_ARG, "enroll_user_source: null context");
  if (c->eos.sr && fn) return fail(ATHK_ERR_ARG, "enroll_user_source: no source terms with special relativity");
  c->user_src = fn;
  c->user_src_data = user;
  return ATHK_OK;
}

extern "C" int athk_set_source_time(athk_ctx *c, double time) {
  if (!c) return fail(ATHK_ERR_ARG, "set_source_time: null context");
  c->src_time = time;
  return ATHK_OK;
}

extern "C" int athk_add_source_terms(athk_ctx *c, double beta_dt) {
  NEED_READY(c);
  const Grid &g = c->g;
  if (!(c->gacc[0] == 0.0 && c->gacc[1] == 0.0 && c->gacc[2] == 0.0)) {
    Range3 r = {g.ks, g.ke, g.js, g.je, g.is, g.ie};
    ConstAccPoint f = {g, c->w, c->u[0], beta_dt, c->gacc[0], c->gacc[1], c->gacc[2], c->eos.iso};
    LAUNCH_NAMED(c, "athk_add_source_terms", launch_points_occ<ATHK_STREAM_MINB>(c->stream, g.nb, r, f));
  }
  if (c->user_src) {   // user-defined source terms come last (hydro_srcterms.cpp:151-154)
    const int ns = c->cfg.nscalars;
    for (int lb = 0; lb < g.nb; ++lb) {
      athk_src_args a;
      a.lb = lb; a.time = c->src_time; a.dt = beta_dt;
      a.is = g.is; a.ie = g.ie; a.js = g.js; a.je = g.je; a.ks = g.ks; a.ke = g.ke;
      a.w = c->w + (idx_t)lb * 5 * g.sn;
      a.bcc = c->cfg.mhd ? c->bcc + (idx_t)lb * 3 * g.sn : nullptr;
      a.r = ns ? c->r + (idx_t)lb * ns * g.sn : nullptr;
      a.u = c->u[0] + (idx_t)lb * 5 * g.sn;
      a.s = ns ? c->s[0] + (idx_t)lb * ns * g.sn : nullptr;
      a.stride_j = g.sj; a.stride_k = g.sk; a.stride_n = g.sn;
      a.nscalars = ns;
      a.stream = (void *)c->stream;
      if (c->user_src(c->user_src_data, &a) != 0)
        return fail(ATHK_ERR_STATE, "user source function failed on block %d", lb);
    }
  }
  return ATHK_OK;
}

// History sums (outputs/history.cpp:196-248).  The reference adds cells serially; here every
// MeshBlock is reduced by one CTA in a fixed order (deterministic, not the serial order: the sums
// agree with the reference to rounding, ~1e-15 relative; the north-star bar is 1e-10), and the
// host adds the per-block partial sums block by block.
struct HistoryCell {
  Grid g;
  Geom geo;
  const Real *u, *bcc, *s;
  int mhd, ns;
  ATHK_HD int nvals() const { return 8 + (mhd ? 3 : 0) + ns; }
  ATHK_HD void operator()(int blk, int k, int j, int i, Real *acc) const {
    const idx_t c = gidx(g, k, j, i);
    const Real vol = geo.dxf[0][blk * geo.PG + i] * geo.dxf[1][blk * geo.PG + j] * geo.dxf[2][blk * geo.PG + k];
    const Real *ub = u + (idx_t)blk * 5 * g.sn + c;
    const Real u_d = ub[0], u_mx = ub[g.sn], u_my = ub[2 * g.sn], u_mz = ub[3 * g.sn];
    acc[0] += vol * u_d;
    acc[1] += vol * u_mx;
    acc[2] += vol * u_my;
    acc[3] += vol * u_mz;
    acc[4] += vol * 0.5 * sqr(u_mx) / u_d;
    acc[5] += vol * 0.5 * sqr(u_my) / u_d;
    acc[6] += vol * 0.5 * sqr(u_mz) / u_d;
    acc[7] += vol * ub[4 * g.sn];
    int n0 = 8;
    if (mhd) {
      const Real *bc = bcc + (idx_t)blk * 3 * g.sn + c;
      acc[8] += vol * 0.5 * bc[0] * bc[0];
      acc[9] += vol * 0.5 * bc[g.sn] * bc[g.sn];
      acc[10] += vol * 0.5 * bc[2 * g.sn] * bc[2 * g.sn];
      n0 = 11;
    }
    for (int n = 0; n < ns; ++n) acc[n0 + n] += vol * s[((idx_t)blk * ns + n) * g.sn + c];
  }
};
#define ATHK_HIST_MAX 19   // 8 + 3 + nscalars(<=8)

#ifndef ATHK_EMULATE
__global__ void __launch_bounds__(256) k_history(const HistoryCell f, Real *out) {
  const Grid &g = f.g;
  const int blk = blockIdx.x;
  const int nv = f.nvals();
  const int ncell = g.nx1 * g.nx2 * g.nx3;
  Real acc[ATHK_HIST_MAX];
#pragma unroll
  for (int v = 0; v < ATHK_HIST_MAX; ++v) acc[v] = 0.0;
  for (int q = threadIdx.x; q < ncell; q += blockDim.x) {
    const int i = q % g.nx1;
    const int t = q / g.nx1;
    f(blk, g.ks + t / g.nx2, g.js + t % g.nx2, g.is + i, acc);
  }
  __shared__ Real sm[8][ATHK_HIST_MAX];
#pragma unroll
  for (int v = 0; v < ATHK_HIST_MAX; ++v) {
    Real m = acc[v];
    for (int o = 16; o > 0; o >>= 1) m += __shfl_xor_sync(0xffffffffu, m, o);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5][v] = m;
  }
  __syncthreads();
  if (threadIdx.x < nv) {
    Real m = 0.0;
    for (int w = 0; w < 8; ++w) m += sm[w][threadIdx.x];
    out[blk * ATHK_HIST_MAX + threadIdx.x] = m;
  }
}
#endif

extern "C" int athk_history(athk_ctx *c, double *sums, int nsums) {
  NEED_READY(c);
  const Grid &g = c->g;
  HistoryCell f = {g, c->geo, c->u[0], c->bcc, c->s[0], c->cfg.mhd, c->cfg.nscalars};
  const int nv = f.nvals();
  if (!sums || nsums < nv) return fail(ATHK_ERR_ARG, "history: need room for %d sums (got %d)", nv, nsums);
  std::vector<Real> h((size_t)g.nb * ATHK_HIST_MAX, 0.0);
#ifndef ATHK_EMULATE
  Real *d = nullptr;
  CU(cudaMalloc((void **)&d, h.size() * sizeof(Real)));
  prof_begin(c, "history");
  k_history<<<g.nb, 256, 0, c->stream>>>(f, d);
  prof_end(c);
  c->launches++;
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpyAsync(h.data(), d, h.size() * sizeof(Real), cudaMemcpyDeviceToHost, c->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  cudaFree(d);
  if (e != cudaSuccess) return fail(ATHK_ERR_CUDA, "history: %s", cudaGetErrorString(e));
#else
  for (int b = 0; b < g.nb; ++b)
    for (int k = g.ks; k <= g.ke; ++k)
      for (int j = g.js; j <= g.je; ++j)
        for (int i = g.is; i <= g.ie; ++i) f(b, k, j, i, &h[(size_t)b * ATHK_HIST_MAX]);
#endif
  for (int v = 0; v < nv; ++v) sums[v] = 0.0;
  for (int b = 0; b < g.nb; ++b)
    for (int v = 0; v < nv; ++v) sums[v] += h[(size_t)b * ATHK_HIST_MAX + v];
  return ATHK_OK;
}
//----------------------------------------------------------------------------------------
// Fused stage: same task sequence, fewer passes over HBM (see athk_sweep.h, athk_sweep_launch_impl.h).
extern "C" int athk_stage(athk_ctx *c, int stage, int order, const athk_stage_wghts *sw,
                          double dt_mesh) {
  int rc = athk_stage_phase(c, 0, stage, order, sw, dt_mesh);
  if (!rc) rc = athk_stage_phase(c, 1, stage, order, sw, dt_mesh);
  if (!rc) rc = athk_stage_phase(c, 2, stage, order, sw, dt_mesh);
  return rc;
}

extern "C" int athk_stage_phase(athk_ctx *c, int phase, int stage, int order,
                                const athk_stage_wghts *sw, double dt_mesh) {
  NEED_READY(c);
  const bool split = phase >= 10;   // 10,13,11,14,12: the same stage with two extra overlap windows
  if (split) phase -= 10;
  if (phase < 0 || phase > (split ? 4 : 2))
    return fail(ATHK_ERR_ARG, "stage_phase: phase must be 0, 1, 2 (or 10, 13, 11, 14, 12)");
  if (!sw) return fail(ATHK_ERR_ARG, "stage: null weights");
  if (sw->gamma_3 != 0.0) return fail(ATHK_ERR_ARG, "stage: gamma_3 != 0 (3-register integrators) not implemented");
  { int rco = check_order(c, order, "stage"); if (rco) return rco; }
  const int ns = c->cfg.nscalars;
  const Grid &g = c->g;
  const bool mhd = c->cfg.mhd != 0;
  const int swap = (sw->gamma_1 == 0.0 && sw->gamma_2 == 1.0 && sw->gamma_3 == 0.0) ? 1 : 0;
  int wa = -2;
  if (!swap) {
    const double w[5] = {sw->gamma_1, sw->gamma_2, 0.0, 0.0, 0.0};
    wa_mode(w, &wa);
  }
  int rc;
  if (phase == 0) {
    SweepArgs a = make_sweep_args(c, dt_mesh, 0, false);
    HydroIntegrate &h = a.hi;
    h.u = c->u[0]; h.u1 = c->u[1];
    h.u_out = swap ? c->u[1] : c->u[0];
    h.u1_out = c->u[1];
    h.delta = sw->delta; h.g1 = sw->gamma_1; h.g2 = sw->gamma_2;
    h.wght = sw->beta * dt_mesh;
    h.swap = swap; h.u1_zero = (stage == 1) ? 1 : 0; h.wa_mode = wa;
    if (ns) {   // the sweeps leave each face's mass flux in component 0 of s_flux[d]
      for (int d = 0; d < 3; ++d) a.mf[d] = c->sflux[d];
      a.mf_bstride = (idx_t)ns * g.sn;
    }
    if (c->eos.sr) {
      rc = athk_sweeps_rs<RS_SR, false, 0>(c, order, a);
    } else if (c->eos.iso && c->cfg.rsolver == (mhd ? ATHK_RS_HLLD : ATHK_RS_HLLE)) {
      rc = mhd ? athk_sweeps_rs<RS_HLLD_ISO, true, 0>(c, order, a) : athk_sweeps_rs<RS_HLLE_ISO, false, 0>(c, order, a);
    } else
    switch (c->cfg.rsolver) {
      case ATHK_RS_ROE: rc = mhd ? athk_sweeps_rs<RS_ROE, true, 0>(c, order, a) : athk_sweeps_rs<RS_ROE, false, 0>(c, order, a); break;
      case ATHK_RS_HLLE: rc = mhd ? athk_sweeps_rs<RS_HLLE, true, 0>(c, order, a) : athk_sweeps_rs<RS_HLLE, false, 0>(c, order, a); break;
      case ATHK_RS_LLF: rc = mhd ? athk_sweeps_rs<RS_LLF, true, 0>(c, order, a) : athk_sweeps_rs<RS_LLF, false, 0>(c, order, a); break;
      default:
#ifndef ATHK_EMULATE
        if (c->cfg.arith == ATHK_ARITH_FAST) {   // contracted, approximately rounded kernel set (athk_fast.cu)
          rc = mhd ? athk_sweeps_rs<RS_HLLD, true, 1>(c, order, a) : athk_sweeps_rs<RS_HLLC, false, 1>(c, order, a);
          break;
        }
#endif
        rc = mhd ? athk_sweeps_rs<RS_HLLD, true, 0>(c, order, a) : athk_sweeps_rs<RS_HLLC, false, 0>(c, order, a);
        break;
    }
    if (rc) return rc;
    if (swap) { Real *t = c->u[0]; c->u[0] = c->u[1]; c->u[1] = t; }
    if ((rc = athk_add_source_terms(c, sw->beta * dt_mesh))) return rc;   // SRC_TERM follows INT_HYD
    if (ns) {   // IntegrateScalars with the same weights (time_integrator.cpp:2376-2403)
      HydroIntegrate hs = h;
      hs.u = c->s[0]; hs.u1 = c->s[1];
      hs.u_out = swap ? c->s[1] : c->s[0];
      hs.u1_out = c->s[1];
      rc = (order == 1) ? launch_scalar_integrate<1>(c, hs)
                        : (order == 2 ? launch_scalar_integrate<2>(c, hs) : launch_scalar_integrate<3>(c, hs));
      if (rc) return rc;
      if (swap) { Real *t = c->s[0]; c->s[0] = c->s[1]; c->s[1] = t; }
    }
    if (!split && (rc = ghost_pack_early(c, 0))) return rc;   // u (and s) are final: send them now
    if (mhd) {
      if ((rc = corner_e_part(c, split ? 1 : 0))) return rc;
      if ((rc = athk_send_flux_correction(c))) return rc;
      if ((rc = athk_remote_pack_emf(c))) return rc;
    }
    return ATHK_OK;
  }
  if (phase == 3) return mhd ? corner_e_part(c, 2) : ATHK_OK;   // interior edges, under the EMF transfer
  if (phase == 4) {   // under the ghost transfer: ghosts from same-GPU neighbours (they must see the
                      // neighbour's U before W(U) floors it, as in the reference), then the active zone
    if ((rc = exchange_ghosts_mode(c, (mhd ? 3 : 1) | (ns ? 4 : 0), 0, 1))) return rc;
    return cons2prim_part(c, 0, 1);
  }
  if (phase == 1) {
    if (mhd) {
      if ((rc = athk_recv_flux_correction(c))) return rc;
      FieldIntegrate fi;
      fi.b = c->b[0]; fi.b1 = c->b[1];
      fi.b_out = swap ? c->b[1] : c->b[0];
      fi.b1_out = c->b[1];
      fi.delta = sw->delta; fi.g1 = sw->gamma_1; fi.g2 = sw->gamma_2;
      fi.wght = sw->beta * dt_mesh;
      fi.swap = swap; fi.b1_zero = (stage == 1) ? 1 : 0; fi.wa_mode = wa;
      Range3 r = {g.ks, g.ke + 1, g.js, g.je + 1, g.is, g.ie + 1};
      CTIntegratePoint f = {g, c->geo, c->e, fi};
      LAUNCH_NAMED(c, "ct_integrate", launch_points_kcol<ATHK_CT_MINB>(c->stream, g.nb, r, c->kt_ct, f));
      if (swap) { Real *t = c->b[0]; c->b[0] = c->b[1]; c->b[1] = t; }
    }
    return ghost_pack_finish(c, 0);
  }
  if ((rc = exchange_ghosts_mode(c, (mhd ? 3 : 1) | (ns ? 4 : 0), 0, split ? 2 : 0))) return rc;
  if ((rc = cons2prim_part(c, 0, split ? 2 : 0))) return rc;
  return athk_apply_physical_boundaries(c);   // PHY_BVAL (no-op without physical boundaries)
}
//----------------------------------------------------------------------------------------
// Remote links (neighbours that live on another GPU).  The transport itself (NCCL send/recv or
// peer copies over NVLink) is done by the host framework between pack and unpack; this library
// packs messages into / reads them from device buffers the caller owns.
extern "C" long long athk_remote_ghost_count(athk_ctx *c, int ox1, int ox2, int ox3) {
  if (!c) return -1;
  return ghost_section_off(c->g, c->cfg.mhd != 0, c->cfg.nscalars, ox1, ox2, ox3, 5);
}

extern "C" long long athk_remote_emf_count(athk_ctx *c, int ox1, int ox2, int ox3) {
  if (!c || !c->cfg.mhd) return 0;
  int dir, od;
  if (!emf_link_section(ox1, ox2, ox3, &dir, &od)) return 0;
  return dir == 0 ? c->sf.n_x1() : (dir == 1 ? c->sf.n_x2() : c->sf.n_x3());
}

extern "C" int athk_remote_configure(athk_ctx *c, int nlinks, const int *lb, const int *ox1,
                                     const int *ox2, const int *ox3,
                                     const long long *ghost_send_off,
                                     const long long *ghost_recv_off,
                                     const long long *emf_send_off, const long long *emf_recv_off) {
  if (!c) return fail(ATHK_ERR_ARG, "null context");
  cudaSetDevice(c->cfg.device);
  if (nlinks < 0) return fail(ATHK_ERR_ARG, "remote_configure: nlinks < 0");
  cudaFree(c->rl_int); cudaFree(c->rl_off);
  c->rl_int = nullptr; c->rl_off = nullptr;
  memset(&c->rl, 0, sizeof(c->rl));
  c->rl.tab = c->slab_tab;
  c->max_ghost_count = c->max_emf_count = 0;
  if (nlinks == 0) return ATHK_OK;
  if (!lb || !ox1 || !ox2 || !ox3 || !ghost_send_off || !ghost_recv_off)
    return fail(ATHK_ERR_ARG, "remote_configure: null array");
  const size_t n = nlinks;
  c->h_rl_int.assign(4 * n, 0);
  c->h_rl_off.assign(5 * n, 0);
  for (size_t q = 0; q < n; ++q) {
    if (lb[q] < 0 || lb[q] >= c->g.nb) return fail(ATHK_ERR_ARG, "remote_configure: block out of range");
    c->h_rl_int[q] = lb[q]; c->h_rl_int[n + q] = ox1[q];
    c->h_rl_int[2 * n + q] = ox2[q]; c->h_rl_int[3 * n + q] = ox3[q];
    c->h_rl_off[q] = ghost_send_off[q]; c->h_rl_off[n + q] = ghost_recv_off[q];
    c->h_rl_off[2 * n + q] = emf_send_off ? emf_send_off[q] : 0;
    c->h_rl_off[3 * n + q] = emf_recv_off ? emf_recv_off[q] : 0;
    idx_t ebase = 0;
    int dir, od;
    if (c->cfg.mhd && emf_link_section(ox1[q], ox2[q], ox3[q], &dir, &od))
      // the neighbour's side that faces this block: neighbour at -1 -> its outer side
      ebase = c->h_rl_off[3 * n + q] - c->sf.face_off(dir, od < 0 ? 1 : 0);
    c->h_rl_off[4 * n + q] = ebase;
    const idx_t gc = athk_remote_ghost_count(c, ox1[q], ox2[q], ox3[q]);
    const idx_t ec = athk_remote_emf_count(c, ox1[q], ox2[q], ox3[q]);
    if (gc > c->max_ghost_count) c->max_ghost_count = gc;
    if (ec > c->max_emf_count) c->max_emf_count = ec;
  }
  int rc = dev_alloc(&c->rl_int, 4 * n, c->stream);
  if (!rc) rc = dev_alloc(&c->rl_off, 5 * n, c->stream);
  if (rc) return rc;
  CU(cudaMemcpyAsync(c->rl_int, c->h_rl_int.data(), 4 * n * sizeof(int), cudaMemcpyHostToDevice, c->stream));
  CU(cudaMemcpyAsync(c->rl_off, c->h_rl_off.data(), 5 * n * sizeof(idx_t), cudaMemcpyHostToDevice, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  c->rl.n = nlinks;
  c->rl.lb = c->rl_int; c->rl.o1 = c->rl_int + n; c->rl.o2 = c->rl_int + 2 * n; c->rl.o3 = c->rl_int + 3 * n;
  c->rl.gsend = c->rl_off; c->rl.grecv = c->rl_off + n; c->rl.esend = c->rl_off + 2 * n;
  c->rl.erecv = c->rl_off + 3 * n; c->rl.ebase = c->rl_off + 4 * n;
  return ATHK_OK;
}

extern "C" int athk_remote_set_buffers(athk_ctx *c, double *send_ghost, double *recv_ghost,
                                       double *send_emf, double *recv_emf) {
  if (!c) return fail(ATHK_ERR_ARG, "null context");
  c->rl.gsbuf = send_ghost; c->rl.grbuf = recv_ghost;
  c->rl.esbuf = send_emf; c->rl.erbuf = recv_emf;
  return ATHK_OK;
}

extern "C" int athk_remote_pack_ghosts(athk_ctx *c, int reg) {
  NEED_READY(c);
  if (c->rl.n == 0) return ATHK_OK;
  if (bad_reg(reg) || !c->u[reg]) return fail(ATHK_ERR_ARG, "remote_pack_ghosts: register %d not allocated", reg);
  if (!c->rl.gsbuf && !c->p2p.on) return fail(ATHK_ERR_STATE, "remote_pack_ghosts: call athk_remote_set_buffers first");
  RemoteLinks rl = c->rl;
  if (c->p2p.on) {   // message number gseq+1 goes straight into the receivers' buffer half (gseq+1)&1
    c->p2p.gseq++;
    rl.gdst = c->p2p.gdst + (c->p2p.gseq & 1) * (size_t)c->rl.n;
  }
  RemoteGhostPack f = {c->g, c->cfg.mhd, c->cfg.nscalars, rl, c->u[reg],
                       c->cfg.mhd ? c->b[reg] : nullptr, c->cfg.nscalars ? c->s[reg] : nullptr, 0};
  LAUNCH_NAMED(c, __func__, launch_links(c->stream, c->rl.n, c->max_ghost_count, f));
  if (c->p2p.on) return athk_p2p_after_pack(c, 0);
  return ATHK_OK;
}

// Peer-to-peer transport only: the cell-centred sections of the next ghost message, packed on the
// side stream as soon as the registers are final (after the sweeps), while the main stream goes on
// with corner E / EMF correction / CT.  ghost_pack_finish adds the face-field sections and raises
// the flags.  (Message n+1 lands in the buffer half the receiver read message n-1 from; it has
// finished with it: it sent message n after that read, and this GPU waited for message n before
// the sweeps that precede this pack.)
static int ghost_pack_early(athk_ctx *c, int reg) {
#ifndef ATHK_EMULATE
  if (!c->p2p.on || c->rl.n == 0 || !c->p2p.side) return ATHK_OK;
  c->p2p.gseq++;
  RemoteLinks rl = c->rl;
  rl.gdst = c->p2p.gdst + (c->p2p.gseq & 1) * (size_t)c->rl.n;
  RemoteGhostPack f = {c->g, c->cfg.mhd, c->cfg.nscalars, rl, c->u[reg],
                       c->cfg.mhd ? c->b[reg] : nullptr, c->cfg.nscalars ? c->s[reg] : nullptr, 1};
  CU(cudaEventRecord(c->p2p.ev_ready, c->stream));
  CU(cudaStreamWaitEvent(c->p2p.side, c->p2p.ev_ready, 0));
  CU(launch_links(c->p2p.side, c->rl.n, c->max_ghost_count, f));
  c->launches++;
  CU(cudaEventRecord(c->p2p.ev_packed, c->p2p.side));
  c->p2p.early_pending = true;
#endif
  return ATHK_OK;
}

static int ghost_pack_finish(athk_ctx *c, int reg) {
#ifndef ATHK_EMULATE
  if (c->p2p.on && c->p2p.early_pending) {
    c->p2p.early_pending = false;
    if (c->cfg.mhd) {
      RemoteLinks rl = c->rl;
      rl.gdst = c->p2p.gdst + (c->p2p.gseq & 1) * (size_t)c->rl.n;
      RemoteGhostPack f = {c->g, c->cfg.mhd, c->cfg.nscalars, rl, c->u[reg], c->b[reg],
                           c->cfg.nscalars ? c->s[reg] : nullptr, 2};
      LAUNCH_NAMED(c, "athk_remote_pack_ghosts", launch_links(c->stream, c->rl.n, c->max_ghost_count, f));
    }
    CU(cudaStreamWaitEvent(c->stream, c->p2p.ev_packed, 0));
    return athk_p2p_after_pack(c, 0);
  }
#endif
  return athk_remote_pack_ghosts(c, reg);
}

extern "C" int athk_remote_pack_emf(athk_ctx *c) {
  NEED_READY(c);
  if (c->rl.n == 0 || !c->cfg.mhd) return ATHK_OK;
  if (!c->rl.esbuf && !c->p2p.on) return fail(ATHK_ERR_STATE, "remote_pack_emf: call athk_remote_set_buffers first");
  RemoteLinks rl = c->rl;
  if (c->p2p.on) {
    c->p2p.eseq++;
    rl.edst = c->p2p.edst + (c->p2p.eseq & 1) * (size_t)c->rl.n;
  }
  RemoteEmfPack f = {c->sf, rl, c->emfbuf};
  LAUNCH_NAMED(c, __func__, launch_links(c->stream, c->rl.n, c->max_emf_count, f));
  if (c->p2p.on) return athk_p2p_after_pack(c, 1);
  return ATHK_OK;
}

//----------------------------------------------------------------------------------------
// Peer-to-peer transport (include/athk.h): allocation / mapping helpers, flag kernels
extern "C" int athk_p2p_alloc(athk_ctx *c, long long nbytes, void **dptr, unsigned char ipc_handle[64]) {
  if (!c || !dptr || nbytes <= 0) return fail(ATHK_ERR_ARG, "p2p_alloc: bad argument");
#ifndef ATHK_EMULATE
  CU(cudaSetDevice(c->cfg.device));
  CU(cudaMalloc(dptr, (size_t)nbytes));
  CU(cudaMemset(*dptr, 0, (size_t)nbytes));
  if (ipc_handle) {
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, *dptr));
    memcpy(ipc_handle, &h, 64);
  }
  return ATHK_OK;
#else
  *dptr = calloc(1, (size_t)nbytes);
  if (ipc_handle) memset(ipc_handle, 0, 64);
  return *dptr ? ATHK_OK : fail(ATHK_ERR_CUDA, "p2p_alloc: out of memory");
#endif
}

extern "C" int athk_p2p_open(athk_ctx *c, const unsigned char ipc_handle[64], void **dptr) {
  if (!c || !dptr || !ipc_handle) return fail(ATHK_ERR_ARG, "p2p_open: bad argument");
#ifndef ATHK_EMULATE
  CU(cudaSetDevice(c->cfg.device));
  cudaIpcMemHandle_t h;
  memcpy(&h, ipc_handle, 64);
  CU(cudaIpcOpenMemHandle(dptr, h, cudaIpcMemLazyEnablePeerAccess));
  return ATHK_OK;
#else
  return fail(ATHK_ERR_STATE, "p2p_open: no inter-process mapping in the host harness");
#endif
}

extern "C" int athk_p2p_close(athk_ctx *c, void *dptr) {
  if (!c) return fail(ATHK_ERR_ARG, "null context");
#ifndef ATHK_EMULATE
  if (dptr) CU(cudaIpcCloseMemHandle(dptr));
#endif
  return ATHK_OK;
}

extern "C" int athk_p2p_free(athk_ctx *c, void *dptr) {
  if (!c) return fail(ATHK_ERR_ARG, "null context");
#ifndef ATHK_EMULATE
  cudaSetDevice(c->cfg.device);
  CU(cudaStreamSynchronize(c->stream));
  if (dptr) CU(cudaFree(dptr));
#else
  free(dptr);
#endif
  return ATHK_OK;
}

extern "C" int athk_p2p_enable_peer(athk_ctx *c, int peer_device) {
  if (!c) return fail(ATHK_ERR_ARG, "null context");
#ifndef ATHK_EMULATE
  if (peer_device == c->cfg.device) return ATHK_OK;
  CU(cudaSetDevice(c->cfg.device));
  int can = 0;
  CU(cudaDeviceCanAccessPeer(&can, c->cfg.device, peer_device));
  if (!can) return fail(ATHK_ERR_STATE, "p2p_enable_peer: device %d cannot access device %d", c->cfg.device, peer_device);
  cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
  if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); e = cudaSuccess; }
  CU(e);
#endif
  return ATHK_OK;
}

extern "C" long long athk_p2p_timeouts(athk_ctx *c) {
  if (!c || !c->p2p.timeouts) return 0;
#ifndef ATHK_EMULATE
  cudaSetDevice(c->cfg.device);
  unsigned long long v = 0;
  cudaStreamSynchronize(c->stream);
  cudaMemcpy(&v, c->p2p.timeouts, sizeof(v), cudaMemcpyDeviceToHost);
  return (long long)v;
#else
  return 0;
#endif
}

extern "C" int athk_remote_set_p2p(athk_ctx *c, int npeers, double *const *ghost_dst, double *const *emf_dst,
                                   double *recv_ghost, long long ghost_half, double *recv_emf, long long emf_half,
                                   unsigned long long *const *peer_flags, unsigned long long *my_flags) {
  if (!c) return fail(ATHK_ERR_ARG, "null context");
  cudaSetDevice(c->cfg.device);
  const int n = c->rl.n;
  if (n == 0 || npeers <= 0) { c->p2p.on = false; return ATHK_OK; }
  if (!ghost_dst || !recv_ghost || !peer_flags || !my_flags)
    return fail(ATHK_ERR_ARG, "remote_set_p2p: null argument");
  if (c->cfg.mhd && (!emf_dst || !recv_emf)) return fail(ATHK_ERR_ARG, "remote_set_p2p: EMF buffers missing");
  athk_ctx::P2P &p = c->p2p;
  cudaFree(p.gdst); cudaFree(p.edst); cudaFree(p.peer_flags); cudaFree(p.timeouts);
  p.gdst = p.edst = nullptr; p.peer_flags = nullptr; p.timeouts = nullptr;
  CU(cudaMalloc((void **)&p.gdst, 2 * (size_t)n * sizeof(Real *)));
  CU(cudaMemcpy(p.gdst, ghost_dst, 2 * (size_t)n * sizeof(Real *), cudaMemcpyHostToDevice));
  if (emf_dst) {
    CU(cudaMalloc((void **)&p.edst, 2 * (size_t)n * sizeof(Real *)));
    CU(cudaMemcpy(p.edst, emf_dst, 2 * (size_t)n * sizeof(Real *), cudaMemcpyHostToDevice));
  }
  CU(cudaMalloc((void **)&p.peer_flags, (size_t)npeers * sizeof(unsigned long long *)));
  CU(cudaMemcpy(p.peer_flags, peer_flags, (size_t)npeers * sizeof(unsigned long long *), cudaMemcpyHostToDevice));
  CU(cudaMalloc((void **)&p.timeouts, sizeof(unsigned long long)));
  {
    const unsigned long long zero = 0;
    CU(cudaMemcpy(p.timeouts, &zero, sizeof(zero), cudaMemcpyHostToDevice));
  }
  p.npeers = npeers;
  p.grecv = recv_ghost; p.ghalf = ghost_half;
  p.erecv = recv_emf; p.ehalf = emf_half;
  p.my_flags = my_flags;
  p.gseq = p.eseq = 0;
  p.on = true;
#ifndef ATHK_EMULATE
  // side stream of the early ghost pack (ghost_pack_early).  Opt-in (ATHK_EARLY_PACK=1): measured at 8
  // GPUs it LOSES 0.2 ms per cycle (9.97 vs 9.76 ms) -- the NVLink-bound pack kernel holds SM slots
  // while corner E / the EMF kernels run, and they slow down by more than the pack hides
  // (profiles/r2_bench_n8_early_pack_variant.json)
  if (!p.side && getenv("ATHK_EARLY_PACK")) {
    CU(cudaStreamCreateWithFlags(&p.side, cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&p.ev_ready, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&p.ev_packed, cudaEventDisableTiming));
  }
#endif
  return ATHK_OK;
}

#ifndef ATHK_EMULATE
// raise this rank's word `which` (0 ghost, 1 emf) in every peer's flag array
__global__ void k_p2p_signal(unsigned long long *const *peer_flags, int npeers, int which, unsigned long long seq) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npeers) return;
  __threadfence_system();   // the pack kernel's peer stores are ordered before the flag
  *(volatile unsigned long long *)(peer_flags[p] + which) = seq;
}
// wait until every peer's word `which` in this rank's flag array has reached seq
__global__ void k_p2p_wait(unsigned long long *my_flags, int npeers, int which, unsigned long long seq,
                           unsigned long long *timeouts) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npeers) return;
  const volatile unsigned long long *f = my_flags + 2 * p + which;
  const long long t0 = clock64();
  while (*f < seq) {
    if (clock64() - t0 > 20000000000LL) {   // ~10 s: the peer is gone; report instead of hanging
      atomicAdd(timeouts, 1ULL);
      break;
    }
    __nanosleep(200);
  }
  __threadfence_system();
}
#endif

// after the pack kernel of message number seq: tell the peers
static int p2p_signal(athk_ctx *c, int which, unsigned long long seq) {
#ifndef ATHK_EMULATE
  prof_begin(c, "p2p_signal");
  k_p2p_signal<<<(c->p2p.npeers + 31) / 32, 32, 0, c->stream>>>(c->p2p.peer_flags, c->p2p.npeers, which, seq);
  prof_end(c);
  CU(cudaGetLastError());
  c->launches++;
#endif
  return ATHK_OK;
}
static int p2p_wait(athk_ctx *c, int which, unsigned long long seq) {
#ifndef ATHK_EMULATE
  prof_begin(c, "p2p_wait");
  k_p2p_wait<<<(c->p2p.npeers + 31) / 32, 32, 0, c->stream>>>(c->p2p.my_flags, c->p2p.npeers, which, seq, c->p2p.timeouts);
  prof_end(c);
  CU(cudaGetLastError());
  c->launches++;
#endif
  return ATHK_OK;
}
int athk_p2p_before_ghost_fill(athk_ctx *c) {
  if (!c->p2p.on || c->rl.n == 0) return ATHK_OK;
  c->rl.grbuf = c->p2p.grecv + (c->p2p.gseq & 1) * c->p2p.ghalf;
  return p2p_wait(c, 0, c->p2p.gseq);
}
int athk_p2p_before_emf_apply(athk_ctx *c) {
  if (!c->p2p.on || c->rl.n == 0 || !c->cfg.mhd) return ATHK_OK;
  c->rl.erbuf = c->p2p.erecv + (c->p2p.eseq & 1) * c->p2p.ehalf;
  return p2p_wait(c, 1, c->p2p.eseq);
}
int athk_p2p_after_pack(athk_ctx *c, int which) {
  return p2p_signal(c, which, which == 0 ? c->p2p.gseq : c->p2p.eseq);
}
