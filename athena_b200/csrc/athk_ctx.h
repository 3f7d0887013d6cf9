// athk_ctx.h -- the context behind include/athk.h's opaque athk_ctx, error plumbing and the launch
// macros; shared by the translation units of libathk.so (athk_api.cu and the per-Riemann-solver
// athk_rs_*.cu, which only hold kernel instantiations).
#ifndef ATHK_CTX_H_
#define ATHK_CTX_H_
#include "../../include/athk.h"

#include <float.h>
#include <math.h>
#include <stdarg.h>
#include <string.h>

#include <string>
#include <vector>

#include "athk_backend.h"
#include "athk_ops.h"
#include "athk_sweep.h"
#include "athk_scalars.h"
#include "athk_bcs.h"
#include "athk_refine.h"
#ifndef ATHK_EMULATE
#include "athk_sweep_kernels.cuh"
#endif

using namespace athk;

int athk_fail(int code, const char *fmt, ...);   // records athk_last_error(), returns code
#define fail athk_fail

#define CU(call)                                                                           \
  do {                                                                                     \
    cudaError_t e_ = (call);                                                               \
    if (e_ != cudaSuccess)                                                                 \
      return fail(ATHK_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_),   \
                  __FILE__, __LINE__);                                                     \
  } while (0)

struct athk_ctx {
  athk_config cfg;
  Grid g;
  Geom geo;
  EosParams eos;
  EmfSurf sf;
  cudaStream_t stream;
  long long launches;
  bool coords_set, nbr_set;
  int nremote;
  // storage
  Real *u[3];       // hydro registers [nb][5][sn]
  Real *w;          // [nb][5][sn]
  Real *w_alt;      // SR only: Hydro::w1, the array the next W(U) writes (swapped with w by every W(U))
  Real *bcc;        // [nb][3][sn]
  Real *b[3];       // field registers [nb][3][sn]
  Real *flux[3];    // [nb][5][sn]
  Real *ef;         // [nb][6][sn]
  Real *wct;        // [nb][3][sn]
  Real *e;          // [nb][3][sn]
  Real *s[3];       // scalar registers [nb][ns][sn]
  Real *r;          // [nb][ns][sn]
  Real *sflux[3];   // [nb][ns][sn]
  Real *emfbuf;     // [nb][sf.per_block]
  // remote (other-GPU) neighbour links
  RemoteLinks rl;
  int *rl_int;        // device [4][n]: lb,o1,o2,o3
  idx_t *rl_off;      // device [5][n]: gsend,grecv,esend,erecv,ebase
  std::vector<int> h_rl_int;
  std::vector<idx_t> h_rl_off;
  idx_t max_ghost_count, max_emf_count;
  SlabTable *slab_tab;   // device
  bool dt_requested, dt_valid;   // NewBlockTimeStep fused into the next / last W(U)
  int kt_corner, kt_ct;          // planes per thread of the k-marching corner-E / CT kernels
  // optional per-kernel timing with CUDA events on the launching stream (athk_profile)
  bool prof_on;
  struct ProfRec { const char *name; cudaEvent_t a, b; };
  std::vector<ProfRec> prof;
  Real *geo_store;  // 9 arrays [nb][PG]
  int *nbr;         // [nb][27]
  Real *dt_blocks;  // [nb]
  std::vector<int> h_nbr;
  Real *stage;                  // dense device staging buffer of athk_upload / athk_download
  size_t stage_count;
  // physical boundaries: block_bcs per local block (enum athk_bc), device copy uploaded lazily
  std::vector<int> h_bcflags;   // [nb][6]
  int *bcflags;                 // device
  bool bc_dirty;
  int bc_any[6];                // does any local block have a physical boundary on face f
  // static mesh refinement (athk_refine_*): coarse buffers, coordinates, per-phase box lists
  struct Refine {
    bool on = false;
    CoarseGrid cg;
    Real *cu = nullptr, *cw = nullptr, *xv = nullptr, *cxv = nullptr;
    Real *cb = nullptr, *cbcc = nullptr;   // MHD: coarse_b_, coarse_bcc_
    Real *cs = nullptr, *cr = nullptr;     // passive scalars: coarse_s_, coarse_r_
    athk_ctx *peers[ATHK_MAX_PEERS] = {};   // contexts of the other GPUs of this process (athk_refine_set_peers)
    int npeers = 0;
    BoxOp *ops[ATHK_REFINE_NPHASES] = {};
    int nops[ATHK_REFINE_NPHASES] = {};
    long long maxvol[ATHK_REFINE_NPHASES] = {};
  } rf;
  athk_bval_fn user_bc[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // enrolled BValFuncs
  void *user_bc_data[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  double bc_time = 0.0, bc_dt = 0.0;
  athk_src_fn user_src = nullptr;   // enrolled SrcTermFunc
  void *user_src_data = nullptr;
  double src_time = 0.0;
  double gacc[3] = {0.0, 0.0, 0.0};   // HydroSourceTerms: <hydro> grav_acc1/2/3 (0 = none)
  std::vector<Real> h_floors;         // per-block (dfloor, pfloor, sfloor); empty = uniform
  Real *d_floors = nullptr;
  std::vector<Real> h_geo;   // host mirror of geo_store
  // peer-to-peer transport of the remote messages (athk_remote_set_p2p)
  struct P2P {
    bool on = false;
    int npeers = 0;
    Real **gdst = nullptr, **edst = nullptr;          // device arrays [2][nlinks] of destination pointers
    Real *grecv = nullptr, *erecv = nullptr;          // own receive buffers (two halves)
    idx_t ghalf = 0, ehalf = 0;
    unsigned long long **peer_flags = nullptr;        // device array [npeers]
    unsigned long long *my_flags = nullptr;           // [npeers][2]
    unsigned long long gseq = 0, eseq = 0;            // messages sent so far
    unsigned long long *timeouts = nullptr;           // device counter
    // the hydro (and scalar) part of a ghost message is final once the sweeps are done: it is packed
    // -- i.e. sent over NVLink -- on a side stream under corner E / EMF correction / CT
    cudaStream_t side = nullptr;
    cudaEvent_t ev_ready = nullptr, ev_packed = nullptr;
    bool early_pending = false;
  } p2p;
};

void athk_prof_begin(athk_ctx *c, const char *name);
void athk_prof_end(athk_ctx *c);
#define prof_begin athk_prof_begin
#define prof_end athk_prof_end
// one kernel launch, optionally bracketed by CUDA events; named after the API entry point
#define LAUNCH_NAMED(c, name, call) \
  do {                            \
    prof_begin((c), (name));      \
    CU(call);                     \
    prof_end(c);                  \
    (c)->launches++;              \
  } while (0)
#define LAUNCH(c, range, functor) \
  LAUNCH_NAMED(c, __func__, launch_points((c)->stream, (c)->g.nb, (range), (functor)))

#endif  // ATHK_CTX_H_
