/* athk.h -- C ABI of the B200 (sm_100a) hydro/MHD MeshBlock-update library (libathk.so).
 *
 * This is the drop-in boundary for the ONE hot path of Athena++ (v24.0): the per-MeshBlock
 * finite-volume update.  Athena++ has no FFI/plugin layer; its physics operators are C++ member
 * functions called from TimeIntegratorTaskList task bodies (src/task_list/time_integrator.cpp:
 * 1606-2318).  Each entry point below replaces one such member function and cites it.  All
 * signatures are plain C: pointers, sizes, scalars.  No torch/CUDA types appear.
 *
 * Execution model: one context == one GPU == the MeshBlocks of one rank.  An operator call acts
 * on ALL local MeshBlocks in one launch (the reference loops blocks on the host); calls are
 * enqueued on the context's stream and return immediately; only athk_sync, athk_download*,
 * athk_new_block_dt and athk_history synchronise.  Every function returns 0 on success, a
 * negative athk_status otherwise; athk_last_error() gives the message (the host side raises it
 * the way ATHENA_ERROR does, src/defs.hpp.in:155-159).  There is no CPU fallback: without a
 * CUDA device athk_create fails.
 *
 * Host array layout is AthenaArray's (src/athena_arrays.hpp:11-14): dense (n,k,j,i), i fastest,
 * ghost zones included: cell arrays (nvar, nc3, nc2, nc1); faces x1f (nc3,nc2,nc1+1),
 * x2f (nc3,nc2+1,nc1), x3f (nc3+1,nc2,nc1); edges x1e (nc3+1,nc2+1,nc1), x2e (nc3+1,nc2,nc1+1),
 * x3e (nc3,nc2+1,nc1+1), with nc = nx + 2*nghost.
 */
#ifndef ATHK_H_
#define ATHK_H_

#ifdef __cplusplus
extern "C" {
#endif

typedef struct athk_ctx athk_ctx;

enum athk_status {
  ATHK_OK = 0,
  ATHK_ERR_ARG = -1,      /* bad argument / unsupported configuration */
  ATHK_ERR_CUDA = -2,     /* CUDA runtime error (message has the cudaError string) */
  ATHK_ERR_NODEVICE = -3, /* no CUDA device: the library never falls back to the CPU */
  ATHK_ERR_STATE = -4     /* call out of order (e.g. geometry not set) */
};

/* Riemann solvers: one is linked per build in the reference (Makefile.in:73) */
enum athk_rsolver { ATHK_RS_HLLC = 0, ATHK_RS_HLLD = 1, ATHK_RS_ROE = 2, ATHK_RS_HLLE = 3, ATHK_RS_LLF = 4 };

/* Equation of state.  ATHK_EOS_ISOTHERMAL (eos/isothermal_hydro.cpp, isothermal_mhd.cpp): p = cs^2 rho,
 * no energy equation; Riemann solvers HLLE for hydro (hlle.cpp, the reference's default) and HLLD
 * for MHD (hlld_iso.cpp).  Registers keep the adiabatic layout; the energy component stays zero and
 * upload / download of U, U1, U2, W move the NHYDRO = 4 components the reference has. */
enum athk_eos { ATHK_EOS_ADIABATIC = 0, ATHK_EOS_ISOTHERMAL = 1,
                /* special-relativistic hydro (configure.py -s: eos/adiabatic_hydro_sr.cpp, hlle_rel.cpp):
                 * w = (rho, u^1, u^2, u^3, pgas) with u^i the four-velocity's spatial components,
                 * u = (D, M^1, M^2, M^3, E); rsolver HLLC, HLLE or LLF, no magnetic fields, uniform meshes;
                 * the block time step is the smallest cell width (c = 1, new_blockdt.cpp:130).  W(U)
                 * starts its iteration from W of two calls ago, like the reference's w / w1 swap
                 * (time_integrator.cpp:2216-2269). */
                ATHK_EOS_ADIABATIC_SR = 2 };

/* ATHK_ARITH_STRICT: the reference's double-precision arithmetic operation by operation (no FMA
 * contraction, correctly rounded divisions and square roots): results identical to the CPU build.
 * ATHK_ARITH_FAST (HLLC / HLLD): contraction on, divisions and square roots refined to ~1 ulp;
 * agrees with the strict set to rounding (tests/test_fast_arith.py), not bit for bit. */
enum athk_arith { ATHK_ARITH_STRICT = 0, ATHK_ARITH_FAST = 1 };

/* what the reference fixes at configure time (configure.py) plus <hydro> runtime parameters */
typedef struct athk_config {
  int nx1, nx2, nx3;   /* active zones per MeshBlock (<meshblock> nx1..3)          */
  int nghost;          /* NGHOST                                                  */
  int nblocks;         /* MeshBlocks owned by this context (this GPU)             */
  int mhd;             /* MAGNETIC_FIELDS_ENABLED                                 */
  int nscalars;        /* NSCALARS                                                */
  int rsolver;         /* enum athk_rsolver                                       */
  int device;          /* CUDA device ordinal                                     */
  int arith;           /* enum athk_arith: which kernel set the fused sweeps run             */
  double gamma;        /* <hydro> gamma                                           */
  double dfloor, pfloor, sfloor; /* floors; <=0 selects sqrt(1024*FLT_MIN) (eos.hpp:161) */
  int eos;             /* enum athk_eos: configure.py --eos                        */
  double iso_sound_speed; /* <hydro> iso_sound_speed (isothermal EOS)              */
  double gamma_max;    /* <hydro> gamma_max (SR velocity ceiling); <=0 selects 1000 (eos.cpp)  */
} athk_config;

/* device-resident arrays addressable through upload/download */
enum athk_array {
  ATHK_U = 0, ATHK_U1 = 1, ATHK_U2 = 2,   /* Hydro::u,u1,u2   (5, nc3,nc2,nc1)  hydro.hpp:43-51 */
  ATHK_W = 3,                             /* Hydro::w                                            */
  ATHK_BCC = 4,                           /* Field::bcc       (3, nc3,nc2,nc1)  field.hpp:38-51  */
  ATHK_B_X1F = 5, ATHK_B_X2F = 6, ATHK_B_X3F = 7,       /* Field::b                               */
  ATHK_B1_X1F = 8, ATHK_B1_X2F = 9, ATHK_B1_X3F = 10,   /* Field::b1                              */
  ATHK_B2_X1F = 11, ATHK_B2_X2F = 12, ATHK_B2_X3F = 13, /* Field::b2                              */
  ATHK_FLUX_X1 = 14, ATHK_FLUX_X2 = 15, ATHK_FLUX_X3 = 16, /* Hydro::flux[3] (5, face shape)      */
  ATHK_E_X1E = 17, ATHK_E_X2E = 18, ATHK_E_X3E = 19,    /* Field::e (edge shapes)                 */
  ATHK_E3_X1F = 20, ATHK_E2_X1F = 21, ATHK_E1_X2F = 22, /* face EMFs written by RiemannSolver     */
  ATHK_E3_X2F = 23, ATHK_E2_X3F = 24, ATHK_E1_X3F = 25,
  ATHK_WGHT_X1F = 26, ATHK_WGHT_X2F = 27, ATHK_WGHT_X3F = 28, /* Field::wght                      */
  ATHK_S = 29, ATHK_S1 = 30, ATHK_S2 = 31, ATHK_R = 32, /* PassiveScalars::s,s1,s2,r              */
  ATHK_SFLUX_X1 = 33, ATHK_SFLUX_X2 = 34, ATHK_SFLUX_X3 = 35,
  ATHK_W1 = 36,   /* Hydro::w1 with ATHK_EOS_ADIABATIC_SR: the array the next W(U) writes -- what it holds
                   * is that call's fallback state and first guess (a pgen may set it, gr_shock_tube.cpp) */
  ATHK_NARRAYS = 37
};

/* time-integrator registers (u,b,s move together only if the caller says so) */
enum athk_reg { ATHK_REG0 = 0, ATHK_REG1 = 1, ATHK_REG2 = 2 };

/* BoundaryFlag of a MeshBlock face that is a physical boundary (bvals/utils/boundary_flag.cpp);
 * faces with a neighbour (block / periodic) are ATHK_BC_NONE */
enum athk_bc { ATHK_BC_NONE = 0, ATHK_BC_REFLECT = 1, ATHK_BC_OUTFLOW = 2, ATHK_BC_USER = 3 };

/* ---- lifecycle ------------------------------------------------------------------------ */
int athk_create(const athk_config *cfg, athk_ctx **out);
void athk_destroy(athk_ctx *ctx);
const char *athk_last_error(void);
int athk_sync(athk_ctx *ctx);
/* number of kernel launches issued by this context so far (bench.py's gpu_launches) */
long long athk_launch_count(const athk_ctx *ctx);
/* Per-kernel timing with CUDA events on the context's stream.  enable != 0 starts recording an
 * event pair around every launch; a non-NULL report receives (after a synchronise) a JSON object
 * {"<entry point or sweep name>": {"ms": total, "launches": n}, ...} of what was recorded since
 * the last report.  Used by bench.py for the live roofline numbers. */
int athk_profile(athk_ctx *ctx, int enable, char *report, int report_len);
/* raw CUDA stream handle (cudaStream_t) so a host framework can order its own work after ours */
void *athk_stream(athk_ctx *ctx);

/* ---- geometry and topology -------------------------------------------------------------
 * Coordinates::x1f/x1v/dx1f... of block lb, host-computed and used verbatim
 * (src/coordinates/coordinates.cpp:100-324, cartesian.cpp:24-60).  xNf has ncN+1 entries,
 * xNv and dxNf have ncN entries. */
int athk_set_block_coords(athk_ctx *ctx, int lb,
                          const double *x1f, const double *x1v, const double *dx1f,
                          const double *x2f, const double *x2v, const double *dx2f,
                          const double *x3f, const double *x3v, const double *dx3f);
/* Same-level neighbour table, the result of BoundaryBase::SearchAndSetNeighbors
 * (src/bvals/bvals_base.cpp:301).  nbr[lb*27 + (ox3+1)*9 + (ox2+1)*3 + (ox1+1)] is
 *   >= 0   local block index on this context,
 *   -1     no neighbour (physical boundary),
 *   <= -2  remote neighbour: -(2 + remote slot), data arrives through athk_remote_* (multi-GPU). */
int athk_set_neighbors(athk_ctx *ctx, const int *nbr);

/* ---- data movement (Mesh::Initialize after pgen, outputs, restart) ---------------------- */
int athk_upload(athk_ctx *ctx, int array, int lb, const double *host);
int athk_download(athk_ctx *ctx, int array, int lb, double *host);
/* as athk_download but stream-ordered: host (pinned) is valid only after athk_sync */
int athk_download_async(athk_ctx *ctx, int array, int lb, double *host);
/* The same for ALL local MeshBlocks in one call and one PCIe transfer per array: host is
 * [nblocks][nvar][n3][n2][n1] dense.  active_only = 0: AthenaArray extents including ghost zones
 * (n = nc + a, as athk_upload); active_only = 1: the active zone only (n = nx + a; for a face
 * field the nx+1 faces of its own direction) -- what a caller that refreshes the ghost zones
 * afterwards (Mesh::Initialize) needs, 17 % fewer bytes for 64^3 blocks.  Stream-ordered; a
 * download's host buffer is valid after athk_sync. */
int athk_upload_all(athk_ctx *ctx, int array, const double *host, int active_only);
int athk_download_all_async(athk_ctx *ctx, int array, double *host, int active_only);
/* raw device pointer + padded strides of an array (for zero-copy consumers; elements) */
int athk_device_array(athk_ctx *ctx, int array, void **dptr, long long *stride_block,
                      long long *stride_var, long long *stride_k, long long *stride_j);

/* ---- operators (each acts on every local MeshBlock) ------------------------------------- */
/* EquationOfState::ConservedToPrimitive (eos/adiabatic_hydro.cpp:38, adiabatic_mhd.cpp:41; MHD
 * also Field::CalculateCellCenteredField field.cpp:112) from register `reg` into w/bcc, over the
 * interior plus every ghost side that has a neighbour (time_integrator.cpp:2206-2212); floors
 * are written back into the conserved register.  Scalars: eos_scalars.cpp:31. */
int athk_cons2prim(athk_ctx *ctx, int reg);
/* Hydro::CalculateFluxes(w,b,bcc,order) (hydro/calculate_fluxes.cpp:38): order 1 donor cell,
 * 2 PLM, 3 PPM; writes flux[3] and, for MHD, face EMFs + CT weights.  dt_mesh is
 * pmy_mesh->dt (the FULL step, hlld.cpp:52,373).  b is taken from register breg. */
int athk_calculate_fluxes(athk_ctx *ctx, int order, double dt_mesh, int breg);
/* PassiveScalars::CalculateFluxes (scalars/calculate_scalar_fluxes.cpp:41) */
int athk_calculate_scalar_fluxes(athk_ctx *ctx, int order);
/* Field::ComputeCornerE(w,bcc) (field/calculate_corner_e.cpp:28, 3-D branch 130-229) */
int athk_compute_corner_e(athk_ctx *ctx);
/* FaceCenteredBoundaryVariable::SendFluxCorrection / ReceiveFluxCorrection (+SetFluxBoundary
 * SameLevel + AverageFluxBoundary), bvals/fc/flux_correction_fc.cpp:51,623,689,1355,1610 */
int athk_send_flux_correction(athk_ctx *ctx);
int athk_recv_flux_correction(athk_ctx *ctx);
/* MeshBlock::WeightedAve (mesh/weighted_ave.cpp:35 cell-centred, :366 face-centred):
 * out = w0*out + w1*in1 + w2*in2 (+w3,w4 unused by vl2/rk2/rk3) with the zero-weight branches */
int athk_weighted_ave_cc(athk_ctx *ctx, int what /*0=hydro u,1=scalars s*/, int out_reg,
                         int in1_reg, int in2_reg, const double wght[5]);
int athk_weighted_ave_fc(athk_ctx *ctx, int out_reg, int in1_reg, int in2_reg,
                         const double wght[5]);
/* AthenaArray::SwapAthenaArray on the register pointers (athena_arrays.hpp:470) */
int athk_swap_registers(athk_ctx *ctx, int what /*0=u,1=b,2=s*/, int reg_a, int reg_b);
/* AthenaArray::ZeroClear on a register (time_integrator.cpp:1608-1619) */
int athk_zero_register(athk_ctx *ctx, int what, int reg);
/* Hydro::AddFluxDivergence(wght,u_out) (hydro/add_flux_divergence.cpp:39) */
int athk_add_flux_divergence(athk_ctx *ctx, double wght, int reg);
int athk_add_scalar_flux_divergence(athk_ctx *ctx, double wght, int reg);
/* Field::CT(wght,b_out) (field/ct.cpp:31) */
int athk_ct(athk_ctx *ctx, double wght, int reg);
/* BoundaryVariable::SendBoundaryBuffers+ReceiveBoundaryBuffers+SetBoundaries for same-level
 * neighbours (bvals/bvals_var.cpp:212-326, cc/bvals_cc.cpp:287-420, fc/bvals_fc.cpp:344-680);
 * what: bit0 hydro u, bit1 field b, bit2 scalars s; taken from register reg. */
int athk_exchange_ghosts(athk_ctx *ctx, int what_mask, int reg);
/* BoundaryValues::block_bcs of local block lb: flags[6] = {inner_x1, outer_x1, inner_x2, outer_x2,
 * inner_x3, outer_x3} as enum athk_bc.  The neighbour table must hold -1 across those faces.
 * Default: all ATHK_BC_NONE (fully periodic / interior). */
int athk_set_boundary_flags(athk_ctx *ctx, int lb, const int flags[6]);
/* BoundaryValues::ApplyPhysicalBoundaries (bvals/bvals.cpp:453): reflecting / outflow functions on
 * the primitives w (and r) and on b, face by face in the reference's order x1, x2, x3, each
 * followed by bcc, PrimitiveToConserved and s = r*rho on that ghost slab.  Call after
 * athk_cons2prim (task PHY_BVAL, time_integrator.cpp:2276); athk_stage does it itself.
 * No-op when no local block has a physical boundary. */
int athk_apply_physical_boundaries(athk_ctx *ctx);
/* User-enrolled boundary functions: BValFunc (athena.hpp:214-217), Mesh::EnrollUserBoundaryFunction
 * (mesh/mesh.cpp:1288), dispatched by BoundaryValues::DispatchBoundaryFunctions (bvals/bvals.cpp:673)
 * for faces flagged ATHK_BC_USER.  The function is a HOST callback invoked once per flagged local
 * MeshBlock at the face's place in the reference's x1, x2, x3 order; it fills the ghost zones of the
 * PRIMITIVES w (and of b, r) of that face by enqueueing its own device work on args->stream
 * (asynchronous, stream-ordered after the library's kernels), exactly what a BValFunc does to
 * `prim` and `b`; the library then forms bcc, U(W) and s on the slab as for the built-in functions.
 * Transverse limits (il..ku) are the ones the reference passes (extended into ghost zones that the
 * neighbour exchange filled).  Arrays use the padded-box layout: element (n,k,j,i) of w at
 * w[n*stride_n + k*stride_k + j*stride_j + i]; b1/b2/b3 likewise without n.  Return non-zero to
 * abort (reported as ATHK_ERR_STATE). */
typedef struct athk_bval_args {
  int lb, face;                  /* local block; 0..5 = inner_x1, outer_x1, inner_x2, ... */
  double time, dt;               /* as passed to BValFunc (athk_set_boundary_time) */
  int il, iu, jl, ju, kl, ku, ngh;
  double *w, *b1, *b2, *b3, *r;  /* DEVICE pointers of this block (b*, r null without MHD / scalars) */
  long long stride_j, stride_k, stride_n;
  int nscalars;
  void *stream;                  /* cudaStream_t the work must be enqueued on */
} athk_bval_args;
typedef int (*athk_bval_fn)(void *user, const athk_bval_args *args);
int athk_enroll_user_boundary(athk_ctx *ctx, int face, athk_bval_fn fn, void *user);
/* User-enrolled explicit source terms: SrcTermFunc (athena.hpp:218-222), Mesh::EnrollUserExplicit
 * SourceFunction (mesh.cpp:1345), called last by HydroSourceTerms::AddSourceTerms (hydro/srcterms/
 * hydro_srcterms.cpp:151-154).  Like the boundary hook, the function is a HOST callback invoked once per
 * local MeshBlock by athk_add_source_terms (and by athk_stage, after INT_HYD); it adds its terms to
 * the conserved register u (and s) of the active zone from the stage-input primitives w (bcc, r) by
 * enqueueing device work on args->stream.  time = beginning of the stage (athk_set_source_time,
 * time_integrator.cpp:1901-1903), dt = beta*dt_mesh.  Same array layout as athk_bval_args. */
typedef struct athk_src_args {
  int lb;
  double time, dt;
  int is, ie, js, je, ks, ke;
  const double *w, *bcc, *r;     /* DEVICE pointers of this block (bcc, r null without MHD / scalars) */
  double *u, *s;
  long long stride_j, stride_k, stride_n;
  int nscalars;
  void *stream;
} athk_src_args;
typedef int (*athk_src_fn)(void *user, const athk_src_args *args);
int athk_enroll_user_source(athk_ctx *ctx, athk_src_fn fn, void *user);
int athk_set_source_time(athk_ctx *ctx, double time);
/* ---- static mesh refinement (hydro and MHD) -------------------------------------------------
 * The host owns the block tree and the neighbour levels (MeshBlockTree, BoundaryBase::
 * SearchAndSetNeighbors) and hands the library, per phase of the boundary exchange, the list of
 * index boxes the reference's Load/Set/Restrict/Prolongate functions loop over; the library holds
 * the coarse buffers and does the arithmetic on the device (athena_b200/csrc/athk_refine.h).
 * Neighbour-table entries towards a coarser or finer neighbour are ATHK_NBR_REFINED: such ghost
 * zones are filled by the box operations, and count as "has a neighbour" for W(U).
 *   kind  1 RestrictCellCenteredValues own u -> own coarse buffer   (mesh_refinement.cpp:116;
 *           LoadBoundaryBufferToCoarser bvals_cc.cpp:309, RestrictGhostCellsOnSameLevel bvals_refine.cpp:263)
 *         2 SetBoundaryFromFiner: u[dst] box <- coarse buffer of src          (bvals_cc.cpp:485)
 *         3 LoadBoundaryBufferToFiner + SetBoundaryFromCoarser: coarse buffer of dst <- u[src]  (:335,:424)
 *         4 ConservedToPrimitive on the coarse buffer                      (bvals_refine.cpp:437)
 *         5 ProlongateCellCenteredValues coarse primitives -> w            (mesh_refinement.cpp:464)
 *         6 PrimitiveToConserved on the prolongated ghost zones            (bvals_refine.cpp:655)
 *         7 flux correction: flux[dir] of dst on a face box <- area-weighted flux of the finer src
 *           (flux_correction_cc.cpp:77-160, 222-275)
 *         8 ApplyPhysicalBoundariesOnCoarseLevel (bvals_refine.cpp:445-530): the reflecting (src = 1) or
 *           outflow (src = 2) function on the coarse primitives, one ghost layer of direction dir copied
 *           from the active coarse cell so[dir]
 * Passive scalars: a box of kind 1, 2, 3, 5, 7 or 8 with pad = 1 acts on s / coarse_s_ / r / coarse_r_ /
 * s_flux instead of the hydro variables (MeshRefinement::pvars_cc_, bvals_refine.cpp:560-575); kinds 4
 * and 6 convert them along with the hydro variables (eos/eos_scalars.cpp).
 * MHD (dir = component of the face field / of the edge EMF):
 *         9 SetBoundarySameLevel on a refined mesh (edge and corner neighbours also set the faces
 *           they share: bvals_fc.cpp:583-680): b[dst].dir box <- b[src].dir
 *        10 RestrictFieldX1/X2/X3 own b -> own coarse_b        (mesh_refinement.cpp:272-460;
 *           LoadBoundaryBufferToCoarser bvals_fc.cpp:403, RestrictGhostCellsOnSameLevel bvals_refine.cpp:318)
 *        11 SetBoundaryFromFiner: b[dst].dir box <- coarse_b of src          (bvals_fc.cpp:813)
 *        12 LoadBoundaryBufferToFiner + SetBoundaryFromCoarser: coarse_b of dst <- b[src] (:476,:690)
 *        13 ProlongateSharedFieldX1/X2/X3 over a coarse face box   (mesh_refinement.cpp:776-1220)
 *        14 ProlongateInternalField over a coarse cell box         (mesh_refinement.cpp:1223)
 *        15 ClearCoarseFluxBoundary: e[dst].dir box = 0            (flux_correction_fc.cpp:1214)
 *        16 LoadFluxBoundaryBufferToCoarser + SetFluxBoundaryFromFiner: e[dst].dir box += the
 *           length-weighted pairs of the finer src's surface EMFs as athk_send_flux_correction packed
 *           them (flux_correction_fc.cpp:315, 944)
 *        17 AverageFluxBoundary: e[dst].dir box *= 1/src            (flux_correction_fc.cpp:1355)
 *        18 as 8 for coarse_b (reflect_fc.cpp, outflow_fc.cpp with ngh = 1): the ghost layer of boundary
 *           direction pad copied from index so[pad] of component dir, negated when src = 1
 *   kinds 4 and 6 on an MHD context also form the cell-centred field of the coarse buffer / of the
 *   prolongated ghost zones (bvals_refine.cpp:437, 640-657).
 * Boxes are inclusive (i, j, k) ranges of the destination; so = source origin matching lo.  Boxes of
 * one list run concurrently: lists whose boxes overwrite each other with different values (face,
 * edge, corner neighbours of the field; successive EMF contributions) are the host's to keep apart. */
#define ATHK_NBR_REFINED (-1000000)
typedef struct athk_box_op {
  int kind, dst, src, dir;
  int lo[3], hi[3];
  int so[3];
  int pad;
  int peer;   /* 0: block src is a block of this context; r + 1: a block of GPU r's context (athk_refine_set_peers) */
} athk_box_op;
enum { ATHK_REFINE_NPHASES = 96 };
/* allocates coarse_cons_ / coarse_prim_ (hydro/hydro.cpp:60-75): 2 x 5 x (nx/2 + 2 nghost)^3 per block;
 * MHD: also coarse_b_ / coarse_bcc_ (field/field.cpp:60-75) */
int athk_refine_enable(athk_ctx *ctx);
/* Coordinates::x?v of the block and of its coarse twin pcoarsec (mesh_refinement.cpp:60-75), lengths
 * nc? and nx?/2 + 2 nghost */
int athk_refine_set_block_coords(athk_ctx *ctx, int lb, const double *x1v, const double *x2v, const double *x3v,
                                 const double *cx1v, const double *cx2v, const double *cx3v);
/* Refined mesh over several GPUs of one process: peers[r] = the context of GPU r (n <= 8; the entry of
 * the calling context itself is ignored).  A box whose source block lives on another GPU names it with
 * peer = r + 1 and the kernel reads that context's arrays directly (peer access must be enabled,
 * athk_p2p_enable_peer); the caller orders the lists across GPUs (athk_sync between dependent lists).
 * Call before athk_refine_set_ops. */
int athk_refine_set_peers(athk_ctx *ctx, int n, athk_ctx *const *peers);
/* The coarse primitives and coarse field of local block lb in the reference's dense layouts
 * (coarse_prim_ [5][cn3][cn2][cn1], coarse_b_.x1f [cn3][cn2][cn1+1], .x2f [cn3][cn2+1][cn1], .x3f
 * [cn3+1][cn2][cn1] with cn = nx/2 + 2 nghost; the field pointers may be NULL), to the host
 * (to_device = 0) or back: a user-enrolled boundary function acts on them between the coarse W(U) and
 * the prolongation (bvals_refine.cpp:445-530).  Synchronous. */
int athk_refine_coarse_transfer(athk_ctx *ctx, int lb, int to_device, double *coarse_prim, double *cb1, double *cb2,
                                double *cb3);
/* the boxes of one phase (0 .. ATHK_REFINE_NPHASES-1; their meaning is the host's), copied to the device */
int athk_refine_set_ops(athk_ctx *ctx, int phase, const athk_box_op *ops, int nops);
/* one launch over all boxes of the phase; reg = the conserved (and field) register the boxes act on */
int athk_refine_run(athk_ctx *ctx, int phase, int reg);
/* time and dt handed to the user functions by the next athk_apply_physical_boundaries / athk_stage
 * (time_integrator.cpp:2276-2290: end-of-stage time and the stage's dt) */
int athk_set_boundary_time(athk_ctx *ctx, double time, double dt);
/* Floors of local block lb.  In the reference every MeshBlock owns an EquationOfState
 * (mesh/meshblock.cpp:160-190) whose floors come from pin->GetOrAddReal("hydro","dfloor"|"pfloor"|
 * "sfloor", sqrt(1024*FLT_MIN)) (eos/adiabatic_mhd.cpp:29-31): the first block gets the default in
 * full precision, the blocks built after it read back the 6-digit string the first call stored
 * (parameter_input.cpp:603-623), so floors may differ from block to block.  Default: athk_config's
 * floors for every block. */
int athk_set_block_floors(athk_ctx *ctx, int lb, double dfloor, double pfloor, double sfloor);
/* HydroSourceTerms (hydro/srcterms/hydro_srcterms.cpp:42-46): the constant accelerations
 * <hydro> grav_acc1 / grav_acc2 / grav_acc3.  All zero (the default) = no source terms. */
int athk_set_constant_acceleration(athk_ctx *ctx, double g1, double g2, double g3);
/* Task SRC_TERM, TimeIntegratorTaskList::AddSourceTerms (task_list/time_integrator.cpp:1890) ->
 * HydroSourceTerms::ConstantAcceleration (hydro/srcterms/constant_acc.cpp:25): u(IM*) += dt*rho*g,
 * u(IEN) += dt*rho*g*v on the active zone, rho and v from the stage-input primitives w; beta_dt =
 * stage_wghts.beta * dt.  Call after athk_integrate_hydro; athk_stage does it itself.  No-op
 * without accelerations. */
int athk_add_source_terms(athk_ctx *ctx, double beta_dt);
/* Hydro::NewBlockTimeStep (hydro/new_blockdt.cpp:45): per-block min dt * cfl, blocks 0..nblocks-1
 * into dt_blocks (may be NULL), returns the min over local blocks in *dt_min. Synchronises. */
int athk_new_block_dt(athk_ctx *ctx, double cfl_number, double *dt_blocks, double *dt_min);
/* Ask the NEXT athk_cons2prim (the last stage's W(U)) to fold Hydro::NewBlockTimeStep into the
 * same pass over w; the following athk_new_block_dt then only reads the per-block minima. */
int athk_request_block_dt(athk_ctx *ctx);
/* History sums (outputs/history.cpp:196-248): mass, 3 momenta, 3 KE, tot-E, [3 ME], [scalars];
 * sums[] accumulates over local blocks (volume weighted). Synchronises. */
int athk_history(athk_ctx *ctx, double *sums, int nsums);

/* ---- fused stage (same results, fewer passes over HBM) ---------------------------------- */
/* One full integrator stage for every local block, equivalent to the task sequence
 * CALC_HYDFLX, CALC_FLDFLX, SEND/RECV_FLDFLX, INT_HYD, INT_FLD, SEND/RECV/SETB_HYD/FLD,
 * CONS2PRIM (time_integrator.cpp:905-1206) with the weights of stage_wghts[stage-1]:
 *   delta, gamma_1..3, beta (time_integrator.cpp:111-606).  order as athk_calculate_fluxes. */
typedef struct athk_stage_wghts {
  double delta, gamma_1, gamma_2, gamma_3, beta;
} athk_stage_wghts;
int athk_stage(athk_ctx *ctx, int stage, int order, const athk_stage_wghts *sw, double dt_mesh);

/* The same stage split at the two points where data of other GPUs is needed, so that the host
 * can run the transport (NCCL send/recv over NVLink) in between:
 *   phase 0: sweeps, corner E, pack EMF surfaces (local + remote links)      -> send/recv EMF
 *   phase 1: EMF correction, CT, pack ghost messages of remote links         -> send/recv ghosts
 *   phase 2: ghost fill (local neighbours + received messages), W(U)
 * athk_stage == phases 0,1,2 back to back (single GPU, or no remote links).
 * Overlap variant 10, 13, 11, 14, 12 (same results): 10 computes corner E on the block surface
 * only before packing, 13 the interior edges (runs while the EMF messages fly), 11 == 1, 14 is
 * the ghost fill from same-GPU neighbours + W(U) on the active zone (runs while the ghost messages
 * fly), 12 the ghost fill from the received messages + W(U) on the ghost shell. */
int athk_stage_phase(athk_ctx *ctx, int phase, int stage, int order, const athk_stage_wghts *sw,
                     double dt_mesh);

/* ---- multi-GPU: neighbours on another GPU ("remote links") ---------------------------------
 * BoundaryVariable::SendBoundaryBuffers / ReceiveBoundaryBuffers across ranks (bvals_var.cpp:
 * 212-262; MPI persistent requests in the reference, bvals_cc.cpp:614-663).  A link is
 * (local block lb, neighbour offset ox1..3); entry -(2+link) in the neighbour table refers to it.
 * The caller owns four device buffers (ghost send/recv, EMF send/recv) and assigns every link an
 * element offset in each, normally grouping links per peer rank so that one message per peer and
 * quantity crosses NVLink.  Message sizes: */
long long athk_remote_ghost_count(athk_ctx *ctx, int ox1, int ox2, int ox3); /* u (+b) (+s) slab */
long long athk_remote_emf_count(athk_ctx *ctx, int ox1, int ox2, int ox3);   /* 0 for corners   */
int athk_remote_configure(athk_ctx *ctx, int nlinks, const int *lb, const int *ox1,
                          const int *ox2, const int *ox3, const long long *ghost_send_off,
                          const long long *ghost_recv_off, const long long *emf_send_off,
                          const long long *emf_recv_off);
int athk_remote_set_buffers(athk_ctx *ctx, double *send_ghost, double *recv_ghost,
                            double *send_emf, double *recv_emf);
/* LoadBoundaryBufferSameLevel for every remote link (cc/bvals_cc.cpp:287, fc/bvals_fc.cpp:344) */
int athk_remote_pack_ghosts(athk_ctx *ctx, int reg);
/* LoadFluxBoundaryBufferSameLevel for every remote link (fc/flux_correction_fc.cpp:51);
 * call after athk_send_flux_correction */
int athk_remote_pack_emf(athk_ctx *ctx);

/* ---- multi-GPU: peer-to-peer transport over NVLink -------------------------------------------
 * CopyVariableBufferSameProcess (bvals_var.cpp:108-119) extended across GPUs: instead of handing
 * the packed messages to NCCL/MPI, the pack kernels store every message straight into the
 * RECEIVER's buffer through a peer-mapped pointer, a one-thread-per-peer kernel then raises the
 * receiver's arrival flag (sequence number, after a system-scope fence), and the receiver's
 * ghost-fill / EMF-correction is preceded by a kernel that waits for the flags of all its peers.
 * Receive buffers are double-buffered by the parity of the sequence number, which together with
 * the stream order of pack -> wait on every rank makes reuse safe without an acknowledgement.
 * How the peer pointers are obtained is the host's business: CUDA IPC between processes
 * (athk_p2p_alloc / athk_p2p_open), or plain pointers inside one process (athk_p2p_enable_peer). */
int athk_p2p_alloc(athk_ctx *ctx, long long nbytes, void **dptr, unsigned char ipc_handle[64]);
int athk_p2p_open(athk_ctx *ctx, const unsigned char ipc_handle[64], void **dptr);
int athk_p2p_close(athk_ctx *ctx, void *dptr);    /* a pointer returned by athk_p2p_open */
int athk_p2p_free(athk_ctx *ctx, void *dptr);     /* a pointer returned by athk_p2p_alloc */
int athk_p2p_enable_peer(athk_ctx *ctx, int peer_device);
/* ghost_dst / emf_dst: [2][nlinks] device pointers, where link q's message goes in the receiver's
 * buffer half 0 / 1; recv_ghost / recv_emf: this context's own buffers, two halves of
 * ghost_half / emf_half elements laid out as athk_remote_configure's *_recv_off describe;
 * peer_flags[p]: the two 64-bit words (ghost, emf) in peer p's flag array that belong to this
 * rank; my_flags: this rank's array [npeers][2] (zero-initialised). */
int athk_remote_set_p2p(athk_ctx *ctx, int npeers, double *const *ghost_dst, double *const *emf_dst,
                        double *recv_ghost, long long ghost_half, double *recv_emf, long long emf_half,
                        unsigned long long *const *peer_flags, unsigned long long *my_flags);
/* number of flag waits that ran into their time limit (0 in a healthy run) */
long long athk_p2p_timeouts(athk_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* ATHK_H_ */
