// athena.hpp -- C++ host side of athena-b200: the reference's Mesh / MeshBlock / Coordinates /
// ParameterInput / TimeIntegratorTaskList / pgen API (names, argument meaning, error behaviour)
// re-implemented as thin glue above the C ABI of libathk.so (include/athk.h).  Field data live in
// HBM; MeshBlock::phydro->u etc. are HOST arrays used by problem generators and outputs only
// (uploaded after ProblemGenerator, downloaded at output times), exactly the access pattern the
// reference's pgens and outputs have.
//
// What the reference fixes at configure time (configure.py: --prob, -b, --flux, --nghost,
// --nscalars) is a run-time choice here (Config), taken from the command line or from the
// athinput's `<comment> configure = ...` line.
//
// Reference: src/athena.hpp, src/athena_arrays.hpp, src/parameter_input.hpp, src/mesh/mesh.hpp,
// src/coordinates/coordinates.hpp, src/hydro/hydro.hpp, src/field/field.hpp,
// src/task_list/task_list.hpp (file:line cited at each definition).
#ifndef ATHENA_B200_HOST_ATHENA_HPP_
#define ATHENA_B200_HOST_ATHENA_HPP_

#include <cstdint>
#include <cstring>
#include <iostream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../../include/athk.h"

using Real = double;   // athena.hpp:39

// athena.hpp:169-192
enum { IDN = 0, IM1 = 1, IM2 = 2, IM3 = 3, IEN = 4 };
enum { IVX = 1, IVY = 2, IVZ = 3, IPR = 4, IBY = 5, IBZ = 6 };
enum { IB1 = 0, IB2 = 1, IB3 = 2 };
enum { X1DIR = 0, X2DIR = 1, X3DIR = 2 };
// NHYDRO: 5, or 4 with the isothermal equation of state (run-time here, see Config)

// defs.hpp.in:137-159
#define PI 3.1415926535897932
#define TWO_PI 6.2831853071795862
#define ONE_3RD 0.33333333333333333
#define TWO_3RD 0.66666666666666667
#define TINY_NUMBER 1.0e-20
#define SQR(x) ((x) * (x))
#define SIGN(x) (((x) < 0.0) ? -1.0 : 1.0)
#define ATHENA_ERROR(x) throw std::runtime_error(x.str().c_str())

// run-time stand-in for the configure-time macros of src/defs.hpp
struct Config {
  std::string problem = "";
  bool mhd = false;           // MAGNETIC_FIELDS_ENABLED
  std::string flux = "";      // RIEMANN_SOLVER (default hllc / hlld)
  int nghost = 2;             // NGHOST
  int nscalars = 0;           // NSCALARS
  bool isothermal = false;    // --eos=isothermal: NON_BAROTROPIC_EOS == 0, NHYDRO == 4
  bool sr = false;            // -s: RELATIVISTIC_DYNAMICS (special relativity, hydro)
  int nhydro() const { return isothermal ? 4 : 5; }
  int nwave() const { return nhydro() + (mhd ? 2 : 0); }
};
extern Config g_config;
#define MAGNETIC_FIELDS_ENABLED (g_config.mhd)
#define NGHOST (g_config.nghost)
#define NSCALARS (g_config.nscalars)
#define NWAVE (g_config.nwave())
#define NON_BAROTROPIC_EOS (!g_config.isothermal)
#define RELATIVISTIC_DYNAMICS (g_config.sr)
#define NHYDRO (g_config.nhydro())

//----------------------------------------------------------------------------------------
// AthenaArray (athena_arrays.hpp:25-171): dense, last index fastest
template <typename T>
class AthenaArray {
 public:
  void NewAthenaArray(int n1) { Alloc(1, 1, 1, n1); }
  void NewAthenaArray(int n2, int n1) { Alloc(1, 1, n2, n1); }
  void NewAthenaArray(int n3, int n2, int n1) { Alloc(1, n3, n2, n1); }
  void NewAthenaArray(int n4, int n3, int n2, int n1) { Alloc(n4, n3, n2, n1); }
  int GetDim1() const { return n1_; }
  int GetDim2() const { return n2_; }
  int GetDim3() const { return n3_; }
  int GetDim4() const { return n4_; }
  std::size_t GetSize() const { return d_.size(); }
  std::size_t GetSizeInBytes() const { return d_.size() * sizeof(T); }
  bool IsAllocated() const { return !d_.empty(); }
  T *data() { return d_.data(); }
  const T *data() const { return d_.data(); }
  T &operator()(int i) { return d_[i]; }
  const T &operator()(int i) const { return d_[i]; }
  T &operator()(int j, int i) { return d_[i + n1_ * j]; }
  const T &operator()(int j, int i) const { return d_[i + n1_ * j]; }
  T &operator()(int k, int j, int i) { return d_[i + n1_ * (j + n2_ * (std::size_t)k)]; }
  const T &operator()(int k, int j, int i) const { return d_[i + n1_ * (j + n2_ * (std::size_t)k)]; }
  T &operator()(int n, int k, int j, int i) { return d_[i + n1_ * (j + n2_ * (k + n3_ * (std::size_t)n))]; }
  const T &operator()(int n, int k, int j, int i) const {
    return d_[i + n1_ * (j + n2_ * (k + n3_ * (std::size_t)n))];
  }
  void ZeroClear() { std::fill(d_.begin(), d_.end(), T()); }
  void SwapAthenaArray(AthenaArray<T> &o) {
    d_.swap(o.d_);
    std::swap(n1_, o.n1_); std::swap(n2_, o.n2_); std::swap(n3_, o.n3_); std::swap(n4_, o.n4_);
  }

 private:
  void Alloc(int n4, int n3, int n2, int n1) {
    n1_ = n1; n2_ = n2; n3_ = n3; n4_ = n4;
    d_.assign((std::size_t)n4 * n3 * n2 * n1, T());
  }
  std::vector<T> d_;
  int n1_ = 0, n2_ = 0, n3_ = 0, n4_ = 0;
};

struct FaceField { AthenaArray<Real> x1f, x2f, x3f; };   // athena.hpp:128-138

//----------------------------------------------------------------------------------------
// ParameterInput (parameter_input.hpp:70-119): <block> / name = value # comment
class ParameterInput {
 public:
  struct Line { std::string name, value, comment; };
  struct Block { std::string name; std::vector<Line> lines; };
  void LoadFromFile(const std::string &fname);
  void LoadFromStream(std::istream &is);
  void ModifyFromCmdline(int argc, char *argv[]);
  void ParameterDump(std::ostream &os) const;
  int DoesParameterExist(const std::string &block, const std::string &name) const;
  bool DoesBlockExist(const std::string &block) const { return Find(block) != nullptr; }
  int GetInteger(const std::string &block, const std::string &name) const;
  Real GetReal(const std::string &block, const std::string &name) const;
  bool GetBoolean(const std::string &block, const std::string &name) const;
  std::string GetString(const std::string &block, const std::string &name) const;
  int GetOrAddInteger(const std::string &block, const std::string &name, int def);
  Real GetOrAddReal(const std::string &block, const std::string &name, Real def);
  bool GetOrAddBoolean(const std::string &block, const std::string &name, bool def);
  std::string GetOrAddString(const std::string &block, const std::string &name, const std::string &def);
  int SetInteger(const std::string &block, const std::string &name, int v);
  Real SetReal(const std::string &block, const std::string &name, Real v);
  std::string SetString(const std::string &block, const std::string &name, const std::string &v);
  const std::vector<Block> &blocks() const { return blocks_; }
  void RollbackNextTime();               // restart: parameter_input.cpp:754
  void ForwardNextTime(Real mesh_time);  // restart: parameter_input.cpp:790

 private:
  const Block *Find(const std::string &b) const;
  Block *FindOrAdd(const std::string &b);
  const Line *FindLine(const std::string &b, const std::string &n, const char *who) const;
  void Add(Block *pb, const std::string &n, const std::string &v, const std::string &c);
  std::vector<Block> blocks_;
};

//----------------------------------------------------------------------------------------
struct RegionSize {   // mesh.hpp / athena.hpp:150-160 (same member order as the reference: the
  Real x1min, x2min, x3min;   // struct is written raw into restart files)
  Real x1max, x2max, x3max;
  Real x1rat, x2rat, x3rat;
  int nx1, nx2, nx3;
};
struct LogicalLocation {   // athena.hpp:104-118
  std::int64_t lx1, lx2, lx3;
  int level;
};

enum class BoundaryFlag { block = -1, undef, reflect, outflow, user, periodic };

class Mesh;
class MeshBlock;

// Coordinates (coordinates/coordinates.hpp, cartesian.cpp): 1-D arrays, uniform Cartesian
class Coordinates {
 public:
  // coarse = the twin MeshRefinement::pcoarsec (coordinates.cpp:33-37, 108-160): nx/2 cells, same NGHOST
  Coordinates(MeshBlock *pmb, Mesh *pm, bool coarse = false);
  AthenaArray<Real> x1f, x2f, x3f, x1v, x2v, x3v, dx1f, dx2f, dx3f, dx1v, dx2v, dx3v;
  Real GetCellVolume(int k, int j, int i) const { return dx1f(i) * dx2f(j) * dx3f(k); }
};

class Hydro {   // hydro/hydro.hpp:32-122 (host mirrors only)
 public:
  explicit Hydro(MeshBlock *pmb);
  AthenaArray<Real> u, w;
  AthenaArray<Real> w1;   // (RELATIVISTIC_DYNAMICS only: what a pgen leaves here seeds the second W(U))
};
class Field {   // field/field.hpp:30-75
 public:
  explicit Field(MeshBlock *pmb);
  FaceField b;
  AthenaArray<Real> bcc;
};
class PassiveScalars {
 public:
  explicit PassiveScalars(MeshBlock *pmb);
  AthenaArray<Real> s, r;
};
class EquationOfState {   // eos/eos.hpp:32-181
 public:
  explicit EquationOfState(ParameterInput *pin);
  Real GetGamma() const { return gamma_; }
  Real GetIsoSoundSpeed() const { return iso_sound_speed_; }
  Real GetGammaMax() const { return gamma_max_; }
  Real GetDensityFloor() const { return dfloor_; }
  Real GetPressureFloor() const { return pfloor_; }
  Real GetScalarFloor() const { return sfloor_; }
 private:
  Real gamma_, dfloor_, pfloor_, sfloor_, iso_sound_speed_ = 0.0, gamma_max_ = 1000.0;
};

// bvals/bvals_interfaces.hpp:120-190 (NeighborBlock) -- what the refinement code needs of it
struct NeighborBlock {
  int gid, lid, level;   // global id; local id on the owning GPU's context
  int rank = 0;          // which GPU owns it
  int ox1, ox2, ox3, fi1, fi2;
  int type;   // 0 face, 1 edge, 2 corner
  int fid;    // BoundaryFace of a face neighbour
};

// mesh/meshblock_tree.hpp: octree of LogicalLocations (3-D), leaves = MeshBlocks
class MeshBlockTree {
 public:
  explicit MeshBlockTree(Mesh *pmesh);
  MeshBlockTree(MeshBlockTree *parent, int ox1, int ox2, int ox3);
  ~MeshBlockTree();
  void CreateRootGrid();
  void AddMeshBlock(LogicalLocation rloc);
  void Refine();
  void GetMeshBlockList(std::vector<LogicalLocation> *list);
  MeshBlockTree *FindNeighbor(LogicalLocation myloc, int ox1, int ox2, int ox3);
  MeshBlockTree *GetLeaf(int ox1, int ox2, int ox3) { return pleaf_[ox1 + (ox2 << 1) + (ox3 << 2)]; }
  LogicalLocation loc_;
  bool has_leaves_ = false;
  int gid_ = -1;
 private:
  Mesh *pmesh_;
  MeshBlockTree *proot_;
  MeshBlockTree *pleaf_[8];
};

class MeshBlock {   // mesh/mesh.hpp:75-200
 public:
  MeshBlock(int gid, int lid, LogicalLocation loc, RegionSize bs, Mesh *pm, ParameterInput *pin);
  ~MeshBlock();
  void ProblemGenerator(ParameterInput *pin);   // dispatches to the selected pgen
  Mesh *pmy_mesh;
  LogicalLocation loc;
  RegionSize block_size;
  int gid, lid;                 // global id; local id on its GPU's context
  int rank = 0;                 // which of the Mesh's GPUs owns the block (the reference's rank)
  int is, ie, js, je, ks, ke;
  int ncells1, ncells2, ncells3;
  Coordinates *pcoord;
  Hydro *phydro;
  Field *pfield;
  PassiveScalars *pscalars;
  EquationOfState *peos;
  Real new_block_dt_;
  // static mesh refinement (bvals_base.hpp): neighbours in SearchAndSetNeighbors' order, their levels
  std::vector<NeighborBlock> neighbor;
  int nblevel[3][3][3];
  // what a user-enrolled boundary function sees of the coarse buffer (built on first use)
  Coordinates *pcoarsec = nullptr;
  AthenaArray<Real> coarse_prim_;
  FaceField coarse_b_;
  BoundaryFlag block_bcs[6] = {BoundaryFlag::block, BoundaryFlag::block, BoundaryFlag::block,
                               BoundaryFlag::block, BoundaryFlag::block, BoundaryFlag::block};
};

// athena.hpp:98-100, 214-217: user-enrolled boundary functions act on the HOST mirrors of the
// primitives and of the face field (see Mesh::EnrollUserBoundaryFunction)
enum BoundaryFace { undef = -1, inner_x1 = 0, outer_x1 = 1, inner_x2 = 2, outer_x2 = 3, inner_x3 = 4, outer_x3 = 5 };
using BValFunc = void (*)(MeshBlock *pmb, Coordinates *pco, AthenaArray<Real> &prim, FaceField &b, Real time, Real dt,
                          int is, int ie, int js, int je, int ks, int ke, int ngh);

using SrcTermFunc = void (*)(MeshBlock *pmb, const Real time, const Real dt, const AthenaArray<Real> &prim,
                             const AthenaArray<Real> &prim_scalar, const AthenaArray<Real> &bcc,
                             AthenaArray<Real> &cons, AthenaArray<Real> &cons_scalar);   // athena.hpp:218-222

typedef void (*PgenFn)(MeshBlock *pmb, ParameterInput *pin);
typedef void (*MeshHookFn)(Mesh *pm, ParameterInput *pin);
struct ProblemGeneratorEntry {
  const char *name;
  MeshHookFn init_user_mesh_data;    // Mesh::InitUserMeshData
  PgenFn problem_generator;          // MeshBlock::ProblemGenerator
  MeshHookFn user_work_after_loop;   // Mesh::UserWorkAfterLoop
};
const ProblemGeneratorEntry *FindProblemGenerator(const std::string &name);

class Mesh {   // mesh/mesh.hpp:206-447
 public:
  Mesh(ParameterInput *pin);
  ~Mesh();
  void Initialize(ParameterInput *pin);   // mesh.cpp:1622
  // restart (mesh.cpp:621-960 + MeshBlock restart constructor, meshblock.cpp:263-420): raw = the
  // whole .rst file, offset = first byte after <par_end>
  void LoadRestart(const std::string &raw, std::size_t offset);
  bool res_flag = false;
  void NewTimeStep();                     // mesh.cpp:1236
  void UserWorkAfterLoop(ParameterInput *pin);
  void OutputCycleDiagnostics();          // mesh.cpp:2304
  void DownloadAll();                     // device -> host mirrors (outputs)
  void ExchangeGhosts();                  // SEND/RECV/SETB of every variable, all GPUs
  MeshBlock *my_blocks(int b) { return blocks_[b]; }
  RegionSize mesh_size;
  BoundaryFlag mesh_bcs[6];
  bool f2, f3;
  int ndim;
  Real start_time, time, tlim, dt, cfl_number;
  int nlim, ncycle, ncycle_out;
  int nbtotal, nblocal, root_level;
  int nrbx1, nrbx2, nrbx3;
  std::vector<LogicalLocation> loclist;
  // MeshBlocks are cut into contiguous Z-order ranges, one per GPU (amr_loadbalance.cpp:73-131:
  // ranklist / nslist / nblist), one device context per GPU; neighbours on another GPU are remote
  // links served by the peer-to-peer transport (include/athk.h).  <mesh> num_gpus, default 1.
  int ngpu = 1;
  std::vector<int> ranklist, nslist, nblist;
  std::vector<athk_ctx *> ctx;            // [ngpu]
  athk_ctx *ctx_of(const MeshBlock *pmb) { return ctx[pmb->rank]; }
  athk_ctx *athk;                         // == ctx[0]
  void SetupRemoteLinks();                // multigpu.cpp
  // static mesh refinement (<mesh> refinement = static, <refinementN> blocks): refinement.cpp
  bool multilevel = false;
  int current_level = 0;
  MeshBlockTree *tree = nullptr;
  void BuildRefinedTree(ParameterInput *pin);
  void SearchAndSetNeighbors(MeshBlock *pmb);
  void SetupRefinement();
  // The box lists (athk_refine_set_ops / athk_refine_run; include/athk.h), as the host numbers them:
  //    0  flux correction of the hydro fluxes from finer faces      66  ... of the scalar fluxes
  //    1  restriction for coarser neighbours (u, s, b)               2  cell-centred from finer / from coarser
  //    3  restriction of same-level ghost zones (u, s, b)            4  W(U) on the coarse buffer
  //    7  reflecting / outflow functions on the coarse buffer        5  prolongation of the primitives (w, r)
  //    6  U(W) (+ bcc, s) on the prolongated ghost zones
  //    8, 9, 10  field from same-level / finer face, edge, corner neighbours
  //   11, 64, 65 field from coarser face, edge, corner neighbours
  //   12  shared faces, 13 internal faces of the prolongated field
  //   14  clear, 16..63 add (one list per face / edge neighbour slot), 15 average: EMFs from finer blocks
  // Order of execution: Mesh::ExchangeGhosts (mesh.cpp), TimeIntegratorTaskList::DoTaskListOneStage.
  // face-centred field + EMF correction boxes; lists[rank][list]
  void SetupFieldRefinement(std::vector<std::vector<std::vector<athk_box_op>>> *lists);
  // user-enrolled boundary functions on the coarse buffer (bvals_refine.cpp:445-530): which function, on
  // which block, over which coarse index range (DispatchBoundaryFunctions' arguments, ngh = 1)
  struct CoarseUserBC { int gid, face, il, iu, jl, ju, kl, ku; };
  std::vector<CoarseUserBC> coarse_user_bcs_;
  Real bc_time_ = 0.0, bc_dt_ = 0.0;   // as handed to athk_set_boundary_time
  void ApplyUserBoundariesOnCoarseLevel();
  void RunRefineList(int list);   // one box list on every GPU, ordered after everything enqueued so far on all of them
  void SyncAll();
  const ProblemGeneratorEntry *pgen;
  bool fused_stage;
  // user history output (mesh.hpp: EnrollUserHistoryOutput)
  typedef Real (*HistoryFn)(MeshBlock *pmb, int iout);
  struct UserHistory { HistoryFn fn; std::string name; int op; /*0 sum,1 max,2 min*/ };
  std::vector<UserHistory> user_history;
  void EnrollUserHistoryOutput(HistoryFn fn, const char *name, int op) { user_history.push_back({fn, name, op}); }
  // Mesh::EnrollUserBoundaryFunction (mesh.cpp:1288-1309).  The function runs on the host, like the
  // reference's: at PHY_BVAL the block's primitives (and face field) are brought to the host
  // mirrors, handed to it, and its ghost zones sent back (a CUDA boundary function can instead be
  // enrolled with athk_enroll_user_boundary and stays on the device).
  void EnrollUserBoundaryFunction(BoundaryFace face, BValFunc my_func);
  BValFunc BoundaryFunction_[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  // Mesh::EnrollUserExplicitSourceFunction (mesh.cpp:1345): host function on the host mirrors, as above
  void EnrollUserExplicitSourceFunction(SrcTermFunc my_func);
  SrcTermFunc UserSourceTerm_ = nullptr;
  MeshBlock *FindLocalBlock(int rank, int lid);

 private:
  std::vector<MeshBlock *> blocks_;
  std::vector<void *> p2p_allocs_;        // per GPU: ghost, emf, flag buffers (athk_p2p_alloc)
  int gpu_device0_ = 0;                   // <mesh> gpu_device: CUDA device of GPU 0 of this Mesh
};

void AthkCheck(int rc, const char *where);   // non-zero return -> ATHENA_ERROR(athk_last_error)

// TimeIntegratorTaskList (task_list/time_integrator.cpp:42-1206)
class TimeIntegratorTaskList {
 public:
  TimeIntegratorTaskList(ParameterInput *pin, Mesh *pm);
  void DoTaskListOneStage(Mesh *pm, int stage);   // task_list.cpp:66
  std::string integrator;
  int nstages;
  Real cfl_limit;
  struct IntegratorWeight { Real delta, gamma_1, gamma_2, gamma_3, beta, ebeta, sbeta; };
  IntegratorWeight stage_wghts[6];
  int xorder;
};

// Outputs (outputs/outputs.hpp): history (.hst) and restart (.rst) writers
class Outputs {
 public:
  Outputs(Mesh *pm, ParameterInput *pin);
  void MakeOutputs(Mesh *pm, ParameterInput *pin, bool wtflag = false);   // outputs.cpp:1490
 private:
  struct OutputParameters {
    std::string block_name, file_type, file_basename, data_format, file_id, variable;
    Real dt, next_time;
    int dcycle, file_number;
    bool include_ghost_zones = false;
  };
  // one column group of a .tab / .vtk file: nvar components of a MeshBlock array from n0 on
  struct OutputField {
    std::string type, name;   // "SCALARS" | "VECTORS", dataset name
    const AthenaArray<Real> *src;
    int n0, nvar;
  };
  std::vector<OutputParameters> types_;
  void WriteHistory(OutputParameters &op, Mesh *pm, ParameterInput *pin);
  void WriteRestart(OutputParameters &op, Mesh *pm, ParameterInput *pin, bool force);
  void WriteTable(OutputParameters &op, Mesh *pm, ParameterInput *pin);
  void WriteVTK(OutputParameters &op, Mesh *pm, ParameterInput *pin);
  static std::vector<OutputField> SelectFields(MeshBlock *pmb, const std::string &variable);
};

#endif  // ATHENA_B200_HOST_ATHENA_HPP_
