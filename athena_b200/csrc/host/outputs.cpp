// outputs.cpp -- history (.hst), restart (.rst), formatted table (.tab) and VTK (.vtk) writers in
// the reference's formats, so the same readers (vis/python/athena_read.py, oracle/refrun.py,
// VisIt / ParaView) parse both codes' artefacts.
//   scheduling      src/outputs/outputs.cpp:1490-1522 (MakeOutputs), :209-330 (<outputN> parsing)
//   history sums    src/outputs/history.cpp:196-248 (serial block-by-block, k,j,i order), header
//                   and row format :376-447
//   restart layout  src/outputs/restart.cpp:42-246
//   variable lists  src/outputs/outputs.cpp:480-760,1310-1390 (LoadOutputData), :1798 (ContainVariable)
//   .tab / .vtk     src/outputs/formatted_table.cpp:36-148, src/outputs/vtk.cpp:55-221 (one file per
//                   MeshBlock; slices and sums are not implemented and are refused)
// Data are downloaded from the device at output times only (Mesh::DownloadAll).
#include <algorithm>
#include <cstdio>
#include <fstream>
#include <limits>

#include "athena.hpp"

Outputs::Outputs(Mesh *pm, ParameterInput *pin) {
  const std::string basename = pin->GetString("job", "problem_id");
  // copy the block list first: GetOrAdd* below may append to blocks
  std::vector<std::string> names;
  for (const ParameterInput::Block &b : pin->blocks())
    if (b.name.compare(0, 6, "output") == 0) names.push_back(b.name);
  for (const std::string &bn : names) {
    OutputParameters op;
    op.block_name = bn;
    // the reference's order (outputs.cpp:206-223): next_time first, so that a disabled block carries
    // it into every restart file too; then the two consistency checks
    op.next_time = pin->GetOrAddReal(bn, "next_time", pm->time);
    op.dt = pin->GetOrAddReal(bn, "dt", 0.0);
    op.dcycle = pin->GetOrAddInteger(bn, "dcycle", 0);
    if (op.dt == 0.0 && op.dcycle == 0) {
      std::stringstream msg;
      msg << "### FATAL ERROR in Outputs constructor" << std::endl
          << "Either dt or dcycle must be specified in " << bn << std::endl;
      ATHENA_ERROR(msg);
    }
    if (op.dt > 0.0 && op.dcycle > 0) {
      std::stringstream msg;
      msg << "### FATAL ERROR in Outputs constructor" << std::endl
          << "dt and dcycle cannot be specified simultaneously in " << bn << std::endl;
      ATHENA_ERROR(msg);
    }
    if (!(op.dt > 0.0 || op.dcycle > 0)) continue;   // disabled block (e.g. dt = -1)
    op.file_number = pin->GetOrAddInteger(bn, "file_number", 0);
    op.file_basename = basename;
    op.file_id = pin->GetOrAddString(bn, "id", "out" + bn.substr(6));   // default id = "outN"
    op.file_type = pin->GetString(bn, "file_type");
    if (op.file_type != "hst" && op.file_type != "rst" && op.file_type != "tab" && op.file_type != "vtk") {
      std::cout << "### Warning in Outputs constructor" << std::endl
                << "output type '" << op.file_type << "' of block <" << bn
                << "> is not implemented by the B200 host (hst, rst, tab and vtk are); skipped" << std::endl;
      continue;
    }
    for (const char *key : {"x1_slice", "x2_slice", "x3_slice"})
      if (pin->DoesParameterExist(bn, key)) {
        std::stringstream msg;
        msg << "### FATAL ERROR in Outputs constructor" << std::endl
            << "'" << key << "' in block <" << bn << ">: sliced outputs are not implemented by the B200 host"
            << std::endl;
        ATHENA_ERROR(msg);
      }
    for (const char *key : {"x1_sum", "x2_sum", "x3_sum"})
      if (pin->GetOrAddBoolean(bn, key, false)) {
        std::stringstream msg;
        msg << "### FATAL ERROR in Outputs constructor" << std::endl
            << "'" << key << "' in block <" << bn << ">: summed outputs are not implemented by the B200 host"
            << std::endl;
        ATHENA_ERROR(msg);
      }
    op.include_ghost_zones = pin->GetOrAddBoolean(bn, "ghost_zones", false);
    pin->GetOrAddBoolean(bn, "orbital_system", false);
    if (op.file_type == "tab" || op.file_type == "vtk") op.variable = pin->GetString(bn, "variable");
    if (op.file_type == "hst" || op.file_type == "tab") {
      op.data_format = pin->GetOrAddString(bn, "data_format", "%12.5e");
      op.data_format.insert(0, " ");
    }
    types_.push_back(op);
  }
  // the reference moves restart outputs to the tail so that they see updated parameters
  std::stable_sort(types_.begin(), types_.end(), [](const OutputParameters &a, const OutputParameters &b) {
    return (a.file_type == "rst") < (b.file_type == "rst");
  });
}

void Outputs::MakeOutputs(Mesh *pm, ParameterInput *pin, bool wtflag) {
  bool downloaded = false;
  for (OutputParameters &op : types_) {
    if (((pm->time == pm->start_time) && (op.next_time <= pm->start_time)) ||
        (op.dt > 0.0 && pm->time >= op.next_time) ||
        (op.dcycle > 0 && pm->ncycle % op.dcycle == 0) || (pm->time >= pm->tlim) ||
        (wtflag && op.file_type == "rst")) {
      if (!downloaded) { pm->DownloadAll(); downloaded = true; }
      if (op.file_type == "hst") WriteHistory(op, pm, pin);
      else if (op.file_type == "tab") WriteTable(op, pm, pin);
      else if (op.file_type == "vtk") WriteVTK(op, pm, pin);
      else WriteRestart(op, pm, pin, wtflag);
    }
  }
}

void Outputs::WriteHistory(OutputParameters &op, Mesh *pm, ParameterInput *pin) {
  const bool mhd = MAGNETIC_FIELDS_ENABLED;
  const int nvars = NHYDRO + 3 + (mhd ? 3 : 0) + NSCALARS;
  const int nuser = static_cast<int>(pm->user_history.size());
  std::vector<Real> h(nvars + nuser, 0.0);
  for (int n = 0; n < nuser; ++n) {
    if (pm->user_history[n].op == 1) h[nvars + n] = std::numeric_limits<Real>::lowest();
    if (pm->user_history[n].op == 2) h[nvars + n] = std::numeric_limits<Real>::max();
  }
  for (int b = 0; b < pm->nblocal; ++b) {
    MeshBlock *pmb = pm->my_blocks(b);
    const AthenaArray<Real> &u = pmb->phydro->u;
    for (int k = pmb->ks; k <= pmb->ke; ++k)
      for (int j = pmb->js; j <= pmb->je; ++j)
        for (int i = pmb->is; i <= pmb->ie; ++i) {
          const Real vol = pmb->pcoord->GetCellVolume(k, j, i);
          const Real &u_d = u(IDN, k, j, i), &u_mx = u(IM1, k, j, i), &u_my = u(IM2, k, j, i),
                     &u_mz = u(IM3, k, j, i);
          h[0] += vol * u_d;
          h[1] += vol * u_mx;
          h[2] += vol * u_my;
          h[3] += vol * u_mz;
          h[4] += vol * 0.5 * SQR(u_mx) / u_d;
          h[5] += vol * 0.5 * SQR(u_my) / u_d;
          h[6] += vol * 0.5 * SQR(u_mz) / u_d;
          if (NON_BAROTROPIC_EOS) h[7] += vol * u(IEN, k, j, i);
          if (mhd) {
            const AthenaArray<Real> &bcc = pmb->pfield->bcc;
            const Real &b1 = bcc(IB1, k, j, i), &b2 = bcc(IB2, k, j, i), &b3 = bcc(IB3, k, j, i);
            h[NHYDRO + 3] += vol * 0.5 * b1 * b1;
            h[NHYDRO + 4] += vol * 0.5 * b2 * b2;
            h[NHYDRO + 5] += vol * 0.5 * b3 * b3;
          }
          for (int n = 0; n < NSCALARS; ++n)
            h[NHYDRO + 3 + (mhd ? 3 : 0) + n] += vol * pmb->pscalars->s(n, k, j, i);
        }
    for (int n = 0; n < nuser; ++n) {
      const Real v = pm->user_history[n].fn(pmb, n);
      Real &acc = h[nvars + n];
      if (pm->user_history[n].op == 0) acc += v;
      else if (pm->user_history[n].op == 1) acc = std::max(acc, v);
      else acc = std::min(acc, v);
    }
  }
  const std::string fname = op.file_basename + ".hst";
  FILE *pfile = std::fopen(fname.c_str(), "a");   // always appended to, as history.cpp:367
  if (pfile == nullptr) {
    std::stringstream msg;
    msg << "### FATAL ERROR in function [OutputType::HistoryFile]" << std::endl
        << "Output file '" << fname << "' could not be opened";
    ATHENA_ERROR(msg);
  }
  if (op.file_number == 0) {
    int iout = 1;
    std::fprintf(pfile, "# Athena++ history data\n");
    std::fprintf(pfile, "# [%d]=time     ", iout++);
    std::fprintf(pfile, "[%d]=dt       ", iout++);
    std::fprintf(pfile, "[%d]=mass     ", iout++);
    std::fprintf(pfile, "[%d]=1-mom    ", iout++);
    std::fprintf(pfile, "[%d]=2-mom    ", iout++);
    std::fprintf(pfile, "[%d]=3-mom    ", iout++);
    std::fprintf(pfile, "[%d]=1-KE     ", iout++);
    std::fprintf(pfile, "[%d]=2-KE     ", iout++);
    std::fprintf(pfile, "[%d]=3-KE     ", iout++);
    if (NON_BAROTROPIC_EOS) std::fprintf(pfile, "[%d]=tot-E   ", iout++);
    if (mhd) {
      std::fprintf(pfile, "[%d]=1-ME    ", iout++);
      std::fprintf(pfile, "[%d]=2-ME    ", iout++);
      std::fprintf(pfile, "[%d]=3-ME    ", iout++);
    }
    for (int n = 0; n < NSCALARS; ++n) std::fprintf(pfile, "[%d]=%d-scalar    ", iout++, n);
    for (int n = 0; n < nuser; ++n) std::fprintf(pfile, "[%d]=%-7s ", iout++, pm->user_history[n].name.c_str());
    std::fprintf(pfile, "\n");
  }
  std::fprintf(pfile, op.data_format.c_str(), pm->time);
  std::fprintf(pfile, op.data_format.c_str(), pm->dt);
  for (std::size_t n = 0; n < h.size(); ++n) std::fprintf(pfile, op.data_format.c_str(), h[n]);
  std::fprintf(pfile, "\n");
  std::fclose(pfile);
  op.file_number++;
  op.next_time += op.dt;
  pin->SetInteger(op.block_name, "file_number", op.file_number);
  pin->SetReal(op.block_name, "next_time", op.next_time);
}

void Outputs::WriteRestart(OutputParameters &op, Mesh *pm, ParameterInput *pin, bool force) {
  char number[8];
  std::snprintf(number, sizeof(number), "%05d", op.file_number);
  std::string fname = op.file_basename + "." + (force ? std::string("final") : std::string(number)) + ".rst";
  if (!force) {
    op.file_number++;
    op.next_time += op.dt;
    pin->SetInteger(op.block_name, "file_number", op.file_number);
    pin->SetReal(op.block_name, "next_time", op.next_time);
  }
  std::ofstream f(fname.c_str(), std::ios::binary);
  if (!f.good()) {
    std::stringstream msg;
    msg << "### FATAL ERROR in function [RestartOutput::WriteOutputFile]" << std::endl
        << "Output file '" << fname << "' could not be opened";
    ATHENA_ERROR(msg);
  }
  std::stringstream ost;
  pin->ParameterDump(ost);
  const std::string sbuf = ost.str();
  f.write(sbuf.data(), sbuf.size());
  MeshBlock *p0 = pm->my_blocks(0);
  std::uint64_t datasize = p0->phydro->u.GetSizeInBytes();
  if (MAGNETIC_FIELDS_ENABLED)
    datasize += p0->pfield->b.x1f.GetSizeInBytes() + p0->pfield->b.x2f.GetSizeInBytes() +
                p0->pfield->b.x3f.GetSizeInBytes();
  if (NSCALARS > 0) datasize += p0->pscalars->s.GetSizeInBytes();
  static_assert(sizeof(RegionSize) == 88, "RegionSize must match the reference's raw layout");
  static_assert(sizeof(LogicalLocation) == 32, "LogicalLocation must match the reference's raw layout");
  f.write(reinterpret_cast<const char *>(&pm->nbtotal), sizeof(int));
  f.write(reinterpret_cast<const char *>(&pm->root_level), sizeof(int));
  f.write(reinterpret_cast<const char *>(&pm->mesh_size), sizeof(RegionSize));
  f.write(reinterpret_cast<const char *>(&pm->time), sizeof(Real));
  f.write(reinterpret_cast<const char *>(&pm->dt), sizeof(Real));
  f.write(reinterpret_cast<const char *>(&pm->ncycle), sizeof(int));
  f.write(reinterpret_cast<const char *>(&datasize), sizeof(std::uint64_t));
  for (int b = 0; b < pm->nblocal; ++b) {
    const double cost = 1.0;
    f.write(reinterpret_cast<const char *>(&pm->my_blocks(b)->loc), sizeof(LogicalLocation));
    f.write(reinterpret_cast<const char *>(&cost), sizeof(double));
  }
  for (int b = 0; b < pm->nblocal; ++b) {
    MeshBlock *pmb = pm->my_blocks(b);
    f.write(reinterpret_cast<const char *>(pmb->phydro->u.data()), pmb->phydro->u.GetSizeInBytes());
    if (MAGNETIC_FIELDS_ENABLED) {
      f.write(reinterpret_cast<const char *>(pmb->pfield->b.x1f.data()), pmb->pfield->b.x1f.GetSizeInBytes());
      f.write(reinterpret_cast<const char *>(pmb->pfield->b.x2f.data()), pmb->pfield->b.x2f.GetSizeInBytes());
      f.write(reinterpret_cast<const char *>(pmb->pfield->b.x3f.data()), pmb->pfield->b.x3f.GetSizeInBytes());
    }
    if (NSCALARS > 0)
      f.write(reinterpret_cast<const char *>(pmb->pscalars->s.data()), pmb->pscalars->s.GetSizeInBytes());
  }
}

//----------------------------------------------------------------------------------------
// variable selection shared by .tab and .vtk

static bool ListHas(const std::string &list, const std::string &word) {   // comma-separated, exact words
  std::size_t a = 0;
  while (a <= list.size()) {
    std::size_t b = list.find(',', a);
    if (b == std::string::npos) b = list.size();
    if (list.compare(a, b - a, word) == 0) return true;
    a = b + 1;
  }
  return false;
}

std::vector<Outputs::OutputField> Outputs::SelectFields(MeshBlock *pmb, const std::string &var) {
  std::vector<OutputField> f;
  const bool cons = ListHas(var, "cons"), prim = ListHas(var, "prim");
  const AthenaArray<Real> &u = pmb->phydro->u, &w = pmb->phydro->w;
  auto scalar = [&](bool on, const char *name, const AthenaArray<Real> &a, int n) {
    if (on) f.push_back({"SCALARS", name, &a, n, 1});
  };
  auto vector = [&](bool on, const char *name, const AthenaArray<Real> &a, int n) {
    if (on) f.push_back({"VECTORS", name, &a, n, 3});
  };
  // the order below is the order of the columns / datasets in the reference's files
  scalar(ListHas(var, "D") || cons, "dens", u, IDN);
  scalar(ListHas(var, "d") || prim, "rho", w, IDN);
  if (NON_BAROTROPIC_EOS) {   // outputs.cpp:521-548
    scalar(ListHas(var, "E") || cons, "Etot", u, IEN);
    scalar(ListHas(var, "p") || prim, "press", w, IPR);
  }
  vector(ListHas(var, "m") || cons, "mom", u, IM1);
  scalar(ListHas(var, "m1"), "mom1", u, IM1);
  scalar(ListHas(var, "m2"), "mom2", u, IM2);
  scalar(ListHas(var, "m3"), "mom3", u, IM3);
  vector(ListHas(var, "v") || prim, "vel", w, IVX);
  scalar(ListHas(var, "vx") || ListHas(var, "v1"), "vel1", w, IVX);
  scalar(ListHas(var, "vy") || ListHas(var, "v2"), "vel2", w, IVY);
  scalar(ListHas(var, "vz") || ListHas(var, "v3"), "vel3", w, IVZ);
  for (int n = 0; n < NSCALARS; ++n) {
    const std::string sn = "s" + std::to_string(n), rn = "r" + std::to_string(n);
    scalar(ListHas(var, sn) || cons, sn.c_str(), pmb->pscalars->s, n);
    scalar(ListHas(var, rn) || prim, rn.c_str(), pmb->pscalars->r, n);
  }
  if (MAGNETIC_FIELDS_ENABLED) {
    const AthenaArray<Real> &bcc = pmb->pfield->bcc;
    vector(ListHas(var, "bcc") || prim || cons, "Bcc", bcc, IB1);
    scalar(ListHas(var, "bcc1"), "Bcc1", bcc, IB1);
    scalar(ListHas(var, "bcc2"), "Bcc2", bcc, IB2);
    scalar(ListHas(var, "bcc3"), "Bcc3", bcc, IB3);
    // face fields are listed by the index of the cell whose lower face they sit on
    const bool b = ListHas(var, "b");
    scalar(ListHas(var, "b1") || b, "B1", pmb->pfield->b.x1f, 0);
    scalar(ListHas(var, "b2") || b, "B2", pmb->pfield->b.x2f, 0);
    scalar(ListHas(var, "b3") || b, "B3", pmb->pfield->b.x3f, 0);
  }
  if (f.empty()) {
    std::stringstream msg;
    msg << "### FATAL ERROR in function [OutputType::LoadOutputData]" << std::endl
        << "Output variable '" << var << "' not implemented" << std::endl
        << "Allowed hydro variables: D,d,E,p,m,m1,m2,m3,v,v1,v2,v3,cons,prim (+ sN,rN, bcc,bcc1-3, b,b1-3)"
        << std::endl;
    ATHENA_ERROR(msg);
  }
  return f;
}

namespace {
struct OutputRange { int is, ie, js, je, ks, ke; };
OutputRange RangeOf(const MeshBlock *pmb, bool ghosts) {
  OutputRange r = {pmb->is, pmb->ie, pmb->js, pmb->je, pmb->ks, pmb->ke};
  if (ghosts) {
    r.is -= NGHOST, r.ie += NGHOST;
    if (r.js != r.je) r.js -= NGHOST, r.je += NGHOST;
    if (r.ks != r.ke) r.ks -= NGHOST, r.ke += NGHOST;
  }
  return r;
}
std::string BlockFileName(const std::string &base, int gid, const std::string &id, int number, const char *ext) {
  char buf[32];
  std::snprintf(buf, sizeof(buf), ".block%d.", gid);
  std::string fname = base + buf + id;
  std::snprintf(buf, sizeof(buf), ".%05d.", number);
  return fname + buf + ext;
}
FILE *OpenOrThrow(const std::string &fname, const char *who) {
  FILE *pfile = std::fopen(fname.c_str(), "w");
  if (pfile == nullptr) {
    std::stringstream msg;
    msg << "### FATAL ERROR in function [" << who << "::WriteOutputFile]" << std::endl
        << "Output file '" << fname << "' could not be opened" << std::endl;
    ATHENA_ERROR(msg);
  }
  return pfile;
}
// VTK legacy binary is big-endian
void WriteFloatsBE(FILE *pfile, std::vector<float> &v, std::size_t n) {
  unsigned char *p = reinterpret_cast<unsigned char *>(v.data());
  const std::uint16_t probe = 1;
  if (*reinterpret_cast<const unsigned char *>(&probe) == 1)
    for (std::size_t i = 0; i < n; ++i) {
      std::swap(p[4 * i], p[4 * i + 3]);
      std::swap(p[4 * i + 1], p[4 * i + 2]);
    }
  std::fwrite(v.data(), sizeof(float), n, pfile);
}
}  // namespace

void Outputs::WriteTable(OutputParameters &op, Mesh *pm, ParameterInput *pin) {
  const char *fmt = op.data_format.c_str();
  for (int b = 0; b < pm->nblocal; ++b) {
    MeshBlock *pmb = pm->my_blocks(b);
    const OutputRange r = RangeOf(pmb, op.include_ghost_zones);
    const std::vector<OutputField> fields = SelectFields(pmb, op.variable);
    FILE *pfile = OpenOrThrow(BlockFileName(op.file_basename, pmb->gid, op.file_id, op.file_number, "tab"),
                              "FormattedTableOutput");
    std::fprintf(pfile, "# Athena++ data at time=%e", pm->time);
    std::fprintf(pfile, "  cycle=%d", pm->ncycle);
    std::fprintf(pfile, "  variables=%s \n", op.variable.c_str());
    std::fprintf(pfile, "#");
    if (r.is != r.ie) std::fprintf(pfile, " i       x1v     ");
    if (r.js != r.je) std::fprintf(pfile, " j       x2v     ");
    if (r.ks != r.ke) std::fprintf(pfile, " k       x3v     ");
    for (const OutputField &f : fields) {
      if (f.type == "VECTORS")
        for (int index = 1; index <= 3; ++index) std::fprintf(pfile, "    %s%d     ", f.name.c_str(), index);
      else
        std::fprintf(pfile, "    %s      ", f.name.c_str());
    }
    std::fprintf(pfile, "\n");
    for (int k = r.ks; k <= r.ke; ++k)
      for (int j = r.js; j <= r.je; ++j)
        for (int i = r.is; i <= r.ie; ++i) {
          if (r.is != r.ie) {
            std::fprintf(pfile, "%04d", i);
            std::fprintf(pfile, fmt, pmb->pcoord->x1v(i));
          }
          if (r.js != r.je) {
            std::fprintf(pfile, " %04d", j);
            std::fprintf(pfile, fmt, pmb->pcoord->x2v(j));
          }
          if (r.ks != r.ke) {
            std::fprintf(pfile, " %04d", k);
            std::fprintf(pfile, fmt, pmb->pcoord->x3v(k));
          }
          for (const OutputField &f : fields)
            for (int n = 0; n < f.nvar; ++n) std::fprintf(pfile, fmt, (*f.src)(f.n0 + n, k, j, i));
          std::fprintf(pfile, "\n");
        }
    std::fclose(pfile);
  }
  op.file_number++;
  op.next_time += op.dt;
  pin->SetInteger(op.block_name, "file_number", op.file_number);
  pin->SetReal(op.block_name, "next_time", op.next_time);
}

void Outputs::WriteVTK(OutputParameters &op, Mesh *pm, ParameterInput *pin) {
  for (int b = 0; b < pm->nblocal; ++b) {
    MeshBlock *pmb = pm->my_blocks(b);
    Coordinates *pc = pmb->pcoord;
    const OutputRange r = RangeOf(pmb, op.include_ghost_zones);
    const std::vector<OutputField> fields = SelectFields(pmb, op.variable);
    FILE *pfile = OpenOrThrow(BlockFileName(op.file_basename, pmb->gid, op.file_id, op.file_number, "vtk"),
                              "VTKOutput");
    std::fprintf(pfile, "# vtk DataFile Version 2.0\n");
    std::fprintf(pfile, "# Athena++ data at time=%e", pm->time);
    std::fprintf(pfile, "  cycle=%d", pm->ncycle);
    std::fprintf(pfile, "  variables=%s \n", op.variable.c_str());
    std::fprintf(pfile, "BINARY\n");
    const int nc1 = r.ie - r.is + 1, nc2 = r.je - r.js + 1, nc3 = r.ke - r.ks + 1;
    // cell data on a rectilinear grid: one more coordinate than cells along every extended axis
    const int np1 = nc1 + (nc1 > 1), np2 = nc2 + (nc2 > 1), np3 = nc3 + (nc3 > 1);
    std::vector<float> data(3 * static_cast<std::size_t>(std::max(np1, std::max(np2, np3))));
    std::fprintf(pfile, "DATASET RECTILINEAR_GRID\n");
    std::fprintf(pfile, "DIMENSIONS %d %d %d\n", np1, np2, np3);
    std::fprintf(pfile, "X_COORDINATES %d float\n", np1);
    if (nc1 == 1) data[0] = static_cast<float>(pc->x1v(r.is));
    else for (int i = r.is; i <= r.ie + 1; ++i) data[i - r.is] = static_cast<float>(pc->x1f(i));
    WriteFloatsBE(pfile, data, np1);
    std::fprintf(pfile, "\nY_COORDINATES %d float\n", np2);
    if (nc2 == 1) data[0] = static_cast<float>(pc->x2v(r.js));
    else for (int j = r.js; j <= r.je + 1; ++j) data[j - r.js] = static_cast<float>(pc->x2f(j));
    WriteFloatsBE(pfile, data, np2);
    std::fprintf(pfile, "\nZ_COORDINATES %d float\n", np3);
    if (nc3 == 1) data[0] = static_cast<float>(pc->x3v(r.ks));
    else for (int k = r.ks; k <= r.ke + 1; ++k) data[k - r.ks] = static_cast<float>(pc->x3f(k));
    WriteFloatsBE(pfile, data, np3);
    std::fprintf(pfile, "\nCELL_DATA %d", nc1 * nc2 * nc3);
    for (const OutputField &f : fields) {
      std::fprintf(pfile, "\n%s %s float\n", f.type.c_str(), f.name.c_str());
      if (f.nvar == 1) std::fprintf(pfile, "LOOKUP_TABLE default\n");
      for (int k = r.ks; k <= r.ke; ++k)
        for (int j = r.js; j <= r.je; ++j) {
          for (int i = r.is; i <= r.ie; ++i)
            for (int n = 0; n < f.nvar; ++n)
              data[f.nvar * (i - r.is) + n] = static_cast<float>((*f.src)(f.n0 + n, k, j, i));
          WriteFloatsBE(pfile, data, static_cast<std::size_t>(f.nvar) * nc1);
        }
    }
    std::fclose(pfile);
  }
  op.file_number++;
  op.next_time += op.dt;
  pin->SetInteger(op.block_name, "file_number", op.file_number);
  pin->SetReal(op.block_name, "next_time", op.next_time);
}
