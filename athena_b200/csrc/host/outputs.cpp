// outputs.cpp -- history (.hst) and restart (.rst) writers in the reference's formats, so the
// same readers (vis/python/athena_read.py, oracle/refrun.py) parse both codes' artefacts.
//   scheduling      src/outputs/outputs.cpp:1490-1522 (MakeOutputs), :209-330 (<outputN> parsing)
//   history sums    src/outputs/history.cpp:196-248 (serial block-by-block, k,j,i order), header
//                   and row format :376-447
//   restart layout  src/outputs/restart.cpp:42-246
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
    op.dt = pin->GetOrAddReal(bn, "dt", 0.0);
    op.dcycle = pin->GetOrAddInteger(bn, "dcycle", 0);
    if (op.dt <= 0.0 && op.dcycle <= 0) continue;   // outputs.cpp:209-223: disabled block
    op.file_type = pin->GetString(bn, "file_type");
    op.file_basename = basename;
    op.next_time = pin->GetOrAddReal(bn, "next_time", pm->time);
    op.file_number = pin->GetOrAddInteger(bn, "file_number", 0);
    op.data_format = pin->GetOrAddString(bn, "data_format", "%12.5e");
    op.data_format.insert(0, " ");
    if (op.file_type != "hst" && op.file_type != "rst") {
      std::cout << "### Warning in Outputs constructor" << std::endl
                << "output type '" << op.file_type << "' of block <" << bn
                << "> is not implemented by the B200 host (hst and rst are); skipped" << std::endl;
      continue;
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
          h[7] += vol * u(IEN, k, j, i);
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
  FILE *pfile = std::fopen(fname.c_str(), op.file_number == 0 ? "w" : "a");
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
    std::fprintf(pfile, "[%d]=tot-E   ", iout++);
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
