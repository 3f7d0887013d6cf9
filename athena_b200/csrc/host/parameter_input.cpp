// parameter_input.cpp -- athinput parser with the reference's grammar and error behaviour
// (src/parameter_input.cpp:120-186 LoadFromStream, :265-305 ParseLine, :352-395
// ModifyFromCmdline, :422-573 Get*, :575-688 GetOrAdd*, :835-876 ParameterDump).
#include <algorithm>
#include <cstdlib>
#include <fstream>

#include <cmath>

#include "athena.hpp"

static std::string trim(const std::string &s) {
  const std::size_t a = s.find_first_not_of(' ');
  if (a == std::string::npos) return "";
  const std::size_t b = s.find_last_not_of(' ');
  return s.substr(a, b - a + 1);
}

void ParameterInput::LoadFromFile(const std::string &fname) {
  std::ifstream f(fname.c_str());
  if (!f.good()) {
    std::stringstream msg;
    msg << "### FATAL ERROR in function [ParameterInput::LoadFromFile]" << std::endl
        << "Input file '" << fname << "' could not be opened";
    ATHENA_ERROR(msg);
  }
  LoadFromStream(f);
}

void ParameterInput::LoadFromStream(std::istream &is) {
  std::string line;
  Block *pb = nullptr;
  while (is.good()) {
    std::getline(is, line);
    line.erase(std::remove(line.begin(), line.end(), '\t'), line.end());
    const std::size_t first = line.find_first_not_of(' ');
    if (line.empty() || first == std::string::npos) continue;
    if (line[first] == '#') continue;
    if (line.compare(first, 9, "<par_end>") == 0) break;
    if (line[first] == '<') {
      const std::size_t last = line.find_first_of('>', first + 1);
      if (last == std::string::npos) {
        std::stringstream msg;
        msg << "### FATAL ERROR in function [ParameterInput::LoadFromStream]" << std::endl
            << "Block name '" << line.substr(first + 1) << "' in the input stream' not properly ended";
        ATHENA_ERROR(msg);
      }
      pb = FindOrAdd(line.substr(first + 1, last - first - 1));
      continue;
    }
    if (pb == nullptr) {
      std::stringstream msg;
      msg << "### FATAL ERROR in function [ParameterInput::LoadFromStream]" << std::endl
          << "Input file must specify a block name before the first parameter = value line";
      ATHENA_ERROR(msg);
    }
    const std::size_t eq = line.find_first_of('=');
    const std::size_t hash = line.find_first_of('#');
    if (eq == std::string::npos) continue;
    const std::string name = trim(line.substr(first, eq - first));
    const std::string value = trim(line.substr(eq + 1, hash == std::string::npos ? std::string::npos : hash - eq - 1));
    const std::string comment = hash == std::string::npos ? "" : line.substr(hash);
    Add(pb, name, value, comment);
  }
}

const ParameterInput::Block *ParameterInput::Find(const std::string &b) const {
  for (const Block &x : blocks_)
    if (x.name == b) return &x;
  return nullptr;
}

ParameterInput::Block *ParameterInput::FindOrAdd(const std::string &b) {
  for (Block &x : blocks_)
    if (x.name == b) return &x;
  blocks_.push_back(Block{b, {}});
  return &blocks_.back();
}

void ParameterInput::Add(Block *pb, const std::string &n, const std::string &v, const std::string &c) {
  for (Line &l : pb->lines)
    if (l.name == n) { l.value = v; l.comment = c; return; }
  pb->lines.push_back(Line{n, v, c});
}

void ParameterInput::ModifyFromCmdline(int argc, char *argv[]) {
  for (int i = 1; i < argc; ++i) {
    const std::string t = argv[i];
    const std::size_t slash = t.find_first_of('/'), eq = t.find_first_of('=');
    if (slash == std::string::npos || eq == std::string::npos || t.compare(0, 2, "--") == 0) continue;
    const std::string block = t.substr(0, slash), name = t.substr(slash + 1, eq - slash - 1),
                      value = t.substr(eq + 1);
    Block *pb = nullptr;
    for (Block &x : blocks_)
      if (x.name == block) pb = &x;
    std::stringstream msg;
    if (pb == nullptr) {
      msg << "### FATAL ERROR in function [ParameterInput::ModifyFromCmdline]" << std::endl
          << "Block name '" << block << "' on command line not found";
      ATHENA_ERROR(msg);
    }
    Line *pl = nullptr;
    for (Line &l : pb->lines)
      if (l.name == name) pl = &l;
    if (pl == nullptr) {   // only EXISTING parameters may be overridden
      msg << "### FATAL ERROR in function [ParameterInput::ModifyFromCmdline]" << std::endl
          << "Parameter '" << name << "' in block '" << block << "' on command line not found";
      ATHENA_ERROR(msg);
    }
    pl->value = value;
  }
}

int ParameterInput::DoesParameterExist(const std::string &block, const std::string &name) const {
  const Block *pb = Find(block);
  if (pb == nullptr) return 0;
  for (const Line &l : pb->lines)
    if (l.name == name) return 1;
  return 0;
}

const ParameterInput::Line *ParameterInput::FindLine(const std::string &b, const std::string &n,
                                                     const char *who) const {
  std::stringstream msg;
  const Block *pb = Find(b);
  if (pb == nullptr) {
    msg << "### FATAL ERROR in function [ParameterInput::" << who << "]" << std::endl
        << "Block name '" << b << "' not found when trying to set value for parameter '" << n << "'";
    ATHENA_ERROR(msg);
  }
  for (const Line &l : pb->lines)
    if (l.name == n) return &l;
  msg << "### FATAL ERROR in function [ParameterInput::" << who << "]" << std::endl
      << "Parameter name '" << n << "' not found in block '" << b << "'";
  ATHENA_ERROR(msg);
}

int ParameterInput::GetInteger(const std::string &b, const std::string &n) const {
  const std::string val = FindLine(b, n, "GetInteger")->value;
  // the reference converts through stod for e-notation (parameter_input.cpp:445-453)
  if (val.find_first_of("eE.") != std::string::npos) return static_cast<int>(std::stod(val));
  return std::atoi(val.c_str());
}
Real ParameterInput::GetReal(const std::string &b, const std::string &n) const {
  return std::atof(FindLine(b, n, "GetReal")->value.c_str());
}
bool ParameterInput::GetBoolean(const std::string &b, const std::string &n) const {
  std::string val = FindLine(b, n, "GetBoolean")->value;
  if (val.compare(0, 1, "0") == 0 || val.compare(0, 1, "1") == 0) return std::atoi(val.c_str()) != 0;
  std::transform(val.begin(), val.end(), val.begin(), ::tolower);
  return val == "true";
}
std::string ParameterInput::GetString(const std::string &b, const std::string &n) const {
  return FindLine(b, n, "GetString")->value;
}

static const char *kDefaultComment = "# Default value added at run time";

int ParameterInput::GetOrAddInteger(const std::string &b, const std::string &n, int def) {
  if (DoesParameterExist(b, n)) return GetInteger(b, n);
  Add(FindOrAdd(b), n, std::to_string(def), kDefaultComment);
  return def;
}
Real ParameterInput::GetOrAddReal(const std::string &b, const std::string &n, Real def) {
  if (DoesParameterExist(b, n)) return GetReal(b, n);
  // Stored as the reference stores it: default stream precision, 6 significant digits
  // (parameter_input.cpp:603-623).  The caller that adds the parameter gets `def` itself, later
  // callers read the string back -- which is why the MeshBlocks after the first see floors of
  // 3.46945e-18 where the first sees sqrt(1024*FLT_MIN) (Mesh constructor, athk_set_block_floors).
  std::stringstream ss;
  ss << def;
  Add(FindOrAdd(b), n, ss.str(), kDefaultComment);
  return def;
}
bool ParameterInput::GetOrAddBoolean(const std::string &b, const std::string &n, bool def) {
  if (DoesParameterExist(b, n)) return GetBoolean(b, n);
  Add(FindOrAdd(b), n, def ? "true" : "false", kDefaultComment);
  return def;
}
std::string ParameterInput::GetOrAddString(const std::string &b, const std::string &n, const std::string &def) {
  if (DoesParameterExist(b, n)) return GetString(b, n);
  Add(FindOrAdd(b), n, def, kDefaultComment);
  return def;
}
int ParameterInput::SetInteger(const std::string &b, const std::string &n, int v) {
  Add(FindOrAdd(b), n, std::to_string(v), "# Updated during run time");
  return v;
}
Real ParameterInput::SetReal(const std::string &b, const std::string &n, Real v) {
  std::stringstream ss;
  ss << v;   // 6 significant digits, as parameter_input.cpp:706-716
  Add(FindOrAdd(b), n, ss.str(), "# Updated during run time");
  return v;
}
std::string ParameterInput::SetString(const std::string &b, const std::string &n, const std::string &v) {
  Add(FindOrAdd(b), n, v, "# Updated during run time");
  return v;
}

void ParameterInput::ParameterDump(std::ostream &os) const {
  os << "#------------------------- PAR_DUMP -------------------------" << std::endl;
  for (const Block &b : blocks_) {
    os << "<" << b.name << ">" << std::endl;
    std::size_t ln = 0, lv = 0;
    for (const Line &l : b.lines) { ln = std::max(ln, l.name.size()); lv = std::max(lv, l.value.size()); }
    for (const Line &l : b.lines) {
      std::string n = l.name, v = l.value;
      n.append(ln - n.size() + 1, ' ');
      v.append(lv - v.size() + 1, ' ');
      os << n << "= " << v << l.comment << std::endl;
    }
  }
  os << "#------------------------- PAR_DUMP -------------------------" << std::endl;
  os << "<par_end>" << std::endl;
}

// Restart files carry every <outputN>'s next_time / file_number as of the dump.  The dump itself
// had already advanced its own block, and an -i file or the command line may change dt: step back
// by one dt here, then ForwardNextTime() moves each block to the first output time that is still
// ahead of (mesh_time - dt) once the final parameters and the mesh time are known.
void ParameterInput::RollbackNextTime() {
  for (Block &b : blocks_) {
    if (b.name.compare(0, 6, "output") != 0) continue;
    Real next_time = static_cast<Real>(std::atof(FindLine(b.name, "next_time", "RollbackNextTime")->value.c_str()));
    next_time -= static_cast<Real>(std::atof(FindLine(b.name, "dt", "RollbackNextTime")->value.c_str()));
    SetReal(b.name, "next_time", next_time);
  }
}

void ParameterInput::ForwardNextTime(Real mesh_time) {
  bool fresh = false;   // (sticky once a block without next_time has been seen, as in the reference)
  for (Block &b : blocks_) {
    if (b.name.compare(0, 6, "output") != 0) continue;
    Real next_time = mesh_time;
    if (DoesParameterExist(b.name, "next_time")) {
      next_time = static_cast<Real>(std::atof(FindLine(b.name, "next_time", "ForwardNextTime")->value.c_str()));
    } else {
      fresh = true;   // an output block added for the restarted run
    }
    const Real dt0 = static_cast<Real>(std::atof(FindLine(b.name, "dt", "ForwardNextTime")->value.c_str()));
    if (!(dt0 > 0.0)) continue;   // dcycle-driven block: nothing to move
    const Real dt = dt0 * static_cast<int>((mesh_time - next_time) / dt0) + dt0;
    if (dt > 0) {
      next_time += dt;
      if (fresh) next_time -= std::fmod(next_time, dt0) + dt0;   // write it immediately
    }
    std::stringstream ss;
    ss << next_time;
    Add(&b, "next_time", ss.str(), "# Updated during run time");
  }
}
