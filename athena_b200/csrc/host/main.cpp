// main.cpp -- driver of the B200 host: argv -> ParameterInput -> Mesh -> TimeIntegratorTaskList
// -> main loop, following src/main.cpp:65-666 (flags :125-203, setup :230-424, loop :457-571,
// wrap-up :584-652).  Prints the reference's own throughput lines (zone-cycles/...second).
#include <signal.h>
#include <unistd.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iterator>

#include "athena.hpp"

// Graceful termination (utils/signal_handler.cpp:34-141, main.cpp:216-218,568-575,613-618): SIGTERM,
// SIGINT and the -t wall-time alarm only raise a flag; the main loop finishes its cycle, and the
// final outputs -- including the restart dump to resume from -- are written as usual.
namespace {
volatile sig_atomic_t g_signal_term = 0, g_signal_int = 0, g_signal_alrm = 0;
void RaiseSignalFlag(int s) {
  if (s == SIGTERM) g_signal_term = 1;
  if (s == SIGINT) g_signal_int = 1;
  if (s == SIGALRM) g_signal_alrm = 1;
  signal(s, RaiseSignalFlag);
}
bool SignalRaised() { return g_signal_term || g_signal_int || g_signal_alrm; }

void ShowConfig() {   // -c: what configure.py fixes in the reference is chosen at run time here
  std::cout << "This Athena++ executable (B200 path) is configured with:" << std::endl
            << "  Problem generator:          " << g_config.problem << std::endl
            << "  Coordinate system:          cartesian" << std::endl
            << "  Equation of state:          adiabatic" << std::endl
            << "  Riemann solver:             "
            << (g_config.flux.empty() ? (g_config.mhd ? "hlld" : "hllc") : g_config.flux) << std::endl
            << "  Magnetic fields:            " << (g_config.mhd ? "ON" : "OFF") << std::endl
            << "  Special relativity:         " << (g_config.sr ? "ON" : "OFF") << std::endl
            << "  Equation of state:          " << (g_config.isothermal ? "isothermal" : "adiabatic") << std::endl
            << "  Number of ghost cells:      " << g_config.nghost << std::endl
            << "  Number of scalars:          " << g_config.nscalars << std::endl
            << "  Floating-point precision:   double" << std::endl
            << "  Device:                     CUDA sm_100a through libathk (include/athk.h)" << std::endl;
}
}  // namespace

static void ParseConfigure(const std::string &s) {
  // the tokens configure.py would take: --prob=<name> -b --flux=<name> --nghost=<n> --nscalars=<n>
  std::stringstream ss(s);
  std::string t;
  while (ss >> t) {
    if (t.compare(0, 7, "--prob=") == 0) g_config.problem = t.substr(7);
    else if (t == "-b") g_config.mhd = true;
    else if (t == "-s") g_config.sr = true;
    else if (t.compare(0, 7, "--flux=") == 0) g_config.flux = t.substr(7);
    else if (t.compare(0, 9, "--nghost=") == 0) g_config.nghost = std::atoi(t.c_str() + 9);
    else if (t.compare(0, 6, "--eos=") == 0) {
      const std::string e = t.substr(6);
      if (e != "adiabatic" && e != "isothermal") {
        std::cout << "### FATAL ERROR in main" << std::endl << "--eos=" << e
                  << ": the B200 host implements the adiabatic and the isothermal equation of state" << std::endl;
        std::exit(1);
      }
      g_config.isothermal = (e == "isothermal");
    }
    else if (t.compare(0, 11, "--nscalars=") == 0) g_config.nscalars = std::atoi(t.c_str() + 11);
  }
}

int main(int argc, char *argv[]) {
  std::string input_filename, restart_filename, run_dir;
  bool narg_flag = false, have_cfg = false, carg_flag = false;
  int wtlim = 0;
  std::string cfg_cmdline;
  for (int i = 1; i < argc; ++i) {
    const std::string a = argv[i];
    if (a == "-i" && i + 1 < argc) input_filename = argv[++i];
    else if (a == "-r" && i + 1 < argc) restart_filename = argv[++i];
    else if (a == "-d" && i + 1 < argc) run_dir = argv[++i];
    else if (a == "-n") narg_flag = true;
    else if (a == "-c") carg_flag = true;
    else if (a == "-t" && i + 1 < argc) {   // -t <hh:mm:ss>
      int wth = 0, wtm = 0, wts = 0;
      std::sscanf(argv[++i], "%d:%d:%d", &wth, &wtm, &wts);
      wtlim = wth * 3600 + wtm * 60 + wts;
    }
    else if (a == "-h") {
      std::cout << "Athena++ (B200 path)" << std::endl
                << "Usage: " << argv[0] << " [options] [block/par=value ...]" << std::endl
                << "  -i <file>   specify input file [athinput]" << std::endl
                << "  -r <file>   restart with this file" << std::endl
                << "  -d <dir>    specify run dir [current dir]" << std::endl
                << "  -n          parse input file and quit" << std::endl
                << "  -c          show configuration and quit" << std::endl
                << "  -t hh:mm:ss wall time limit for final output" << std::endl
                << "  --prob=<pgen> [-b] [-s] [--eos=..] [--flux=..] [--nghost=n] [--nscalars=n]  configure-time choices "
                   "of the reference (default: the athinput's <comment> configure line)" << std::endl;
      return 0;
    } else if (a.compare(0, 2, "--") == 0 || a == "-b" || a == "-s") {
      cfg_cmdline += a + " ";
      have_cfg = true;
    }
  }
  if (carg_flag && input_filename.empty() && restart_filename.empty()) {
    if (have_cfg) ParseConfigure(cfg_cmdline);
    ShowConfig();
    return 0;
  }
  if (input_filename.empty() && restart_filename.empty()) {
    std::cout << "### FATAL ERROR in main" << std::endl << "No input file or restart file is specified." << std::endl;
    return 1;
  }
  try {
    ParameterInput *pinput = new ParameterInput;
    std::string restart_raw;
    std::size_t restart_data = 0;
    if (!restart_filename.empty()) {   // main.cpp:231-239: parameters first from the restart file ...
      std::ifstream rf(restart_filename.c_str(), std::ios::binary);
      if (!rf.good()) {
        std::stringstream msg;
        msg << "### FATAL ERROR in function [IOWrapper:Open]" << std::endl
            << "Input file '" << restart_filename << "' could not be opened" << std::endl;
        ATHENA_ERROR(msg);
      }
      restart_raw.assign(std::istreambuf_iterator<char>(rf), std::istreambuf_iterator<char>());
      const std::size_t end = restart_raw.find("<par_end>\n");
      if (end == std::string::npos) {
        std::stringstream msg;
        msg << "### FATAL ERROR in function [ParameterInput::LoadFromFile]" << std::endl
            << "<par_end> not found in restart file '" << restart_filename << "'" << std::endl;
        ATHENA_ERROR(msg);
      }
      restart_data = end + 10;
      std::istringstream par(restart_raw.substr(0, restart_data));
      pinput->LoadFromStream(par);
      pinput->RollbackNextTime();
    }
    if (!input_filename.empty()) pinput->LoadFromFile(input_filename);   // ... then -i overrides them
    pinput->ModifyFromCmdline(argc, argv);
    if (have_cfg) ParseConfigure(cfg_cmdline);
    else if (pinput->DoesParameterExist("comment", "configure")) ParseConfigure(pinput->GetString("comment", "configure"));
    if (carg_flag) {
      ShowConfig();
      return 0;
    }
    signal(SIGTERM, RaiseSignalFlag);
    signal(SIGINT, RaiseSignalFlag);
    signal(SIGALRM, RaiseSignalFlag);
    if (wtlim > 0) alarm(wtlim);
    if (!run_dir.empty() && chdir(run_dir.c_str()) != 0) {
      std::stringstream msg;
      msg << "### FATAL ERROR in function [ChangeRunDir]" << std::endl << "Cannot cd to directory '" << run_dir << "'";
      ATHENA_ERROR(msg);
    }
    Mesh *pmesh = new Mesh(pinput);
    if (!restart_filename.empty()) {
      pmesh->LoadRestart(restart_raw, restart_data);
      std::string().swap(restart_raw);
      pinput->ForwardNextTime(pmesh->time);   // main.cpp:304-308
    }
    TimeIntegratorTaskList *ptlist = new TimeIntegratorTaskList(pinput, pmesh);
    if (narg_flag) {
      pinput->ParameterDump(std::cout);
      return 0;
    }
    pmesh->Initialize(pinput);
    Outputs *pouts = new Outputs(pmesh, pinput);
    pouts->MakeOutputs(pmesh, pinput);

    std::cout << "\nSetup complete, entering main loop...\n" << std::endl;
    std::uint64_t mbcnt = 0;
    const auto t0 = std::chrono::steady_clock::now();
    while ((pmesh->time < pmesh->tlim) && (pmesh->nlim < 0 || pmesh->ncycle < pmesh->nlim)) {
      pmesh->OutputCycleDiagnostics();
      for (int stage = 1; stage <= ptlist->nstages; ++stage) {
        // NEW_DT rides on the last stage's W(U) pass (same per-cell values, exact min)
        if (stage == ptlist->nstages)
          for (athk_ctx *a : pmesh->ctx) AthkCheck(athk_request_block_dt(a), "main");
        ptlist->DoTaskListOneStage(pmesh, stage);
      }
      pmesh->ncycle++;
      pmesh->time += pmesh->dt;
      mbcnt += pmesh->nbtotal;
      pmesh->NewTimeStep();
      if (pmesh->time < pmesh->tlim) pouts->MakeOutputs(pmesh, pinput);   // final output happens later (main.cpp:546)
      if (SignalRaised()) break;
    }
    if (wtlim > 0) alarm(0);
    for (athk_ctx *a : pmesh->ctx) AthkCheck(athk_sync(a), "main");
    const double wsec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    pmesh->OutputCycleDiagnostics();
    pmesh->UserWorkAfterLoop(pinput);
    pouts->MakeOutputs(pmesh, pinput, true);   // final outputs (main.cpp:589)

    if (g_signal_term) std::cout << std::endl << "Terminating on Terminate signal" << std::endl;
    else if (g_signal_int) std::cout << std::endl << "Terminating on Interrupt signal" << std::endl;
    else if (g_signal_alrm) std::cout << std::endl << "Terminating on wall-time limit" << std::endl;
    else if (pmesh->ncycle == pmesh->nlim) std::cout << std::endl << "Terminating on cycle limit" << std::endl;
    else std::cout << std::endl << "Terminating on time limit" << std::endl;
    char buf[256];
    std::snprintf(buf, sizeof(buf), "time=%.16e cycle=%d\ntlim=%.16e nlim=%d\n", pmesh->time, pmesh->ncycle,
                  pmesh->tlim, pmesh->nlim);
    std::cout << buf << std::endl;
    const MeshBlock *pb = pmesh->my_blocks(0);
    const std::uint64_t zonecycles = mbcnt * static_cast<std::uint64_t>(pb->block_size.nx1) * pb->block_size.nx2 *
                                     pb->block_size.nx3;   // main.cpp:639-641
    std::snprintf(buf, sizeof(buf),
                  "zone-cycles = %llu\ngpu wtime used = %.16e\nzone-cycles/gpu_wsecond = %.16e\n"
                  "kernel launches = %lld\n",
                  (unsigned long long)zonecycles, wsec, zonecycles / wsec, [&] {
                    long long n = 0;
                    for (athk_ctx *a : pmesh->ctx) n += athk_launch_count(a);
                    return n;
                  }());
    std::cout << buf << std::endl;
    delete pouts;
    delete ptlist;
    delete pmesh;
    delete pinput;
  } catch (std::bad_alloc &ba) {
    std::cout << "### FATAL ERROR in main" << std::endl << "memory allocation failed: " << ba.what() << std::endl;
    return 1;
  } catch (std::exception const &ex) {
    std::cout << ex.what() << std::endl;   // the reference prints the ATHENA_ERROR message and exits(1)
    return 1;
  }
  return 0;
}
