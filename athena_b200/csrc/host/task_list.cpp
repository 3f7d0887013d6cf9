// task_list.cpp -- TimeIntegratorTaskList: integrator weights and the per-stage operator
// sequence.  Weights and limits follow src/task_list/time_integrator.cpp:111-242 (vl2), :243-337
// (rk1), :338-465 (rk2), :466-606 (rk3), the CFL clamp :893-903; the task bodies :1606-2318.
// The reference runs a per-MeshBlock task DAG with polling (task_list.cpp:30-89); every operator
// here acts on all MeshBlocks of the rank in one launch, so a stage is a stream-ordered sequence
// (dependencies of the DAG = launch order; Receive* tasks = stream order).
#include "athena.hpp"

TimeIntegratorTaskList::TimeIntegratorTaskList(ParameterInput *pin, Mesh *pm) {
  integrator = pin->GetOrAddString("time", "integrator", "vl2");
  std::stringstream msg;
  auto set = [&](int n, Real delta, Real g1, Real g2, Real g3, Real beta) {
    stage_wghts[n] = IntegratorWeight{delta, g1, g2, g3, beta};
  };
  if (integrator == "vl2") {
    nstages = 2;
    cfl_limit = 1.0;
    if (pm->ndim == 2) cfl_limit = 0.5;
    if (pm->ndim == 3) cfl_limit = 0.5;
    set(0, 1.0, 0.0, 1.0, 0.0, 0.5);
    set(1, 0.0, 0.0, 1.0, 0.0, 1.0);
  } else if (integrator == "rk1") {
    nstages = 1;
    cfl_limit = 1.0;
    set(0, 1.0, 0.0, 1.0, 0.0, 1.0);
  } else if (integrator == "rk2") {
    nstages = 2;
    cfl_limit = 1.0;
    set(0, 1.0, 0.0, 1.0, 0.0, 1.0);
    set(1, 0.0, 0.5, 0.5, 0.0, 0.5);
  } else if (integrator == "rk3") {
    nstages = 3;
    cfl_limit = 1.0;
    set(0, 1.0, 0.0, 1.0, 0.0, 1.0);
    set(1, 0.0, 0.25, 0.75, 0.0, 0.25);
    set(2, 0.0, TWO_3RD, ONE_3RD, 0.0, TWO_3RD);
  } else {
    msg << "### FATAL ERROR in TimeIntegratorTaskList constructor" << std::endl
        << "integrator=" << integrator << " not valid time integrator" << std::endl;
    ATHENA_ERROR(msg);
  }
  // Reconstruction constructor (reconstruct/reconstruction.cpp:39-101)
  const std::string input_recon = pin->GetOrAddString("time", "xorder", "2");
  if (input_recon == "1") xorder = 1;
  else if (input_recon == "2") xorder = 2;
  else if (input_recon == "3") xorder = 3;
  else {
    msg << "### FATAL ERROR in Reconstruction constructor" << std::endl
        << "xorder=" << input_recon << " not valid choice for reconstruction" << std::endl;
    ATHENA_ERROR(msg);
  }
  if (xorder == 3 && NGHOST < 3) {   // reconstruction.cpp:117-127
    msg << "### FATAL ERROR in Reconstruction constructor" << std::endl
        << "xorder=" << input_recon << " (PPM) reconstruction selected, but nghost=" << NGHOST << std::endl
        << "Reconfigure with --nghost=XXX with XXX > 2" << std::endl;
    ATHENA_ERROR(msg);
  }
  Real cfl_number = pin->GetReal("time", "cfl_number");
  if (cfl_number > cfl_limit) {   // time_integrator.cpp:893-903
    std::cout << "### Warning in TimeIntegratorTaskList constructor" << std::endl
              << "User CFL number " << cfl_number << " must be smaller than " << cfl_limit
              << " for integrator=" << integrator << " in " << pm->ndim << "D simulation" << std::endl
              << "Setting to limit" << std::endl;
    cfl_number = cfl_limit;
  }
  pm->cfl_number = cfl_number;
}

void TimeIntegratorTaskList::DoTaskListOneStage(Mesh *pm, int stage) {
  athk_ctx *a = pm->athk;
  const IntegratorWeight &sw = stage_wghts[stage - 1];
  const bool mhd = MAGNETIC_FIELDS_ENABLED;
  // CalculateHydroFlux: donor cell on the VL2 predictor (time_integrator.cpp:1700-1701)
  const int order = (integrator == "vl2" && stage == 1) ? 1 : xorder;
  const Real dt = pm->dt;
  if (pm->fused_stage) {
    athk_stage_wghts w = {sw.delta, sw.gamma_1, sw.gamma_2, sw.gamma_3, sw.beta};
    AthkCheck(athk_stage(a, stage, order, &w, dt), "TimeIntegratorTaskList (athk_stage)");
    return;
  }
  const char *who = "TimeIntegratorTaskList::DoTaskListOneStage";
  if (stage == 1) {   // StartupTaskList (:1608-1619)
    AthkCheck(athk_zero_register(a, 0, 1), who);
    if (mhd) AthkCheck(athk_zero_register(a, 1, 1), who);
    if (NSCALARS > 0) AthkCheck(athk_zero_register(a, 2, 1), who);   // ps->s1.ZeroClear() :1628-1630
  }
  AthkCheck(athk_calculate_fluxes(a, order, dt, 0), who);                 // CALC_HYDFLX
  if (NSCALARS > 0) AthkCheck(athk_calculate_scalar_fluxes(a, order), who);   // CALC_SCLRFLX :2329-2340
  if (mhd) {
    AthkCheck(athk_compute_corner_e(a), who);                            // CALC_FLDFLX
    AthkCheck(athk_send_flux_correction(a), who);                        // SEND_FLDFLX
    AthkCheck(athk_recv_flux_correction(a), who);                        // RECV_FLDFLX
  }
  // IntegrateHydro (:1797-1847)
  const Real w1[5] = {1.0, sw.delta, 0.0, 0.0, 0.0};
  const Real w2[5] = {sw.gamma_1, sw.gamma_2, sw.gamma_3, 0.0, 0.0};
  const bool swap = (w2[0] == 0.0 && w2[1] == 1.0 && w2[2] == 0.0);
  AthkCheck(athk_weighted_ave_cc(a, 0, 1, 0, 2, w1), who);
  if (swap) AthkCheck(athk_swap_registers(a, 0, 0, 1), who);
  else AthkCheck(athk_weighted_ave_cc(a, 0, 0, 1, 2, w2), who);
  AthkCheck(athk_add_flux_divergence(a, sw.beta * dt, 0), who);
  AthkCheck(athk_add_source_terms(a, sw.beta * dt), who);               // SRC_TERM :1890-1913
  if (NSCALARS > 0) {                                                   // INT_SCLR :2376-2403
    AthkCheck(athk_weighted_ave_cc(a, 1, 1, 0, 2, w1), who);
    if (swap) AthkCheck(athk_swap_registers(a, 2, 0, 1), who);
    else AthkCheck(athk_weighted_ave_cc(a, 1, 0, 1, 2, w2), who);
    AthkCheck(athk_add_scalar_flux_divergence(a, sw.beta * dt, 0), who);
  }
  if (mhd) {   // IntegrateField (:1853-1885)
    AthkCheck(athk_weighted_ave_fc(a, 1, 0, 2, w1), who);
    if (swap) AthkCheck(athk_swap_registers(a, 1, 0, 1), who);
    else AthkCheck(athk_weighted_ave_fc(a, 0, 1, 2, w2), who);
    AthkCheck(athk_ct(a, sw.beta * dt, 0), who);
  }
  AthkCheck(athk_exchange_ghosts(a, 1 | (mhd ? 2 : 0) | (NSCALARS > 0 ? 4 : 0), 0), who);   // SEND/RECV/SETB_HYD,_FLD,_SCLR
  AthkCheck(athk_cons2prim(a, 0), who);                            // CONS2PRIM
  AthkCheck(athk_apply_physical_boundaries(a), who);               // PHY_BVAL :2276-2311
}
