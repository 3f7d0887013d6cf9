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
  // (ebeta / sbeta: fraction of dt at the end / beginning of the stage, handed to boundary / source functions)
  auto set = [&](int n, Real delta, Real g1, Real g2, Real g3, Real beta, Real ebeta, Real sbeta) {
    stage_wghts[n] = IntegratorWeight{delta, g1, g2, g3, beta, ebeta, sbeta};
  };
  if (integrator == "vl2") {
    nstages = 2;
    cfl_limit = 1.0;
    if (pm->ndim == 2) cfl_limit = 0.5;
    if (pm->ndim == 3) cfl_limit = 0.5;
    set(0, 1.0, 0.0, 1.0, 0.0, 0.5, 0.5, 0.0);
    set(1, 0.0, 0.0, 1.0, 0.0, 1.0, 1.0, 0.5);
  } else if (integrator == "rk1") {
    nstages = 1;
    cfl_limit = 1.0;
    set(0, 1.0, 0.0, 1.0, 0.0, 1.0, 1.0, 0.0);
  } else if (integrator == "rk2") {
    nstages = 2;
    cfl_limit = 1.0;
    set(0, 1.0, 0.0, 1.0, 0.0, 1.0, 1.0, 0.0);
    set(1, 0.0, 0.5, 0.5, 0.0, 0.5, 1.0, 1.0);
  } else if (integrator == "rk3") {
    nstages = 3;
    cfl_limit = 1.0;
    set(0, 1.0, 0.0, 1.0, 0.0, 1.0, 1.0, 0.0);
    set(1, 0.0, 0.25, 0.75, 0.0, 0.25, 0.5, 1.0);
    set(2, 0.0, TWO_3RD, ONE_3RD, 0.0, TWO_3RD, 1.0, 0.5);
  } else if (integrator == "rk4" || integrator == "ssprk5_4") {
    // RK4()4[2S] of Ketcheson (2010), time_integrator.cpp:607-735; SSPRK(5,4) of Gottlieb (2009) in
    // 3S* form, :737-880 (its fourth stage also updates the third register, see DoTaskListOneStage)
    if (integrator == "rk4") {
      nstages = 4;
      cfl_limit = 1.3925;
      set(0, 1.0, 0.0, 1.0, 0.0, 1.193743905974738, 0.0, 0.0);
      set(1, 0.217683334308543, 0.121098479554482, 0.721781678111411, 0.0, 0.099279895495783, 0.0, 0.0);
      set(2, 1.065841341361089, -3.843833699660025, 2.121209265338722, 0.0, 1.131678018054042, 0.0, 0.0);
      set(3, 0.0, 0.546370891121863, 0.198653035682705, 0.0, 0.310665766509336, 0.0, 0.0);
    } else {
      nstages = 5;
      cfl_limit = 1.508;
      set(0, 1.0, 0.0, 1.0, 0.0, 0.391752226571890, 0.0, 0.0);
      set(1, 0.0, 0.555629506348765, 0.444370493651235, 0.0, 0.368410593050371, 0.0, 0.0);
      set(2, 0.517231671970585, 0.379898148511597, 0.0, 0.620101851488403, 0.251891774271694, 0.0, 0.0);
      set(3, 0.096059710526147, 0.821920045606868, 0.0, 0.178079954393132, 0.544974750228521, 0.0, 0.0);
      set(4, 0.0, 0.386708617503268, 1.0, 1.0, 0.226007483236906, 0.0, 0.0);
      pm->fused_stage = false;   // three registers: the operator sequence
    }
    // beginning / end of stage as fractions of dt (:722-735, :866-879)
    Real temp = 0.0, temp_prev = 0.0;
    stage_wghts[0].sbeta = 0.0;
    for (int l = 0; l < nstages - 1; l++) {
      temp_prev = temp;
      temp = temp_prev + stage_wghts[l].delta * stage_wghts[l].sbeta;
      stage_wghts[l].ebeta = stage_wghts[l].gamma_1 * temp_prev + stage_wghts[l].gamma_2 * temp +
                             stage_wghts[l].gamma_3 * 0.0 + stage_wghts[l].beta;
      stage_wghts[l + 1].sbeta = stage_wghts[l].ebeta;
    }
    stage_wghts[nstages - 1].ebeta = 1.0;
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
  const IntegratorWeight &sw = stage_wghts[stage - 1];
  const bool mhd = MAGNETIC_FIELDS_ENABLED;
  // CalculateHydroFlux: donor cell on the VL2 predictor (time_integrator.cpp:1700-1701)
  const int order = (integrator == "vl2" && stage == 1) ? 1 : xorder;
  const Real dt = pm->dt;
  const char *who = "TimeIntegratorTaskList::DoTaskListOneStage";
  // Every call acts on all MeshBlocks of one GPU and only enqueues work, so a task is issued to
  // every GPU before the next one (the reference's per-MeshBlock task loop, task_list.cpp:66-117,
  // with the GPUs running concurrently); messages between GPUs travel inside the pack calls.
  auto all = [&](auto &&fn) { for (athk_ctx *a : pm->ctx) AthkCheck(fn(a), who); };
  // what PhysicalBoundary hands to user-enrolled boundary functions (:2284-2287)
  all([&](athk_ctx *a) { return athk_set_boundary_time(a, pm->time + sw.ebeta * dt, sw.beta * dt); });
  pm->bc_time_ = pm->time + sw.ebeta * dt;   // ... and Prolongation to ProlongateBoundaries (:2150-2153)
  pm->bc_dt_ = sw.beta * dt;
  // ... and AddSourceTerms to user source functions: the time at the beginning of the stage (:1901-1903)
  all([&](athk_ctx *a) { return athk_set_source_time(a, pm->time + sw.sbeta * dt); });
  if (pm->fused_stage) {
    athk_stage_wghts w = {sw.delta, sw.gamma_1, sw.gamma_2, sw.gamma_3, sw.beta};
    if (pm->ngpu == 1) {
      AthkCheck(athk_stage(pm->athk, stage, order, &w, dt), "TimeIntegratorTaskList (athk_stage)");
    } else {   // the same stage cut at its two exchange points
      for (int phase = 0; phase < 3; ++phase)
        all([&](athk_ctx *a) { return athk_stage_phase(a, phase, stage, order, &w, dt); });
    }
    return;
  }
  const bool ssprk = (integrator == "ssprk5_4");
  if (stage == 1) {   // StartupTaskList (:1608-1619)
    const Real copy[5] = {0.0, 1.0, 0.0, 0.0, 0.0};   // u2 = u (:1612-1613, :1631-1632)
    all([&](athk_ctx *a) { return athk_zero_register(a, 0, 1); });
    if (ssprk) all([&](athk_ctx *a) { return athk_weighted_ave_cc(a, 0, 2, 0, 1, copy); });
    if (mhd) {
      all([&](athk_ctx *a) { return athk_zero_register(a, 1, 1); });
      if (ssprk) {   // :1620-1626
        std::stringstream msg;
        msg << "### FATAL ERROR in TimeIntegratorTaskList::StartupTaskList\n"
            << "integrator=" << integrator << " is currently incompatible with MHD." << std::endl;
        ATHENA_ERROR(msg);
      }
    }
    if (NSCALARS > 0) {
      all([&](athk_ctx *a) { return athk_zero_register(a, 2, 1); });   // ps->s1.ZeroClear() :1628-1630
      if (ssprk) all([&](athk_ctx *a) { return athk_weighted_ave_cc(a, 1, 2, 0, 1, copy); });
    }
  }
  all([&](athk_ctx *a) { return athk_calculate_fluxes(a, order, dt, 0); });                    // CALC_HYDFLX
  if (pm->multilevel) {   // SEND/RECV_HYDFLX: fluxes on faces shared with finer blocks (:1757-1795)
    pm->RunRefineList(0);
    if (pm->ngpu > 1) pm->SyncAll();
  }
  if (NSCALARS > 0) all([&](athk_ctx *a) { return athk_calculate_scalar_fluxes(a, order); });  // CALC_SCLRFLX :2329-2340
  if (NSCALARS > 0 && pm->multilevel) {   // SEND/RECV_SCLRFLX (:2342-2374)
    pm->RunRefineList(66);
    if (pm->ngpu > 1) pm->SyncAll();
  }
  if (mhd) {
    all([&](athk_ctx *a) { return athk_compute_corner_e(a); });                                // CALC_FLDFLX
    all([&](athk_ctx *a) { return athk_send_flux_correction(a); });                            // SEND_FLDFLX
    if (pm->ngpu > 1) all([&](athk_ctx *a) { return athk_remote_pack_emf(a); });
    all([&](athk_ctx *a) { return athk_recv_flux_correction(a); });                            // RECV_FLDFLX
    if (pm->multilevel) {   // ... from finer blocks (flux_correction_fc.cpp:1648-1680): clear, add neighbour by
                            // neighbour (lists 16..63, most of them empty), average
      pm->RunRefineList(14);
      for (int list = 16; list < 16 + 48; ++list) pm->RunRefineList(list);
      pm->RunRefineList(15);
      if (pm->ngpu > 1) pm->SyncAll();
    }
  }
  // IntegrateHydro (:1797-1847)
  const Real w1[5] = {1.0, sw.delta, 0.0, 0.0, 0.0};
  const Real w2[5] = {sw.gamma_1, sw.gamma_2, sw.gamma_3, 0.0, 0.0};
  const bool swap = (w2[0] == 0.0 && w2[1] == 1.0 && w2[2] == 0.0);
  all([&](athk_ctx *a) { return athk_weighted_ave_cc(a, 0, 1, 0, 2, w1); });
  if (swap) all([&](athk_ctx *a) { return athk_swap_registers(a, 0, 0, 1); });
  else all([&](athk_ctx *a) { return athk_weighted_ave_cc(a, 0, 0, 1, 2, w2); });
  all([&](athk_ctx *a) { return athk_add_flux_divergence(a, sw.beta * dt, 0); });
  // the penultimate stage of SSPRK(5,4) also forms part of u^(n+1) in the third register (:1828-1841)
  const Real neg[5] = {-1.0, 0.0, 0.0, 0.0, 0.0};
  const Real wght_ssp = 0.063692468666290 * dt;
  if (ssprk && stage == 4) {
    all([&](athk_ctx *a) { return athk_weighted_ave_cc(a, 0, 2, 1, 2, neg); });
    all([&](athk_ctx *a) { return athk_add_flux_divergence(a, wght_ssp, 2); });
  }
  all([&](athk_ctx *a) { return athk_add_source_terms(a, sw.beta * dt); });                    // SRC_TERM :1890-1913
  if (NSCALARS > 0) {                                                                          // INT_SCLR :2376-2403
    all([&](athk_ctx *a) { return athk_weighted_ave_cc(a, 1, 1, 0, 2, w1); });
    if (swap) all([&](athk_ctx *a) { return athk_swap_registers(a, 2, 0, 1); });
    else all([&](athk_ctx *a) { return athk_weighted_ave_cc(a, 1, 0, 1, 2, w2); });
    all([&](athk_ctx *a) { return athk_add_scalar_flux_divergence(a, sw.beta * dt, 0); });
    if (ssprk && stage == 4) {   // :2405-2416
      all([&](athk_ctx *a) { return athk_weighted_ave_cc(a, 1, 2, 1, 2, neg); });
      all([&](athk_ctx *a) { return athk_add_scalar_flux_divergence(a, wght_ssp, 2); });
    }
  }
  if (mhd) {   // IntegrateField (:1853-1885)
    all([&](athk_ctx *a) { return athk_weighted_ave_fc(a, 1, 0, 2, w1); });
    if (swap) all([&](athk_ctx *a) { return athk_swap_registers(a, 1, 0, 1); });
    else all([&](athk_ctx *a) { return athk_weighted_ave_fc(a, 0, 1, 2, w2); });
    all([&](athk_ctx *a) { return athk_ct(a, sw.beta * dt, 0); });
  }
  pm->ExchangeGhosts();                                                // SEND/RECV/SETB_HYD,_FLD,_SCLR
  all([&](athk_ctx *a) { return athk_cons2prim(a, 0); });              // CONS2PRIM
  all([&](athk_ctx *a) { return athk_apply_physical_boundaries(a); }); // PHY_BVAL :2276-2311
}
