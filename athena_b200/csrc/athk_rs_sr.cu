// athk_rs_sr.cu -- kernel instantiations of the flux operator and the fused sweeps for special-
// relativistic hydro (hydro/rsolvers/hydro/hlle_rel.cpp, hllc_rel.cpp, llf_rel.cpp; one kernel
// set, the solver is chosen by EosParams::sr_solver); arithmetic in athk_rsolvers.h.
#include "athk_sweep_launch_impl.h"

ATHK_INSTANTIATE_RS(athk::RS_SR, false)
