// athk_rs_llf.cu -- kernel instantiations of the flux operator and the fused sweeps for
// --flux=llf (hydro/rsolvers/hydro/llf.cpp, hydro/rsolvers/mhd/llf_mhd.cpp); arithmetic in athk_rsolvers.h.
#include "athk_sweep_launch_impl.h"

ATHK_INSTANTIATE_RS(athk::RS_LLF, false)
ATHK_INSTANTIATE_RS(athk::RS_LLF, true)
