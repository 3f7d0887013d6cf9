// athk_rs_roe.cu -- kernel instantiations of the flux operator and the fused sweeps for
// --flux=roe (hydro/rsolvers/hydro/roe.cpp, hydro/rsolvers/mhd/roe_mhd.cpp); arithmetic in athk_rsolvers.h.
#include "athk_sweep_launch_impl.h"

ATHK_INSTANTIATE_RS(athk::RS_ROE, false)
ATHK_INSTANTIATE_RS(athk::RS_ROE, true)
