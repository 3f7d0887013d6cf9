// athk_rs_hlle.cu -- kernel instantiations of the flux operator and the fused sweeps for
// --flux=hlle (hydro/rsolvers/hydro/hlle.cpp, hydro/rsolvers/mhd/hlle_mhd.cpp); arithmetic in athk_rsolvers.h.
#include "athk_sweep_launch_impl.h"

ATHK_INSTANTIATE_RS(athk::RS_HLLE, false)
ATHK_INSTANTIATE_RS(athk::RS_HLLE, true)
