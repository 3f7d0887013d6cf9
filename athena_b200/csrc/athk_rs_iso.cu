// athk_rs_iso.cu -- kernel instantiations of the flux operator and the fused sweeps for the isothermal
// equation of state: HLLE for hydro (hydro/rsolvers/hydro/hlle.cpp, barotropic branch) and HLLD for MHD
// (hydro/rsolvers/mhd/hlld_iso.cpp); arithmetic in athk_rsolvers.h.
#include "athk_sweep_launch_impl.h"

ATHK_INSTANTIATE_RS(athk::RS_HLLE_ISO, false)
ATHK_INSTANTIATE_RS(athk::RS_HLLD_ISO, true)
