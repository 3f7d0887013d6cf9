// athk_fast.cu -- the fast-arithmetic set of the fused sweeps (athk_config.arith = ATHK_ARITH_FAST).
// The ONLY translation unit compiled with FMA contraction on (--fmad=true, see Makefile): same
// kernels, math policy MathFast (athk_math.h).  Results agree with the strict (bit-exact) set to
// rounding; the strict set is what parity is claimed for.
#include "athk_ctx.h"
#include "athk_sweep_launch_impl.h"

ATHK_INSTANTIATE_RS_FAST(athk::RS_HLLC, false)
ATHK_INSTANTIATE_RS_FAST(athk::RS_HLLD, true)
