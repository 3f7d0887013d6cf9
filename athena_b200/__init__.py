"""athena_b200 -- B200-native (sm_100a) per-MeshBlock hydro/MHD finite-volume update behind the
Athena++ operator API.  The compute path is the CUDA library athena_b200/lib/libathk.so (C ABI in
include/athk.h); this package only binds it.  Importing the package does not touch the GPU;
``capi.load()`` raises if the CUDA library has not been built -- there is no CPU fallback."""
from . import capi  # noqa: F401

__all__ = ["capi", "pyhost"]
