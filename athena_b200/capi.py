"""ctypes prototypes for the C ABI declared in include/athk.h.

``bind(cdll)`` attaches argtypes/restypes to a loaded library handle; ``load()`` loads the
in-tree CUDA library ``athena_b200/lib/libathk.so`` and fails loudly when it is missing.
There is no CPU implementation behind this module.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libathk.so")

# enum athk_array (include/athk.h)
ARRAYS = dict(
    U=0, U1=1, U2=2, W=3, BCC=4, B_X1F=5, B_X2F=6, B_X3F=7, B1_X1F=8, B1_X2F=9, B1_X3F=10,
    B2_X1F=11, B2_X2F=12, B2_X3F=13, FLUX_X1=14, FLUX_X2=15, FLUX_X3=16, E_X1E=17, E_X2E=18,
    E_X3E=19, E3_X1F=20, E2_X1F=21, E1_X2F=22, E3_X2F=23, E2_X3F=24, E1_X3F=25, WGHT_X1F=26,
    WGHT_X2F=27, WGHT_X3F=28, S=29, S1=30, S2=31, R=32, SFLUX_X1=33, SFLUX_X2=34, SFLUX_X3=35, W1=36)
RS_HLLC, RS_HLLD, RS_ROE, RS_HLLE, RS_LLF = range(5)
ARITH_STRICT, ARITH_FAST = 0, 1
EOS_ADIABATIC, EOS_ISOTHERMAL, EOS_ADIABATIC_SR = 0, 1, 2

# (nvar, a1, a2, a3): host shape = (nvar, nc3+a3, nc2+a2, nc1+a1); nvar None = nscalars
SHAPES = {}
for _n in ("U", "U1", "U2", "W", "W1"):
    SHAPES[_n] = (5, 0, 0, 0)
SHAPES["BCC"] = (3, 0, 0, 0)
for _p in ("B", "B1", "B2"):
    SHAPES[_p + "_X1F"] = (1, 1, 0, 0)
    SHAPES[_p + "_X2F"] = (1, 0, 1, 0)
    SHAPES[_p + "_X3F"] = (1, 0, 0, 1)
SHAPES.update(FLUX_X1=(5, 1, 0, 0), FLUX_X2=(5, 0, 1, 0), FLUX_X3=(5, 0, 0, 1),
              E_X1E=(1, 0, 1, 1), E_X2E=(1, 1, 0, 1), E_X3E=(1, 1, 1, 0),
              E3_X1F=(1, 1, 0, 0), E2_X1F=(1, 1, 0, 0), E1_X2F=(1, 0, 1, 0), E3_X2F=(1, 0, 1, 0),
              E2_X3F=(1, 0, 0, 1), E1_X3F=(1, 0, 0, 1),
              WGHT_X1F=(1, 1, 0, 0), WGHT_X2F=(1, 0, 1, 0), WGHT_X3F=(1, 0, 0, 1),
              S=(None, 0, 0, 0), S1=(None, 0, 0, 0), S2=(None, 0, 0, 0), R=(None, 0, 0, 0),
              SFLUX_X1=(None, 1, 0, 0), SFLUX_X2=(None, 0, 1, 0), SFLUX_X3=(None, 0, 0, 1))


class Config(C.Structure):
    _fields_ = [("nx1", C.c_int), ("nx2", C.c_int), ("nx3", C.c_int), ("nghost", C.c_int),
                ("nblocks", C.c_int), ("mhd", C.c_int), ("nscalars", C.c_int),
                ("rsolver", C.c_int), ("device", C.c_int), ("arith", C.c_int),
                ("gamma", C.c_double), ("dfloor", C.c_double), ("pfloor", C.c_double),
                ("sfloor", C.c_double), ("eos", C.c_int), ("iso_sound_speed", C.c_double),
                ("gamma_max", C.c_double)]


class StageWghts(C.Structure):
    _fields_ = [("delta", C.c_double), ("gamma_1", C.c_double), ("gamma_2", C.c_double),
                ("gamma_3", C.c_double), ("beta", C.c_double)]


_P = C.c_void_p
_D = C.POINTER(C.c_double)
_I = C.POINTER(C.c_int)
_LL = C.POINTER(C.c_longlong)

# name -> (restype, argtypes); every symbol include/athk.h declares
PROTOTYPES = {
    "athk_create": (C.c_int, [C.POINTER(Config), C.POINTER(_P)]),
    "athk_destroy": (None, [_P]),
    "athk_last_error": (C.c_char_p, []),
    "athk_sync": (C.c_int, [_P]),
    "athk_launch_count": (C.c_longlong, [_P]),
    "athk_stream": (_P, [_P]),
    "athk_profile": (C.c_int, [_P, C.c_int, C.c_char_p, C.c_int]),
    "athk_set_block_coords": (C.c_int, [_P, C.c_int] + [_D] * 9),
    "athk_set_neighbors": (C.c_int, [_P, _I]),
    "athk_upload": (C.c_int, [_P, C.c_int, C.c_int, _D]),
    "athk_download": (C.c_int, [_P, C.c_int, C.c_int, _D]),
    "athk_download_async": (C.c_int, [_P, C.c_int, C.c_int, _D]),
    "athk_upload_all": (C.c_int, [_P, C.c_int, _D, C.c_int]),
    "athk_download_all_async": (C.c_int, [_P, C.c_int, _D, C.c_int]),
    "athk_device_array": (C.c_int, [_P, C.c_int, C.POINTER(_P), _LL, _LL, _LL, _LL]),
    "athk_cons2prim": (C.c_int, [_P, C.c_int]),
    "athk_calculate_fluxes": (C.c_int, [_P, C.c_int, C.c_double, C.c_int]),
    "athk_calculate_scalar_fluxes": (C.c_int, [_P, C.c_int]),
    "athk_compute_corner_e": (C.c_int, [_P]),
    "athk_send_flux_correction": (C.c_int, [_P]),
    "athk_recv_flux_correction": (C.c_int, [_P]),
    "athk_weighted_ave_cc": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int, _D]),
    "athk_weighted_ave_fc": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, _D]),
    "athk_swap_registers": (C.c_int, [_P, C.c_int, C.c_int, C.c_int]),
    "athk_zero_register": (C.c_int, [_P, C.c_int, C.c_int]),
    "athk_add_flux_divergence": (C.c_int, [_P, C.c_double, C.c_int]),
    "athk_add_scalar_flux_divergence": (C.c_int, [_P, C.c_double, C.c_int]),
    "athk_ct": (C.c_int, [_P, C.c_double, C.c_int]),
    "athk_exchange_ghosts": (C.c_int, [_P, C.c_int, C.c_int]),
    "athk_new_block_dt": (C.c_int, [_P, C.c_double, _D, _D]),
    "athk_request_block_dt": (C.c_int, [_P]),
    "athk_history": (C.c_int, [_P, _D, C.c_int]),
    "athk_set_boundary_flags": (C.c_int, [_P, C.c_int, C.POINTER(C.c_int)]),
    "athk_apply_physical_boundaries": (C.c_int, [_P]),
    "athk_enroll_user_source": (C.c_int, [_P, _P, _P]),
    "athk_set_source_time": (C.c_int, [_P, C.c_double]),
    "athk_refine_enable": (C.c_int, [_P]),
    "athk_refine_set_block_coords": (C.c_int, [_P, C.c_int] + [_D] * 6),
    "athk_refine_set_peers": (C.c_int, [_P, C.c_int, _P]),
    "athk_refine_coarse_transfer": (C.c_int, [_P, C.c_int, C.c_int] + [_D] * 4),
    "athk_refine_set_ops": (C.c_int, [_P, C.c_int, _P, C.c_int]),
    "athk_refine_run": (C.c_int, [_P, C.c_int, C.c_int]),
    "athk_enroll_user_boundary": (C.c_int, [_P, C.c_int, _P, _P]),
    "athk_set_boundary_time": (C.c_int, [_P, C.c_double, C.c_double]),
    "athk_set_block_floors": (C.c_int, [_P, C.c_int, C.c_double, C.c_double, C.c_double]),
    "athk_set_constant_acceleration": (C.c_int, [_P, C.c_double, C.c_double, C.c_double]),
    "athk_add_source_terms": (C.c_int, [_P, C.c_double]),
    "athk_stage": (C.c_int, [_P, C.c_int, C.c_int, C.POINTER(StageWghts), C.c_double]),
    "athk_stage_phase": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.POINTER(StageWghts), C.c_double]),
    "athk_remote_ghost_count": (C.c_longlong, [_P, C.c_int, C.c_int, C.c_int]),
    "athk_remote_emf_count": (C.c_longlong, [_P, C.c_int, C.c_int, C.c_int]),
    "athk_remote_configure": (C.c_int, [_P, C.c_int, _I, _I, _I, _I, _LL, _LL, _LL, _LL]),
    "athk_remote_set_buffers": (C.c_int, [_P, _P, _P, _P, _P]),
    "athk_remote_pack_ghosts": (C.c_int, [_P, C.c_int]),
    "athk_remote_pack_emf": (C.c_int, [_P]),
    "athk_p2p_alloc": (C.c_int, [_P, C.c_longlong, C.POINTER(_P), C.c_char_p]),
    "athk_p2p_open": (C.c_int, [_P, C.c_char_p, C.POINTER(_P)]),
    "athk_p2p_close": (C.c_int, [_P, _P]),
    "athk_p2p_free": (C.c_int, [_P, _P]),
    "athk_p2p_enable_peer": (C.c_int, [_P, C.c_int]),
    "athk_remote_set_p2p": (C.c_int, [_P, C.c_int, C.POINTER(_P), C.POINTER(_P), _P, C.c_longlong, _P,
                                      C.c_longlong, C.POINTER(_P), _P]),
    "athk_p2p_timeouts": (C.c_longlong, [_P]),
}


class BvalArgs(C.Structure):
    """struct athk_bval_args (include/athk.h)"""
    _fields_ = [("lb", C.c_int), ("face", C.c_int), ("time", C.c_double), ("dt", C.c_double),
                ("il", C.c_int), ("iu", C.c_int), ("jl", C.c_int), ("ju", C.c_int), ("kl", C.c_int), ("ku", C.c_int),
                ("ngh", C.c_int),
                ("w", C.c_void_p), ("b1", C.c_void_p), ("b2", C.c_void_p), ("b3", C.c_void_p), ("r", C.c_void_p),
                ("stride_j", C.c_longlong), ("stride_k", C.c_longlong), ("stride_n", C.c_longlong),
                ("nscalars", C.c_int), ("stream", C.c_void_p)]


BVAL_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(BvalArgs))   # athk_bval_fn


def bind(cdll):
    """Attach prototypes; raises AttributeError if the library lacks a declared symbol."""
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(cdll, name)
        fn.restype = res
        fn.argtypes = args
    return cdll


_lib = None


def load():
    """Load the product CUDA library. No fallback: a missing build is an error."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "athena_b200: %s not found. Build it with `python -c 'import __graft_entry__ as g; "
                "g.build()'` (nvcc, sm_100a). There is no CPU fallback." % LIB_PATH)
        _lib = bind(C.CDLL(LIB_PATH))
    return _lib


class AthkError(RuntimeError):
    pass


def check(lib, rc):
    if rc != 0:
        raise AthkError("athk error %d: %s" % (rc, lib.athk_last_error().decode()))
