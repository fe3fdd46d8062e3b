"""ctypes view of include/reak_b200.h (the C-ABI of libreak_b200.so).

The structures below mirror `rkb_element`, `rkb_base_frame` and `rkb_chain_desc` field by
field.  `load_library()` fails loudly when the CUDA library has not been built: there is no
CPU fallback behind this package.
"""
import ctypes as C
import os

RKB_MAX_COORDS = 16
RKB_MAX_FREE = 1

# enum rkb_kind
REVOLUTE_3D, PRISMATIC_3D, FREE_3D, RIGID_LINK_3D, INERTIA_3D, INERTIA_GEN, ACTUATOR_GEN = 1, 2, 3, 4, 5, 6, 7
TORSION_SPRING_3D, TORSION_DAMPER_3D, SPRING_3D, DAMPER_3D = 8, 9, 10, 11
RIGID_LINK_GEN, SPRING_GEN, DAMPER_GEN, COORD_GEN = 12, 13, 14, 15
REVOLUTE_2D, PRISMATIC_2D, FREE_2D, RIGID_LINK_2D, INERTIA_2D = 17, 18, 19, 20, 21
TORSION_SPRING_2D, TORSION_DAMPER_2D, SPRING_2D, DAMPER_2D = 24, 25, 26, 27

MEM_HOST, MEM_DEVICE = 0, 1
LAYOUT_AOS, LAYOUT_SOA, LAYOUT_BLOCKED = 0, 2, 4
STATUS_SINGULAR, STATUS_NONFINITE = 1, 2

OK, ERR_INVALID, ERR_UNSUPPORTED, ERR_DIMENSION, ERR_CUDA, ERR_NOMEM, ERR_INTEGRATION = 0, -1, -2, -3, -4, -5, -6


class rkb_element(C.Structure):
    _fields_ = [("kind", C.c_int32), ("frame_a", C.c_int32), ("frame_b", C.c_int32),
                ("coord", C.c_int32), ("aux", C.c_int32), ("reserved", C.c_int32),
                ("upstream", C.c_uint64), ("p", C.c_double * 12)]


class rkb_base_frame(C.Structure):
    _fields_ = [("position", C.c_double * 3), ("quat", C.c_double * 4), ("velocity", C.c_double * 3),
                ("ang_velocity", C.c_double * 3), ("acceleration", C.c_double * 3),
                ("ang_acceleration", C.c_double * 3)]


class rkb_chain_desc(C.Structure):
    _fields_ = [("dim", C.c_int32), ("n_elements", C.c_int32), ("n_frames", C.c_int32),
                ("n_coords", C.c_int32), ("n_inputs", C.c_int32), ("base_frame", C.c_int32),
                ("base", rkb_base_frame), ("elements", C.POINTER(rkb_element))]


assert C.sizeof(rkb_element) == 128, C.sizeof(rkb_element)


class rkb_shape(C.Structure):
    _fields_ = [("kind", C.c_int32), ("anchor", C.c_int32), ("position", C.c_double * 3), ("quat", C.c_double * 4),
                ("dims", C.c_double * 3)]


SHAPE_PLANE, SHAPE_SPHERE, SHAPE_CCYLINDER, SHAPE_CYLINDER, SHAPE_BOX = 1, 2, 3, 4, 5
SHAPE_CIRCLE, SHAPE_CRECT, SHAPE_RECTANGLE = 6, 7, 8   # planar shapes (planar chains)
PROXY_MAX_SHAPES = 16


class rkb_rollout_opts(C.Structure):
    _fields_ = [("scheme", C.c_int32), ("n_intervals", C.c_int32), ("steps_per_interval", C.c_int32),
                ("reserved", C.c_int32), ("dt", C.c_double)]


class rkb_steer_opts(C.Structure):
    _fields_ = [("time_step", C.c_double), ("dt", C.c_double), ("goal_proximity", C.c_double), ("substeps", C.c_int32),
                ("max_intervals", C.c_int32), ("saturate_first", C.c_int32), ("reserved", C.c_int32),
                ("u_lower", C.c_void_p), ("u_upper", C.c_void_p), ("du_lower", C.c_void_p), ("du_upper", C.c_void_p)]


CREATE_INTERPRETER, CREATE_GENERAL = 1, 2
OPT_SPLIT_MAX_SAMPLES, OPT_FUSED_STEER, OPT_FUSED_SEQUENCE, OPT_HOST_PIPELINE, OPT_AUTO_SPECIALIZE = 1, 2, 3, 4, 5
SCHEME_EULER, SCHEME_MIDPOINT, SCHEME_RK4, SCHEME_RK5 = 1, 2, 4, 5
SCHEMES = {"euler": SCHEME_EULER, "midpoint": SCHEME_MIDPOINT, "rk4": SCHEME_RK4, "runge_kutta4": SCHEME_RK4,
           "rk5": SCHEME_RK5, "runge_kutta5": SCHEME_RK5}

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RKB_LIB_PATH") or os.path.join(_PKG_DIR, "lib", "libreak_b200.so")
_lib = None

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)

# name -> (restype, argtypes); every symbol include/reak_b200.h declares
SYMBOLS = {
    "rkb_version": (C.c_int, []),
    "rkb_build_id": (C.c_char_p, []),
    "rkb_strerror": (C.c_char_p, [C.c_int]),
    "rkb_last_cuda_error": (C.c_char_p, []),
    "rkb_chain_create": (C.c_int, [C.POINTER(rkb_chain_desc), C.POINTER(C.c_void_p)]),
    "rkb_chain_create_ex": (C.c_int, [C.POINTER(rkb_chain_desc), C.c_uint, C.POINTER(C.c_void_p)]),
    "rkb_chain_set_option": (C.c_int, [C.c_void_p, C.c_int, C.c_longlong]),
    "rkb_chain_get_option": (C.c_longlong, [C.c_void_p, C.c_int]),
    "rkb_chain_destroy": (None, [C.c_void_p]),
    "rkb_chain_state_dim": (C.c_int, [C.c_void_p]),
    "rkb_chain_input_dim": (C.c_int, [C.c_void_p]),
    "rkb_chain_dof": (C.c_int, [C.c_void_p]),
    "rkb_chain_is_serial": (C.c_int, [C.c_void_p]),
    "rkb_chain_shape": (C.c_uint64, [C.c_void_p]),
    "rkb_chain_kernel_shape": (C.c_uint64, [C.c_void_p]),
    "rkb_chain_specialize": (C.c_int, [C.c_void_p, C.c_int]),
    "rkb_chain_is_specialized": (C.c_int, [C.c_void_p]),
    "rkb_eval": (C.c_int, [C.c_void_p, C.c_int, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                           C.c_uint, C.c_void_p]),
    "rkb_rollout_rk4": (C.c_int, [C.c_void_p, C.c_int, C.c_size_t, C.c_void_p, C.c_void_p, C.c_double, C.c_int,
                                  C.c_void_p, C.c_void_p, C.c_uint, C.c_void_p]),
    "rkb_rollout_rk4_inputs": (C.c_int, [C.c_void_p, C.c_int, C.c_size_t, C.c_void_p, C.c_void_p, C.c_double, C.c_int,
                                         C.c_void_p, C.c_void_p, C.c_uint, C.c_void_p]),
    "rkb_rollout": (C.c_int, [C.c_void_p, C.c_int, C.c_size_t, C.c_void_p, C.c_void_p, C.POINTER(rkb_rollout_opts),
                              C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint, C.c_void_p]),
    "rkb_rollout_rk4_scatter": (C.c_int, [C.c_void_p, C.c_int, C.c_size_t, C.c_void_p, C.c_void_p, C.c_double, C.c_int, C.c_int,
                                          C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_size_t, C.c_uint, C.c_void_p]),
    "rkb_rollout_rk4_multi": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.c_size_t, C.c_void_p, C.c_void_p, C.c_double,
                                        C.c_int, C.c_void_p, C.c_void_p]),
    "rkb_gen_forces": (C.c_int, [C.c_void_p, C.c_int, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.c_uint, C.c_void_p]),
    "rkb_mass_matrix": (C.c_int, [C.c_void_p, C.c_int, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_uint, C.c_void_p]),
    "rkb_chain_frame_count": (C.c_int, [C.c_void_p]),
    "rkb_chain_wave_samples": (C.c_longlong, [C.c_void_p, C.c_int]),
    "rkb_frames": (C.c_int, [C.c_void_p, C.c_int, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint, C.c_void_p]),
    "rkb_proxy_create": (C.c_int, [C.c_void_p, C.POINTER(rkb_shape), C.c_int, C.POINTER(rkb_shape), C.c_int, C.POINTER(C.c_void_p)]),
    "rkb_proxy_destroy": (None, [C.c_void_p]),
    "rkb_proxy_set_option": (C.c_int, [C.c_void_p, C.c_int, C.c_longlong]),
    "rkb_proxy_specialize": (C.c_int, [C.c_void_p, C.c_int]),
    "rkb_proxy_is_specialized": (C.c_int, [C.c_void_p]),
    "rkb_proxy_source": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "rkb_steer_checked_specialize": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int]),
    "rkb_steer_checked_is_specialized": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "rkb_steer_checked_source": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]),
    "rkb_proxy_finder_count": (C.c_int, [C.c_void_p]),
    "rkb_proxy_finder": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "rkb_proxy_program": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "rkb_chain_program": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "rkb_is_free": (C.c_int, [C.c_void_p, C.c_int, C.c_size_t, C.c_void_p, C.POINTER(C.c_void_p), C.c_int, C.c_void_p, C.c_uint, C.c_void_p]),
    "rkb_min_distance": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_uint, C.c_void_p]),
    "rkb_collision_points": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_size_t, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_uint, C.c_void_p]),
    "rkb_twist_shaping_rows": (C.c_int, [C.c_void_p]),
    "rkb_twist_shaping_mcm": (C.c_int, [C.c_void_p, C.c_void_p]),
    "rkb_twist_shaping": (C.c_int, [C.c_void_p, C.c_int, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint, C.c_void_p]),
    "rkb_frame_jacobian": (C.c_int, [C.c_void_p, C.c_int, C.c_size_t, C.c_void_p, C.c_int, C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint, C.c_void_p]),
    "rkb_linearize": (C.c_int, [C.c_void_p, C.c_int, C.c_size_t, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p,
                                C.c_uint, C.c_void_p]),
    "rkb_steer_batch": (C.c_int, [C.c_void_p, C.c_int, C.c_size_t, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_double, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_uint, C.c_void_p]),
    "rkb_steer_feedback": (C.c_int, [C.c_void_p, C.c_int, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.POINTER(rkb_steer_opts), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint, C.c_void_p]),
    "rkb_steer_feedback_checked": (C.c_int, [C.c_void_p, C.c_int, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                             C.POINTER(rkb_steer_opts), C.POINTER(C.c_void_p), C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                             C.c_void_p, C.c_void_p, C.c_uint, C.c_void_p]),
    "rkb_rkx_read": (C.c_int, [C.c_char_p, C.c_void_p, C.c_void_p, C.c_int, C.c_char_p, C.c_size_t]),
    "rkb_rkx_load": (C.c_int, [C.c_char_p, C.c_uint, C.POINTER(C.c_void_p), C.c_char_p, C.c_size_t]),
    "rkb_nearest": (C.c_int, [C.c_int, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_void_p, C.c_void_p,
                              C.c_void_p, C.c_uint, C.c_void_p]),
    "rkb_nearest_last_error": (C.c_char_p, []),
    "rkb_last_kernel_ms": (C.c_double, [C.c_void_p]),
    "rkb_launch_count": (C.c_uint64, [C.c_void_p]),
    "rkb_host_pin": (C.c_int, [C.c_void_p, C.c_size_t]),
    "rkb_host_unpin": (C.c_int, [C.c_void_p]),
    "rkb_measure_fp64_peak": (C.c_int, [C.c_int, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
}


def load_library(path=None):
    """dlopen libreak_b200.so and bind every declared symbol.  Raises if it is missing."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.isfile(p):
        raise RuntimeError(
            "reak_b200: CUDA library %s is not built (run `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C reak_b200/csrc`). There is no CPU fallback." % p)
    lib = C.CDLL(p)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if path is None:
        _lib = lib
    return lib


class RkbError(RuntimeError):
    def __init__(self, code, what=""):
        self.code = code
        lib = load_library()
        msg = lib.rkb_strerror(code).decode()
        cu = lib.rkb_last_cuda_error().decode()
        RuntimeError.__init__(self, "%s%s (code %d)%s" % (what + ": " if what else "", msg, code,
                                                          " [cuda: %s]" % cu if cu and code == ERR_CUDA else ""))


def check(code, what=""):
    if code != 0:
        raise RkbError(code, what)
