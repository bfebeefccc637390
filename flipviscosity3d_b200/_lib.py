"""ctypes binding of libflip_b200.so (C ABI in include/flip_b200.h).

The product path is the nvcc/sm_100a library under flipviscosity3d_b200/lib/.  There is no CPU
fallback: if the library is missing or no CUDA device is usable, loading / flip_create raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_LIB = os.path.join(_HERE, "lib", "libflip_b200.so")


class flip_stats(C.Structure):
    _fields_ = [
        ("substeps", C.c_int64), ("particles", C.c_int64), ("kernel_launches", C.c_int64),
        ("pressure_iterations", C.c_int32), ("pressure_converged", C.c_int32),
        ("pressure_active_blocks", C.c_int32),
        ("viscosity_iterations", C.c_int32), ("viscosity_converged", C.c_int32),
        ("viscosity_active_blocks", C.c_int32), ("viscosity_applied", C.c_int32),
        ("reserved", C.c_int32),
        ("pressure_residual", C.c_double), ("viscosity_residual", C.c_double),
        ("pressure_rhs_max", C.c_double), ("viscosity_rhs_max", C.c_double),
        ("stage_ms", C.c_float * 8),
        ("pressure_solve_ms", C.c_float), ("viscosity_solve_ms", C.c_float),
        ("pressure_unknowns", C.c_int64), ("viscosity_unknowns", C.c_int64),
        ("viscosity_setup_ms", C.c_float), ("reserved2", C.c_float),
    ]


# every symbol include/flip_b200.h declares: name -> (restype, argtypes)
_H = C.c_void_p
_FP = C.POINTER(C.c_float)
_BP = C.POINTER(C.c_uint8)
SYMBOLS = {
    "flip_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_float, C.POINTER(_H)]),
    "flip_destroy": (C.c_int, [_H]),
    "flip_last_error": (C.c_char_p, [_H]),
    "flip_synchronize": (C.c_int, [_H]),
    "flip_set_solid_sdf": (C.c_int, [_H, _FP]),
    "flip_set_particles": (C.c_int, [_H, _FP, C.c_int64]),
    "flip_get_particles": (C.c_int, [_H, _FP, C.c_int64, C.POINTER(C.c_int64)]),
    "flip_reset_boundary": (C.c_int, [_H]),
    "flip_add_boundary_mesh": (C.c_int, [_H, _FP, C.c_int, C.POINTER(C.c_int32), C.c_int, C.c_int]),
    "flip_add_liquid_mesh": (C.c_int, [_H, _FP, C.c_int, C.POINTER(C.c_int32), C.c_int, C.POINTER(C.c_int64)]),
    "flip_mesh_sdf": (C.c_int, [_H, _FP, C.c_int, C.POINTER(C.c_int32), C.c_int, _FP]),
    "flip_srand": (C.c_int, [C.c_uint]),
    "flip_rand": (C.c_int, []),
    "flip_get_positions_async": (C.c_int, [_H, _FP, C.c_int64, C.POINTER(C.c_int64)]),
    "flip_output_wait": (C.c_int, [_H]),
    "flip_num_particles": (C.c_int, [_H, C.POINTER(C.c_int64)]),
    "flip_set_viscosity_uniform": (C.c_int, [_H, C.c_float]),
    "flip_set_viscosity_grid": (C.c_int, [_H, _FP]),
    "flip_set_gravity": (C.c_int, [_H, C.c_float, C.c_float, C.c_float]),
    "flip_advance": (C.c_int, [_H, C.c_float, C.POINTER(C.c_int)]),
    "flip_substep": (C.c_int, [_H, C.c_float]),
    "flip_cfl": (C.c_int, [_H, _FP]),
    "flip_stage_update_liquid_sdf": (C.c_int, [_H]),
    "flip_stage_advect_velocity_field": (C.c_int, [_H]),
    "flip_stage_add_body_force": (C.c_int, [_H, C.c_float]),
    "flip_stage_apply_viscosity": (C.c_int, [_H, C.c_float]),
    "flip_stage_project": (C.c_int, [_H, C.c_float]),
    "flip_stage_constrain": (C.c_int, [_H]),
    "flip_stage_advect_particles": (C.c_int, [_H, C.c_float]),
    "flip_solve_pressure": (C.c_int, [_H, C.c_float]),
    "flip_apply_pressure": (C.c_int, [_H, C.c_float]),
    "flip_extrapolate": (C.c_int, [_H]),
    "flip_viscosity_volumes": (C.c_int, [_H]),
    "flip_get_field": (C.c_int, [_H, C.c_int, _FP]),
    "flip_set_field": (C.c_int, [_H, C.c_int, _FP]),
    "flip_get_valid": (C.c_int, [_H, C.c_int, _BP]),
    "flip_set_valid": (C.c_int, [_H, C.c_int, _BP]),
    "flip_set_param": (C.c_int, [_H, C.c_char_p, C.c_double]),
    "flip_get_stats": (C.c_int, [_H, C.POINTER(flip_stats)]),
    "flip_time_kernel": (C.c_int, [_H, C.c_char_p, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_uint64)]),
    "flip_event_record": (C.c_int, [_H, C.c_int]),
    "flip_event_elapsed_ms": (C.c_int, [_H, C.c_int, C.c_int, C.POINTER(C.c_float)]),
    "flip_dist_unique_id": (C.c_int, [C.c_void_p]),
    "flip_dist_init": (C.c_int, [_H, C.c_int, C.c_int, C.c_void_p]),
    "flip_dist_p2p_blob_size": (C.c_int, []),
    "flip_dist_p2p_export": (C.c_int, [_H, C.c_void_p]),
    "flip_dist_p2p_import": (C.c_int, [_H, C.c_void_p]),
    "flip_host_alloc": (C.c_int, [C.POINTER(C.c_void_p), C.c_uint64]),
    "flip_host_free": (C.c_int, [C.c_void_p]),
    "flip_version": (C.c_char_p, []),
}


def load_library(path=None):
    """Load the C-ABI library and declare every prototype.  Raises if it is missing."""
    path = path or DEFAULT_LIB
    if not os.path.exists(path):
        raise RuntimeError(
            "flipviscosity3d_b200: %s not found. Build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'` (needs nvcc; targets sm_100a). There is no CPU fallback." % path)
    lib = C.CDLL(path)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    return lib


_default = None


def default_library():
    global _default
    if _default is None:
        _default = load_library()
    return _default
