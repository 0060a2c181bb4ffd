"""ctypes loader for librst_b200.so (the C ABI declared in include/rst_b200.h).

The product path is CUDA-only: importing works without a GPU (so the build and the exported symbols can be
checked on a CPU box), but creating a solver without the compiled library or without a CUDA device raises --
there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "librst_b200.so")

OK = 0
ERR_INVALID, ERR_CUDA, ERR_ZERO_PIVOT, ERR_NAN, ERR_UNSUPPORTED, ERR_NOT_FACTORIZED = -1, -2, -3, -4, -5, -6

c_f64_p = C.POINTER(C.c_double)
c_i64_p = C.POINTER(C.c_int64)
c_int_p = C.POINTER(C.c_int)


class Desc(C.Structure):
    _fields_ = [
        ("problem_key", C.c_char_p), ("n_steps", C.c_int64), ("scheme", C.c_int), ("trap_time", C.c_int),
        ("t0", C.c_double), ("t_end", C.c_double), ("mesh", c_f64_p), ("n_bc", C.c_int), ("bc_var", c_int_p),
        ("bc_side", c_int_p), ("bc_val", c_f64_p), ("params", c_f64_p), ("n_params", C.c_int), ("lo", c_f64_p),
        ("hi", c_f64_p), ("rtol", c_f64_p), ("device", C.c_int), ("stream", C.c_void_p), ("slab", C.c_int),
        ("rank", C.c_int), ("world", C.c_int),
    ]


class NewtonOpts(C.Structure):
    _fields_ = [("max_iterations", C.c_int), ("max_jac", C.c_int), ("max_damp_iter", C.c_int),
                ("damp_factor", C.c_double), ("abs_tol", C.c_double)]


class NewtonStats(C.Structure):
    _fields_ = [("status", C.c_int), ("iterations", C.c_int), ("linear_solves", C.c_int), ("jac_recalcs", C.c_int),
                ("factorizations", C.c_int), ("last_s0", C.c_double), ("last_s1", C.c_double),
                ("last_conv", C.c_double), ("last_fbound", C.c_double), ("last_lambda", C.c_double),
                ("ms_assemble", C.c_double), ("ms_factor", C.c_double), ("ms_solve", C.c_double),
                ("ms_total", C.c_double), ("solves_reused", C.c_int)]


class FrozenOpts(C.Structure):
    _fields_ = [("strategy", C.c_int), ("strategy_params", C.c_double * 3), ("tolerance", C.c_double),
                ("max_iterations", C.c_int)]


FROZEN_STRATEGIES = {"Naive": 0, "Frozen_naive": 1, "every_m": 2, "at_high_morm": 3, "at_low_speed": 4, "complex": 5}


class Scalars(C.Structure):
    _fields_ = [("sumsq_step", C.c_double), ("tol_sum", C.c_double), ("fbound", C.c_double),
                ("nan_flag", C.c_double), ("device_flags", C.c_double)]


# LinearSolverPolicy / FallbackPolicy (solver_policy.rs:10-23), declaration order
LINEAR_SOLVER_POLICIES = {"Auto": 0, "ForceBlockTridiagonal": 1, "ForceBlockTridiagonalConsistent": 2, "ForceBanded": 3,
                          "ForceGeneralBandedPartialPivot": 4, "ForceFaerSparse": 5}
FALLBACK_POLICIES = {"Never": 0, "ToFaerSparse": 1}

ALLGATHER_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p)

# every symbol include/rst_b200.h declares: name -> (restype, argtypes)
_H = C.c_void_p
SYMBOLS = {
    "rst_b200_version": (C.c_int, []),
    "rst_b200_device_count": (C.c_int, []),
    "rst_b200_problem_count": (C.c_int, []),
    "rst_b200_problem_key": (C.c_char_p, [C.c_int]),
    "rst_b200_problem_info": (C.c_int, [C.c_char_p, c_int_p, c_int_p, C.POINTER(C.c_char_p)]),
    "rst_b200_last_error": (C.c_char_p, []),
    "rst_b200_create": (C.c_int, [C.POINTER(Desc), C.POINTER(_H)]),
    "rst_b200_destroy": (None, [_H]),
    "rst_b200_n_unknowns": (C.c_int64, [_H]),
    "rst_b200_set_params": (C.c_int, [_H, c_f64_p, C.c_int]),
    "rst_b200_sync": (C.c_int, [_H]),
    "rst_b200_set_state": (C.c_int, [_H, c_f64_p, C.c_size_t]),
    "rst_b200_get_state": (C.c_int, [_H, c_f64_p, C.c_size_t]),
    "rst_b200_get_full_solution": (C.c_int, [_H, c_f64_p, C.c_size_t]),
    "rst_b200_device_ptr": (C.c_void_p, [_H, C.c_int]),
    "rst_b200_eval_residual": (C.c_int, [_H, C.c_int]),
    "rst_b200_eval_jacobian": (C.c_int, [_H, C.c_int]),
    "rst_b200_factor": (C.c_int, [_H]),
    "rst_b200_solve": (C.c_int, [_H]),
    "rst_b200_factor_solve": (C.c_int, [_H]),
    "rst_b200_reduce": (C.c_int, [_H, C.c_int, C.POINTER(Scalars)]),
    "rst_b200_make_trial": (C.c_int, [_H, C.c_double]),
    "rst_b200_accept_trial": (C.c_int, [_H]),
    "rst_b200_checkpoint": (C.c_int, [_H, C.c_int]),
    "rst_b200_new_grid": (C.c_int, [_H, C.c_int, c_f64_p, C.c_double, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "rst_b200_new_grid_fetch": (C.c_int, [_H, c_f64_p, C.c_size_t, c_f64_p, C.c_size_t, C.POINTER(C.c_int32), C.c_size_t]),
    "rst_b200_create_refined": (C.c_int, [_H, C.POINTER(_H)]),
    "rst_b200_newton_solve": (C.c_int, [_H, C.POINTER(NewtonOpts), C.POINTER(NewtonStats)]),
    "rst_b200_newton_iterate": (C.c_int, [_H, C.POINTER(NewtonOpts), C.c_int, C.POINTER(NewtonStats)]),
    "rst_b200_newton_frozen": (C.c_int, [_H, C.POINTER(FrozenOpts), C.POINTER(NewtonStats)]),
    "rst_b200_solve_in_place": (C.c_int, [_H, c_f64_p, C.c_size_t]),
    "rst_b200_set_linear_solver_config": (C.c_int, [_H, C.c_int, C.c_int, C.c_int]),
    "rst_b200_get_linear_solver_config": (C.c_int, [_H, c_int_p, c_int_p, c_int_p]),
    "rst_b200_solve_in_place_with_refinement": (C.c_int, [_H, c_f64_p, C.c_size_t, C.c_int, C.c_double, c_int_p, c_f64_p]),
    "rst_b200_solve_multiple_in_place": (C.c_int, [_H, c_f64_p, C.c_size_t, C.c_size_t]),
    "rst_b200_jacobian_structure": (C.c_int64, [_H, c_i64_p, c_i64_p, c_i64_p, c_i64_p, c_int_p, c_int_p]),
    "rst_b200_aot_bind": (C.c_int, [_H]),
    "rst_b200_aot_bind_thread": (C.c_int, [_H]),
    "rst_b200_manifest_h": (C.c_int64, [_H, C.c_char_p, C.c_size_t]),
    "rustedscithe_aot_eval_residual": (C.c_int, [c_f64_p, C.c_size_t, c_f64_p, C.c_size_t]),
    "rustedscithe_aot_eval_jacobian_values": (C.c_int, [c_f64_p, C.c_size_t, c_f64_p, C.c_size_t]),
    "rst_b200_set_allgather": (C.c_int, [_H, ALLGATHER_FN, C.c_void_p]),
    "rst_b200_p2p_get_handle": (C.c_int, [_H, C.c_void_p, C.c_size_t]),
    "rst_b200_p2p_open": (C.c_int, [_H, C.c_void_p, C.c_size_t]),
    "rst_b200_p2p_selftest": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, c_i64_p, c_i64_p, c_int_p]),
    "rst_b200_dist_range": (C.c_int, [_H, c_i64_p, c_i64_p]),
    "rst_b200_factor_local": (C.c_int, [_H]),
    "rst_b200_factor_reduced": (C.c_int, [_H]),
    "rst_b200_solve_local_forward": (C.c_int, [_H]),
    "rst_b200_solve_reduced_backward": (C.c_int, [_H]),
    "rst_b200_reduce_local": (C.c_int, [_H, C.c_int]),
    "rst_b200_reduce_finish": (C.c_int, [_H, C.POINTER(Scalars)]),
    "rst_b200_btd_create": (C.c_int, [C.c_int64, C.c_int, C.c_int, C.c_void_p, C.POINTER(_H)]),
    "rst_b200_btd_destroy": (None, [_H]),
    "rst_b200_btd_n": (C.c_int64, [_H]),
    "rst_b200_btd_factor_from": (C.c_int, [_H, c_f64_p, c_f64_p, c_f64_p]),
    "rst_b200_btd_solve_in_place": (C.c_int, [_H, c_f64_p, C.c_size_t]),
    "rst_b200_btd_solve_multiple_in_place": (C.c_int, [_H, c_f64_p, C.c_size_t, C.c_size_t]),
    "rst_b200_memcpy": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]),
    "rst_b200_host_alloc": (C.c_int, [C.c_size_t, C.POINTER(C.c_void_p)]),
    "rst_b200_host_free": (C.c_int, [C.c_void_p]),
    "rst_b200_launch_count": (C.c_int64, [_H]),
}

_libs = {}


class RstB200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"rst_b200 error {code}: {msg}")
        self.code = code


def load(path: str | None = None):
    """dlopen a library and bind every declared symbol.  Raises if it has not been built."""
    path = os.path.abspath(path or LIB_PATH)
    if path not in _libs:
        if not os.path.exists(path):
            raise RuntimeError(
                f"{path} is missing: build it with `python -m rustedscithe_b200.build` "
                "(there is no CPU fallback for the B200 path)")
        L = C.CDLL(path)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)  # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _libs[path] = L
    return _libs[path]


def has_problem(L, key: str) -> bool:
    return any(L.rst_b200_problem_key(i) == key.encode() for i in range(L.rst_b200_problem_count()))


def check(rc, L=None):
    if rc != OK:
        raise RstB200Error(rc, ((L or load()).rst_b200_last_error() or b"").decode())
    return rc
