"""ctypes binding of the C-ABI library (include/pararealml_b200.h).

The product path has no CPU fallback: if the library is missing or a call
fails, a ``RuntimeError`` is raised.
"""
import ctypes
import os
from ctypes import (
    POINTER,
    Structure,
    byref,
    c_char_p,
    c_double,
    c_int,
    c_longlong,
    c_void_p,
)

LIB_NAME = "libpararealml_b200.so"
LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), LIB_NAME)

INTEGRATOR_CODES = {"forward_euler": 0, "explicit_midpoint": 1, "rk4": 2}


class PlanDesc(Structure):
    _fields_ = [
        ("n_dims", c_int),
        ("shape", c_int * 3),
        ("y_dim", c_int),
        ("n_dt", c_int),
        ("n_alg", c_int),
        ("n_lap", c_int),
        ("block", c_int * 3),
        ("fused", c_int),
        ("fused_tile", c_int * 2),
        ("fused_zc", c_int),
        ("fused_threads", c_int),
        ("fused_smem", c_int * 2),
        ("small_threads", c_int),
        ("zrep", c_int),
    ]


class Tables(Structure):
    _fields_ = [
        ("neu", c_void_p * 6),
        ("neu_stride", c_longlong * 6),
        ("dir", c_void_p * 6),
        ("dir_stride", c_longlong * 6),
        ("coord", c_void_p * 3),
        ("aux", c_void_p * 4),
    ]


class Workspace(Structure):
    _fields_ = [
        ("u_a", c_void_p),
        ("u_b", c_void_p),
        ("acc", c_void_p),
        ("lap_rhs", c_void_p),
        ("jac_a", c_void_p),
        ("jac_b", c_void_p),
        ("partials", c_void_p),
        ("flags", c_void_p),
        ("t_dev", c_void_p),
        ("t_capacity", c_longlong),
    ]


IC_MAX_COMPONENTS = 16


class IcMesh(Structure):
    _fields_ = [
        ("n_dims", c_int),
        ("shape", c_int * 3),
        ("coord", c_int),
        ("axis_dev", c_void_p * 3),
        ("trig_dev", c_void_p * 4),
    ]


class IcGaussianParams(Structure):
    _fields_ = [
        ("mean", c_double * 3),
        ("whiten", c_double * 9),
        ("log_norm", c_double),
        ("multiplier", c_double),
    ]


# every symbol declared in include/pararealml_b200.h
SYMBOLS = {
    "pml_ic_gaussian": (
        c_int,
        [POINTER(IcMesh), c_int, POINTER(IcGaussianParams), c_void_p, c_void_p],
    ),
    "pml_ic_separable": (
        c_int,
        [POINTER(IcMesh), c_int, POINTER(c_void_p), POINTER(c_double), c_void_p,
         c_void_p],
    ),
    "pml_apply_dirichlet": (c_int, [c_void_p, c_void_p, c_longlong, c_void_p]),
    "pml_last_error": (c_char_p, []),
    "pml_version": (c_int, []),
    "pml_plan_create": (
        c_int,
        [c_char_p, POINTER(PlanDesc), c_char_p, POINTER(c_void_p)],
    ),
    "pml_compile_to_cubin": (c_int, [c_char_p, c_char_p]),
    "pml_plan_destroy": (c_int, [c_void_p]),
    "pml_plan_set_tables": (c_int, [c_void_p, POINTER(Tables)]),
    "pml_plan_launches": (c_longlong, [c_void_p]),
    "pml_fdm_run": (
        c_int,
        [
            c_void_p, c_int, POINTER(Workspace), c_void_p, c_void_p,
            c_longlong, POINTER(c_double), c_int, c_double, c_longlong,
            c_void_p, c_double, c_longlong, POINTER(c_int), c_void_p,
        ],
    ),
    "pml_fdm_run_batch": (
        c_int,
        [
            c_void_p, c_int, POINTER(Workspace), c_void_p, c_void_p,
            c_longlong, c_int, c_longlong, c_longlong, c_longlong,
            POINTER(c_double), c_int, c_double, c_longlong, c_void_p,
        ],
    ),
    "pml_fdm_phase_count": (c_int, [c_void_p, c_int]),
    "pml_fdm_phase": (
        c_int,
        [
            c_void_p, c_int, POINTER(Workspace), c_void_p, c_void_p, c_double,
            c_double, c_longlong, c_int, POINTER(c_void_p), c_void_p,
        ],
    ),
    "pml_fdm_phase_planes": (
        c_int,
        [
            c_void_p, c_int, POINTER(Workspace), c_void_p, c_void_p, c_double,
            c_double, c_longlong, c_int, c_int, c_int, POINTER(c_void_p), c_void_p,
        ],
    ),
    "pml_eval_rhs": (
        c_int, [c_void_p, c_void_p, c_void_p, c_double, c_longlong, c_void_p]
    ),
    "pml_jacobi_run": (
        c_int,
        [
            c_void_p, POINTER(Workspace), c_void_p, c_void_p, c_void_p,
            c_longlong, c_double, c_longlong, POINTER(c_int), c_void_p,
        ],
    ),
    "pml_aos_to_soa": (
        c_int, [c_void_p, c_void_p, c_longlong, c_int, c_longlong, c_void_p]
    ),
    "pml_soa_to_aos": (
        c_int, [c_void_p, c_void_p, c_longlong, c_int, c_longlong, c_void_p]
    ),
    "pml_parareal_correction": (
        c_int, [c_void_p, c_void_p, c_void_p, c_longlong, c_void_p]
    ),
    "pml_parareal_update": (
        c_int,
        [
            c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
            c_longlong, c_int, c_void_p,
        ],
    ),
    "pml_parareal_shift": (
        c_int,
        [c_void_p, c_longlong, c_longlong, c_void_p, c_void_p, c_longlong, c_void_p],
    ),
}

_lib = None


def lib():
    """Loads the library once; raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_NAME} has not been built (run `python -c 'import "
                "__graft_entry__ as g; g.build()'` or `make -C "
                "pararealml_b200/csrc`); the B200 operators have no CPU "
                "fallback"
            )
        handle = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes) in SYMBOLS.items():
            fn = getattr(handle, name)
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = handle
    return _lib


def check(code: int):
    if code != 0:
        msg = lib().pml_last_error()
        raise RuntimeError(
            "pararealml_b200 native call failed: "
            + (msg.decode(errors="replace") if msg else f"code {code}")
        )


def compile_to_cubin(source: str, path: str):
    check(lib().pml_compile_to_cubin(source.encode(), path.encode()))


__all__ = [
    "PlanDesc", "Tables", "Workspace", "lib", "check", "compile_to_cubin",
    "INTEGRATOR_CODES", "byref", "c_void_p", "c_int", "c_double",
]
