"""Code generation, NVRTC compilation (no GPU needed) and the C ABI surface."""
import ctypes
import os
import re

import numpy as np
import pytest

import pararealml_b200 as ns
from golden import cases
from pararealml_b200 import _native
from pararealml_b200.operators.fdm import codegen
from pararealml_b200.operators.fdm.fdm_operator import plan_overrides
from pararealml_b200.operators.fdm.lowering import lower_problem

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _source(case):
    ivp = case.build(ns)
    cp = ivp.constrained_problem
    low = lower_problem(cp)
    y0 = ivp.initial_condition.discrete_y_0(True)
    return low, codegen.generate_source(low.spec(**plan_overrides(cp, low, y0)))


def test_library_exports_every_symbol_of_the_header():
    header = open(os.path.join(ROOT, "include", "pararealml_b200.h")).read()
    declared = set(re.findall(r"\b(pml_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_native.SYMBOLS), declared ^ set(_native.SYMBOLS)
    lib = _native.lib()
    for name in declared:
        assert hasattr(lib, name)
    assert lib.pml_version() >= 100


def test_struct_layouts_match_the_header():
    assert ctypes.sizeof(_native.PlanDesc) == 4 * 20
    assert ctypes.sizeof(_native.Tables) == 8 * (6 * 4 + 3 + 4)
    assert ctypes.sizeof(_native.Workspace) == 8 * 10


@pytest.mark.parametrize("case", cases.FDM_CASES, ids=lambda c: c.name)
def test_every_case_generates_source(case):
    low, src = _source(case)
    assert f"#define PML_C {low.y_dim}" in src
    assert "pml_rhs_dt" in src and "PML_GENERATED_RHS" not in src
    # nothing is lambdified: the right-hand side is printed C
    assert "out[0] =" in src


def test_nvrtc_compiles_generated_kernels_for_sm_100a(tmp_path):
    _, src = _source(cases.FDM_BY_NAME["shallow_water_polar_rk4"])
    out = str(tmp_path / "k.cubin")
    _native.compile_to_cubin(src, out)
    blob = open(out, "rb").read()
    assert blob[:4] == b"\x7fELF" and len(blob) > 10000


def test_nvrtc_reports_errors():
    with pytest.raises(RuntimeError, match="NVRTC"):
        _native.compile_to_cubin("this is not CUDA", "/tmp/pml_bad.cubin")


def test_integer_powers_become_multiplication_chains():
    _, src = _source(cases.FDM_BY_NAME["cahn_hilliard_3d_rk4"])
    assert "pow(" not in src.split("pml_rhs_aux")[1].split("}")[0]


def test_passthrough_only_for_static_problems():
    low, src = _source(cases.FDM_BY_NAME["cahn_hilliard_3d_rk4"])
    assert "#define PML_PASSTHROUGH 1" in src
    low, src = _source(cases.FDM_BY_NAME["wave_2d_dynamic_rk4"])
    assert "#define PML_PASSTHROUGH 0" in src


def test_vector_laplacian_symbol_fails_like_the_reference():
    class Eq(ns.DifferentialEquation):
        def __init__(self):
            super().__init__(2, 2, [(0, 1)])

        @property
        def symbolic_equation_system(self):
            s = self.symbols
            return ns.SymbolicEquationSystem(
                [s.y_vector_laplacian[0, 1, 0], s.y_vector_laplacian[0, 1, 1]]
            )

    mesh = ns.Mesh([(0.0, 1.0), (0.0, 1.0)], [0.25, 0.25])
    bc = ns.NeumannBoundaryCondition(lambda x, t: np.zeros((len(x), 2)), is_static=True)
    cp = ns.ConstrainedProblem(Eq(), mesh, [(bc, bc)] * 2)
    with pytest.raises(KeyError):
        plan_overrides(cp, lower_problem(cp), None)


def test_too_few_points_is_rejected():
    eq = ns.DiffusionEquation(1)
    mesh = ns.Mesh([(0.0, 1.0)], [1.0])
    bc = ns.NeumannBoundaryCondition(lambda x, t: np.zeros((len(x), 1)), is_static=True)
    cp = ns.ConstrainedProblem(eq, mesh, [(bc, bc)])
    with pytest.raises(ValueError):
        codegen.generate_source(lower_problem(cp).spec())


def test_product_path_fails_loudly_without_a_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from pararealml_b200.operators.fdm import FDMOperator, RK4, ThreePointCentralDifferenceMethod

    case = cases.FDM_BY_NAME["diffusion_2d_rk4"]
    op = FDMOperator(RK4(), ThreePointCentralDifferenceMethod(), case.d_t)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        op.solve(case.build(ns))
