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


@pytest.mark.parametrize(
    "shape,y_dim",
    [((512, 512, 512), 3), ((64, 48, 40), 1), ((4096, 4096), 3), ((150, 600), 1),
     ((20, 21, 22), 6), ((9, 9, 4), 2)],
)
@pytest.mark.parametrize("variant", [1, 2])
def test_fused_tile_fits_the_hardware(shape, y_dim, variant, monkeypatch):
    """Geometry of the fused stage-pair kernels: thread, TMA box and shared
    memory limits of sm_100a hold for every mesh the default rule accepts."""
    for key in ("PML_FUSE", "PML_FTILE", "PML_FDEPTH", "PML_FZC", "PML_FMIN_BLOCKS",
                "PML_FROWS", "PML_FSYNC"):
        monkeypatch.delenv(key, raising=False)
    monkeypatch.setenv("PML_FVARIANT", str(variant))
    tile = codegen.default_fused(shape, y_dim, y_dim, False)
    assert tile is not None and tile.variant == variant
    hy = 1 if len(shape) == 3 else 0
    assert tile.tx % 2 == 0 and tile.tx + 4 <= 256
    # one thread per cell of the stage-A tile, or per ``rows`` cells of a column
    assert (tile.tx + 2) * (tile.ty + 2 * hy) <= tile.threads * tile.rows
    assert tile.threads <= 1024
    if variant == 2:
        assert (tile.tx + 2) % 32 == 0 and (tile.ty + 2 * hy) % tile.rows == 0
    assert tile.threads % 32 == 0
    assert tile.smem_first <= tile.smem_pointwise <= 227 * 1024 - 2048
    assert tile.min_blocks * (tile.smem_pointwise + 1024) <= 228 * 1024
    min_regs = 96 if variant == 1 else 80
    assert tile.min_blocks * tile.threads * min_regs <= 65536 or tile.min_blocks == 1
    assert 1 <= tile.zc <= shape[0]


def test_fused_pairs_are_skipped_where_they_do_not_apply(monkeypatch):
    monkeypatch.delenv("PML_FUSE", raising=False)
    # odd contiguous extent (TMA rows start on 16-byte boundaries), 1-D meshes,
    # systems with algebraic components
    assert codegen.default_fused((21, 21), 1, 1, False) is None
    assert codegen.default_fused((101,), 1, 1, False) is None
    assert codegen.default_fused((64, 64, 64), 2, 1, True) is None
    monkeypatch.setenv("PML_FUSE", "0")
    assert codegen.default_fused((64, 64, 64), 3, 3, False) is None


def test_nvrtc_compiles_the_fused_pair_kernels(tmp_path, monkeypatch):
    """TMA (UTMALDG) and mbarrier (SYNCS) code paths build for sm_100a."""
    import subprocess

    monkeypatch.delenv("PML_FUSE", raising=False)
    monkeypatch.setenv("PML_SMALL", "0")
    eq = ns.BurgersEquation(3, 100.0)
    mesh = ns.Mesh([(0.0, 1.0)] * 3, [1.0 / 39, 1.0 / 47, 1.0 / 63])
    bc = ns.NeumannBoundaryCondition(lambda x, t: np.zeros((len(x), 3)), is_static=True)
    cp = ns.ConstrainedProblem(eq, mesh, [(bc, bc)] * 3)
    low = lower_problem(cp)
    spec = low.spec(**plan_overrides(cp, low, None))
    assert spec.fused is not None
    src = codegen.generate_source(spec)
    out = str(tmp_path / "fused.cubin")
    _native.compile_to_cubin(src, out)
    sass = subprocess.run(
        ["cuobjdump", "-sass", out], capture_output=True, text=True
    ).stdout
    assert "pml_fused_rk4_12" in sass and "pml_fused_rk4_34" in sass
    assert "UTMALDG" in sass and "SYNCS" in sass


def _burgers_problem(neumann_value=0.0, static=True):
    eq = ns.BurgersEquation(3, 100.0)
    mesh = ns.Mesh([(0.0, 1.0)] * 3, [1.0 / 15, 1.0 / 17, 1.0 / 19])
    bc = ns.NeumannBoundaryCondition(
        lambda x, t: np.full((len(x), 3), neumann_value), is_static=static
    )
    return ns.ConstrainedProblem(eq, mesh, [(bc, bc)] * 3)


def test_static_zero_flux_faces_are_compiled_in():
    """Faces whose static Neumann table is 0.0 everywhere need no table
    look-ups (PML_NEU_ZERO_MASK); any other value, or a dynamic condition,
    keeps them."""
    low = lower_problem(_burgers_problem(0.0))
    assert low.neu_mask == 63 and low.neu_zero_mask == 63
    assert "#define PML_NEU_ZERO_MASK 63" in codegen.generate_source(low.spec())
    assert lower_problem(_burgers_problem(0.25)).neu_zero_mask == 0
    assert lower_problem(_burgers_problem(0.0, static=False)).neu_zero_mask == 0
    # mixed: one Dirichlet face pair, the rest zero flux
    eq = ns.DiffusionEquation(2, 0.1)
    mesh = ns.Mesh([(0.0, 1.0)] * 2, [0.1, 0.1])
    dirichlet = ns.DirichletBoundaryCondition(
        lambda x, t: np.zeros((len(x), 1)), is_static=True)
    flux = ns.NeumannBoundaryCondition(
        lambda x, t: np.zeros((len(x), 1)), is_static=True)
    low = lower_problem(ns.ConstrainedProblem(eq, mesh, [(dirichlet, dirichlet), (flux, flux)]))
    assert low.neu_mask == 0b1100 and low.neu_zero_mask == 0b1100 and low.dir_mask == 0b0011


def test_interior_variant_folds_mesh_constants(monkeypatch):
    """The all-interior instantiation of the generated right-hand side enters
    first derivatives as constant x raw difference (shared products found by
    SymPy's cse) and the Cartesian Laplacian as one FMA chain; the boundary
    variants keep the plain primitives."""
    low = lower_problem(_burgers_problem())
    src = codegen.generate_source(low.spec())
    body = src[src.index("void pml_rhs_dt("):]
    body = body[: body.index("\n}\n")]
    fast, general = body.split("} else {")
    assert "if constexpr (IM == PML_IM_ALL)" in fast
    assert fast.count("pml_d1raw_at<") == 9 and "pml_d1_at<" not in fast
    assert re.search(r"const double X0 = PML_INV2H0\*L\d+;", fast)
    assert general.count("pml_d1_at<") == 9 and "pml_d1raw_at<" not in general
    assert fast.count("pml_lap_at<IM>") == 3 and general.count("pml_lap_at<IM>") == 3
    monkeypatch.setenv("PML_FOLD_D1", "0")
    monkeypatch.setenv("PML_FAST_LAPLACIAN", "0")
    plain = codegen.generate_source(low.spec())
    assert "pml_d1raw_at<0>(S" not in plain and "pml_lap_at<IM>" not in plain
    assert plain.count("pml_d2_at<") == 9
    # a polar mesh has metric factors in its gradients: only plain first
    # derivatives are folded, the Laplacian keeps its coordinate-system form
    _, polar = _source(cases.FDM_BY_NAME["shallow_water_polar_rk4"])
    assert "pml_lap_at<IM>" not in polar
