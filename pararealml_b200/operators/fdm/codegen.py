"""SymPy system -> CUDA C prelude for the fixed stage-kernel template.

Replaces the reference's per-symbol NumPy evaluators and ``sp.lambdify``
(``operators/symbol_mapper.py:160-253``, ``operators/fdm/fdm_symbol_mapper.py
:45-158``): every free symbol of the right-hand sides becomes an inline
expression over the template's stencil primitives (``pml_d1_at``,
``pml_d2_at``, ``pml_d2m_at``; csrc/fdm_template.cuh), with the coordinate
system algebra of ``numerical_differentiator.py:114-870`` written out, and the
right-hand sides themselves are printed with SymPy's C printer.  Nothing is
lambdified to NumPy.
"""
import hashlib
import os
import re
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import sympy as sp
from sympy.printing.c import C99CodePrinter

TEMPLATE_PATH = os.path.join(
    os.path.dirname(os.path.abspath(__file__)), "..", "..", "csrc",
    "fdm_template.cuh",
)

COORD_CODES = {"CARTESIAN": 0, "POLAR": 1, "CYLINDRICAL": 2, "SPHERICAL": 3}
KIND_CODES = {"D_Y_OVER_D_T": 0, "Y": 1, "Y_LAPLACIAN": 2}


def c_double(v: float) -> str:
    """Exact C literal of a double (hex float keeps every bit)."""
    v = float(v)
    if v != v:
        return "PML_NAN"
    if v in (float("inf"), float("-inf")):
        raise ValueError("infinite constant")
    return f"({v.hex()})"


class _CudaPrinter(C99CodePrinter):
    """C printer with exact constants and multiplication chains for small
    integer powers (``pow`` is two orders of magnitude slower in fp64)."""

    def __init__(self, names: Dict[str, str]):
        super().__init__({"allow_unknown_functions": False})
        self._names = names

    def _print_Symbol(self, expr):
        return self._names[expr.name]

    def _print_Float(self, expr):
        return c_double(float(expr))

    def _print_Integer(self, expr):
        return f"{int(expr)}.0"

    def _print_Rational(self, expr):
        return f"({int(expr.p)}.0/{int(expr.q)}.0)"

    def _print_Pow(self, expr):
        base, exp = expr.base, expr.exp
        if exp.is_Integer and 1 <= abs(int(exp)) <= 8:
            b = self._print(base)
            chain = "*".join([f"({b})"] * abs(int(exp)))
            return f"({chain})" if int(exp) > 0 else f"(1.0/({chain}))"
        if exp == sp.Rational(1, 2):
            return f"sqrt({self._print(base)})"
        if exp == sp.Rational(-1, 2):
            return f"(1.0/sqrt({self._print(base)}))"
        return f"pow({self._print(base)}, {self._print(exp)})"


@dataclass
class ProblemSpec:
    """Everything the generated kernels are specialised on."""

    shape: Tuple[int, ...]  # mesh vertices shape, () for an ODE
    d_x: Tuple[float, ...]
    coord: str  # CARTESIAN | POLAR | CYLINDRICAL | SPHERICAL
    y_dim: int
    rhs: Sequence[sp.Expr]
    kinds: Sequence[str]  # LHS kind name per equation
    neu_mask: int = 0  # bit (axis * 2 + side): a Neumann table exists
    dir_mask: int = 0
    neu_zero_mask: int = 0  # ... and holds 0.0 everywhere (static zero flux)
    passthrough: bool = False
    coherent_loads: bool = False
    eval_only: bool = False  # the differentiator entry points
    block: Optional[Tuple[int, int, int]] = None
    #: tile of the fused stage-pair kernels, or None to generate none
    fused: Optional["FusedTile"] = None
    #: threads of the single-block time-loop kernel for small meshes (0: none)
    small_threads: int = 0
    #: cells along axis 0 every thread of a stage kernel walks
    zrep: int = 1


def default_block(shape) -> Tuple[int, int, int]:
    nd = len(shape)
    if nd <= 1:
        return (256, 1, 1)
    if nd == 2:
        return (128, 2, 1) if shape[1] >= 128 else (32, 8, 1)
    # measured on B200 (512^3 Burgers RK4): small row-shaped blocks win
    return (32, 4, 1) if shape[2] >= 32 else (16, 4, 4)


N_SMS = 148
#: cells along axis 0 per thread of the Jacobi sweep kernel (2-D / 3-D meshes)
JACOBI_REP = 4
SMALL_MESH_CELLS = 8192


def default_zrep(shape) -> int:
    """Cells along axis 0 per thread in the stage kernels (2-D / 3-D meshes)."""
    if len(shape) < 2:
        return 1
    if os.environ.get("PML_ZREP"):
        return max(1, int(os.environ["PML_ZREP"]))
    # measured on B200 (512^3 Burgers RK4): 1, 2, 4, 8 are within 1 %
    return 1


def default_small(shape) -> int:
    """Meshes of at most SMALL_MESH_CELLS cells (and ODE systems) run their
    whole time loop in one thread block."""
    if os.environ.get("PML_SMALL", "1") == "0":
        return 0
    cells = int(np.prod(shape)) if len(shape) else 1
    if cells > SMALL_MESH_CELLS:
        return 0
    return int(min(1024, max(32, 32 * -(-cells // 32))))


@dataclass(frozen=True)
class FusedTile:
    """Geometry of the fused stage-pair kernels (csrc/fdm_template.cuh)."""

    tx: int  # tile cells along the contiguous mesh axis
    ty: int  # tile cells along axis 1 (3-D meshes; 1 otherwise)
    zc: int  # planes of the marching axis per thread block
    depth: int  # iterations the TMA copies run ahead
    threads: int  # one thread per cell of the tile + halo 1
    smem_first: int  # dynamic shared memory of the stage 1+2 / midpoint kernels
    smem_pointwise: int  # ... of the stage 3+4 kernel (adds y and acc rings)
    min_blocks: int
    #: 1 = one thread per cell of the stage-A tile, every stencil operand read
    #: from shared memory; 2 = column-marching: a thread owns ``rows`` rows of
    #: one column and keeps three planes of both stages' inputs in registers
    variant: int = 1
    rows: int = 1
    #: marching variant: 0 = __syncthreads per plane, 1 = per-warp mbarrier
    #: arrivals with one plane of slack (rings one slot deeper)
    sync: int = 0


SMEM_PER_BLOCK_MAX = 227 * 1024
SMEM_PER_SM = 228 * 1024


def default_fused(shape, y_dim, n_dt=None, passthrough=False) -> Optional[FusedTile]:
    """Tile of the fused stage-pair kernels: the stage-A tile (tile + halo 1)
    has one thread per cell, rows are moved by the TMA unit and therefore have
    to start on 16-byte boundaries (even extent of the contiguous axis)."""
    mode = os.environ.get("PML_FUSE", "auto")
    if mode == "0":
        return None
    nd = len(shape)
    n_dt = y_dim if n_dt is None else n_dt
    if nd < 2 or any(n < 3 for n in shape) or n_dt < 1 or shape[-1] % 2:
        return None
    if any(n > 65534 for n in shape[1:]):
        # the pair kernels pack the in-plane coordinates into 16-bit fields
        return None
    # by default only systems whose components are all time-stepped: algebraic
    # (LHS.Y) and Poisson components are stencil inputs that never change
    # within a step, which the unfused kernels serve from L2 at no extra cost
    # (measured: Cahn-Hilliard 256^3 runs 1.8x faster unfused)
    if mode != "1" and n_dt != y_dim:
        return None
    variant = int(os.environ.get("PML_FVARIANT", "2"))
    if variant == 2:
        tile = _marching_tile(shape, y_dim, n_dt, passthrough)
        if tile is not None:
            return tile
    if os.environ.get("PML_FTILE"):
        tx, ty = (int(v) for v in os.environ["PML_FTILE"].split(","))
    elif nd == 3:
        # tile + halo 1 = 32 cells per row: one warp per row of the stage-A
        # tile (conflict-free shared-memory rows, warp-uniform row predicates);
        # 6 rows + prefetch depth 2: two thread blocks of 256 threads share an
        # SM with 128 registers per thread, so one computes while the other
        # waits at its per-plane barrier (measured on B200, 512^3 Burgers RK4:
        # 7.6 ms/step; depth 1 7.9, 30x8 at 96 registers 8.0-10.1, 30x16 with
        # one block per SM 8.6-9.9)
        tx, ty = 30, 6
    else:
        # measured on B200 (4096^2 polar shallow water RK4): 126 and 94 give
        # 25.1 Gcell-steps/s, 222 gives 22.8, 446 gives 20.5
        tx, ty = 126, 1
    if nd == 2:
        ty = 1
    tx = max(2, min(tx, shape[-1] + (shape[-1] % 2)))
    tx -= tx % 2
    if nd == 3:
        ty = max(1, min(ty, shape[1]))
    depth = int(os.environ.get("PML_FDEPTH", "2"))
    n_ring = n_dt if passthrough else y_dim
    hy = 1 if nd == 3 else 0

    def pad16(n):
        return -(-n // 16) * 16

    def geometry(tx, ty, depth):
        mw, mh = tx + 2, ty + 2 * hy
        iw, ih = tx + 4, ty + 4 * hy
        threads = 32 * -(-(mw * mh) // 32)
        # component planes of the TMA-fed rings are padded to 128 bytes
        first = 8 * ((depth + 3) * n_ring * pad16(iw * ih) + pad16(4 * n_ring * mw * mh))
        pointwise = first + 8 * (depth + 1) * n_dt * (
            pad16(iw * mh) + pad16(tx * ty)
        )
        return threads, first, pointwise

    while True:
        threads, first, pointwise = geometry(tx, ty, depth)
        # a TMA box is at most 256 elements along each dimension
        if (threads <= 1024 and tx + 4 <= 256
                and pointwise + 2048 <= SMEM_PER_BLOCK_MAX):
            break
        if depth > 1:
            depth -= 1
        elif nd == 3 and ty > 2:
            ty //= 2
        elif tx > 8:
            tx = (tx // 2) - ((tx // 2) % 2)
        else:
            return None
    tiles = -(-shape[-1] // tx) * (-(-shape[1] // ty) if nd == 3 else 1)
    if os.environ.get("PML_FZC"):
        zc = int(os.environ["PML_FZC"])
    else:
        # enough thread blocks for ~8 waves, at least 32 planes per block so
        # that the two extra planes a chunk recomputes stay cheap
        per_sm = max(1, min(SMEM_PER_SM // (pointwise + 1024), 2048 // threads))
        chunks = max(1, -(-(8 * N_SMS * per_sm) // tiles))
        # (2048^2 diffusion: 16 rows per chunk 47.7 Gcell-steps/s, 32: 43.7,
        # 64: 31.1 -- small meshes need the extra thread blocks)
        zc = min(64, max(32 if nd == 3 else 16, -(-shape[0] // chunks)))
    zc = max(1, min(zc, shape[0]))
    # resident blocks per SM by shared memory and threads, capped so that the
    # compiler keeps ~96 registers per thread (rotating stage-A results)
    per_sm = max(1, min(SMEM_PER_SM // (pointwise + 1024), 2048 // threads,
                        65536 // (96 * threads)))
    min_blocks = int(os.environ.get("PML_FMIN_BLOCKS", str(per_sm)))
    return FusedTile(tx, ty, zc, depth, threads, first, pointwise, min_blocks)


def _marching_tile(shape, y_dim, n_dt, passthrough) -> Optional[FusedTile]:
    """Geometry of the column-marching stage-pair kernels: rows of the
    stage-A tile (tile + halo 1) are whole warps, so the tile is 32 k - 2
    cells wide; a thread owns ``rows`` consecutive rows of one column (3-D
    meshes), so the stage-A tile is ``rows * row groups`` rows high."""
    nd = len(shape)
    hy = 1 if nd == 3 else 0
    n_ring = n_dt if passthrough else y_dim
    depth = int(os.environ.get("PML_FDEPTH", "2"))
    # rows per thread (3-D).  Measured on B200, 512^3 Burgers RK4, ms per step
    # (tile, warps per SM): 1 row 30x14, 16 warps: 5.38; 1 row 30x6, 2 x 8
    # warps: 5.67; 2 rows 30x14, 8 warps at 220 registers: 5.85; 1 row 30x16,
    # 18 warps at 112 registers: 6.48; 2 rows 30x18, 10 warps: 6.80
    rows = int(os.environ.get("PML_FROWS", "1")) if nd == 3 else 1
    sync = int(os.environ.get("PML_FSYNC", "0"))

    def pad16(n):
        return -(-n // 16) * 16

    if os.environ.get("PML_FTILE"):
        tx, ty = (int(v) for v in os.environ["PML_FTILE"].split(","))
    elif nd == 3:
        # 16 warps of 32 columns at one row per thread, 8 at two
        tx, ty = 30, (16 if rows == 1 else 8 * rows) - 2
    else:
        tx, ty = 126, 1
    if nd == 2:
        ty = 1
    # the tile is not wider than the mesh rounded up to whole warps
    tx = min(tx, 32 * -(-(shape[-1] + 2) // 32) - 2)
    tx = max(30, 32 * ((tx + 2) // 32) - 2)
    if nd == 3:
        # whole row groups, not (much) higher than the mesh
        ty = max(rows, min(ty, rows * -(-(shape[1] + 2) // rows)))
        ty = rows * -(-(ty + 2) // rows) - 2
        if ty < 1:
            return None

    def geometry(tx, ty, depth):
        mw, mh = tx + 2, ty + 2 * hy
        iw, ih = tx + 4, ty + 4 * hy
        threads = 32 * (mw // 32) * (mh // rows)
        in_slot = n_ring * pad16(iw * ih)
        mid_slot = n_ring * mw * mh
        y_slot = n_dt * pad16(iw * mh)
        acc_slot = n_dt * pad16(tx * ty)
        pad = mw + 8  # unguarded (interior) rows read one row past the ring
        first = 8 * ((depth + 4 + sync) * in_slot + 4 * mid_slot + pad)
        pointwise = 8 * ((depth + 3 + sync) * (in_slot + y_slot)
                         + (depth + 1 + sync) * acc_slot + 4 * mid_slot + pad)
        return threads, first, pointwise

    while True:
        threads, first, pointwise = geometry(tx, ty, depth)
        if (threads <= 1024 and tx + 4 <= 256
                and max(first, pointwise) + 1024 <= SMEM_PER_BLOCK_MAX):
            break
        if depth > 1:
            depth -= 1
        elif nd == 3 and ty + 2 > 2 * rows:
            ty -= rows
        elif tx > 30:
            tx -= 32
        else:
            return None
    tiles = -(-shape[-1] // tx) * (-(-shape[1] // ty) if nd == 3 else 1)
    # resident blocks per SM: shared memory, threads, and registers (a thread
    # holds 2 x 3 planes of its rows' components plus stage A's increments)
    regs = 64 + 2 * rows * (6 * n_ring + 3 * n_dt)
    per_sm = max(1, min(SMEM_PER_SM // (max(first, pointwise) + 1024),
                        2048 // threads, 65536 // (min(regs, 255) * threads)))
    if os.environ.get("PML_FZC"):
        zc = int(os.environ["PML_FZC"])
    else:
        # a chunk recomputes two planes of stage A: long chunks, but enough
        # thread blocks for ~6 waves
        chunks = max(1, -(-(6 * N_SMS * per_sm) // tiles))
        zc = min(128, max(32 if nd == 3 else 16, -(-shape[0] // chunks)))
    zc = max(1, min(zc, shape[0]))
    min_blocks = int(os.environ.get("PML_FMIN_BLOCKS", str(per_sm)))
    return FusedTile(tx, ty, zc, depth, threads, first, pointwise, min_blocks,
                     variant=2, rows=rows, sync=sync)


class _LeafBuilder:
    """Turns symbol names into C expressions and records the stencil
    primitives they need (emitted once each)."""

    def __init__(self, spec: ProblemSpec):
        self.spec = spec
        self.nd = len(spec.shape)
        self.cs = spec.coord
        self.prims: Dict[str, str] = {}
        self.order: List[str] = []

    def _prim(self, name: str, code: str) -> str:
        if name not in self.prims:
            self.prims[name] = code
            self.order.append(name)
        return name

    # primitives ----------------------------------------------------------
    def Y(self, c):
        return self._prim(f"Y{c}", f"S.template rel<0, 0, 0>({c}, c)")

    def D1(self, c, a):
        return self._prim(
            f"D1_{c}_{a}",
            f"pml_d1_at<{a}, IM>(a, S, {c}, c)",
        )

    def D2(self, c, a, b=None):
        if b is None or a == b:
            return self._prim(
                f"D2_{c}_{a}",
                f"pml_d2_at<{a}, IM>(a, S, {c}, c)",
            )
        return self._prim(
            f"D2M_{c}_{a}_{b}",
            f"pml_d2m_at<{a}, {b}, IM>(a, S, {c}, c)",
        )

    def X(self, a):
        return self._prim(f"X{a}", f"__ldg(a.coord[{a}] + c.i{a})")

    def IR(self):
        return self._prim("IR", "__ldg(a.aux[0] + c.i0)")

    def SIN(self):
        return self._prim("SINP", "__ldg(a.aux[1] + c.i2)")

    def COS(self):
        return self._prim("COSP", "__ldg(a.aux[2] + c.i2)")

    def IS(self):
        return self._prim("ISINP", "__ldg(a.aux[3] + c.i2)")

    def IRS(self):
        return self._prim("IRS", f"({self.IR()} * {self.IS()})")

    # leaves (numerical_differentiator.py:114-870) --------------------------
    def _check_axis(self, a):
        if not 0 <= a < self.nd:
            raise ValueError(f"x axis {a} out of range for {self.nd} dims")

    def gradient(self, c, a):
        self._check_axis(a)
        d = self.D1(c, a)
        if self.cs == "CARTESIAN":
            return d
        if self.cs == "SPHERICAL":
            if a == 0:
                return d
            return f"({d} * {self.IRS()})" if a == 1 else f"({d} * {self.IR()})"
        return f"({d} * {self.IR()})" if a == 1 else d

    def hessian(self, c, a, b):
        self._check_axis(a)
        self._check_axis(b)
        d2 = self.D2(c, a, b)
        cs = self.cs
        if cs == "CARTESIAN":
            return d2
        ir = self.IR()
        axes = {a, b}
        if cs == "SPHERICAL":
            if a == 0 and b == 0:
                return d2
            if a == 1 and b == 1:
                return (
                    f"(({self.D1(c, 0)} + ({d2} * {self.IS()} + {self.COS()} * "
                    f"{self.D1(c, 2)}) * {self.IRS()}) * {ir})"
                )
            if a == 2 and b == 2:
                return f"(({d2} * {ir} + {self.D1(c, 0)}) * {ir})"
            if axes == {0, 1}:
                return f"(({d2} - {self.D1(c, 1)} * {ir}) * {self.IRS()})"
            if axes == {0, 2}:
                return f"(({d2} - {self.D1(c, 2)} * {ir}) * {ir})"
            return (
                f"(({self.SIN()} * {d2} - {self.COS()} * {self.D1(c, 1)}) * "
                f"({self.IRS()} * {self.IRS()}))"
            )
        if a != 1 and b != 1:
            return d2
        if a == 1 and b == 1:
            return f"(({d2} * {ir} + {self.D1(c, 0)}) * {ir})"
        if axes == {0, 1}:
            return f"(({d2} - {self.D1(c, 1)} * {ir}) * {ir})"
        return f"({d2} * {ir})"

    def divergence(self, comps):
        if len(comps) != self.nd:
            raise ValueError("divergence needs one component per axis")
        cs = self.cs
        if cs == "CARTESIAN":
            return "(" + " + ".join(self.D1(c, i) for i, c in enumerate(comps)) + ")"
        ir = self.IR()
        if cs == "SPHERICAL":
            return (
                f"({self.D1(comps[0], 0)} + ({self.D1(comps[2], 2)} + 2.0 * "
                f"{self.Y(comps[0])} + ({self.D1(comps[1], 1)} + {self.COS()} * "
                f"{self.Y(comps[2])}) * {self.IS()}) * {ir})"
            )
        out = (
            f"({self.D1(comps[0], 0)} + ({self.Y(comps[0])} + "
            f"{self.D1(comps[1], 1)}) * {ir})"
        )
        if cs == "CYLINDRICAL":
            out = f"({out} + {self.D1(comps[2], 2)})"
        return out

    def curl(self, comps, ind):
        nd = self.nd
        if not 2 <= nd <= 3:
            raise ValueError("curl needs 2 or 3 spatial dimensions")
        if len(comps) != nd:
            raise ValueError("curl needs one component per axis")
        if nd == 2 and ind != 0:
            raise ValueError("2D curl only has component 0")
        if not 0 <= ind < nd:
            raise ValueError(f"curl index {ind} out of range")
        v = comps
        cs = self.cs
        if cs == "CARTESIAN":
            if nd == 2 or ind == 2:
                return f"({self.D1(v[1], 0)} - {self.D1(v[0], 1)})"
            if ind == 0:
                return f"({self.D1(v[2], 1)} - {self.D1(v[1], 2)})"
            return f"({self.D1(v[0], 2)} - {self.D1(v[2], 0)})"
        ir = self.IR()
        if cs == "SPHERICAL":
            if ind == 0:
                return (
                    f"(({self.D1(v[1], 2)} + ({self.COS()} * {self.Y(v[1])} - "
                    f"{self.D1(v[2], 1)}) * {self.IS()}) * {ir})"
                )
            if ind == 1:
                return (
                    f"({self.D1(v[2], 0)} + ({self.Y(v[2])} - "
                    f"{self.D1(v[0], 2)}) * {ir})"
                )
            return (
                f"(-{self.D1(v[1], 0)} + ({self.D1(v[0], 1)} * {self.IS()} - "
                f"{self.Y(v[1])}) * {ir})"
            )
        if cs == "POLAR" or ind == 2:
            return (
                f"({self.D1(v[1], 0)} + ({self.Y(v[1])} - {self.D1(v[0], 1)}) "
                f"* {ir})"
            )
        if ind == 0:
            return f"({self.D1(v[2], 1)} * {ir} - {self.D1(v[1], 2)})"
        return f"({self.D1(v[0], 2)} - {self.D1(v[2], 0)})"

    def laplacian(self, c):
        cs = self.cs
        if cs == "CARTESIAN":
            if os.environ.get("PML_FAST_LAPLACIAN", "1") != "0":
                # one FMA chain over the axes (fdm_template.cuh::pml_lap_at)
                return self._prim(f"LAP_{c}", f"pml_lap_at<IM>(a, S, {c}, c)")
            return "(" + " + ".join(self.D2(c, a) for a in range(self.nd)) + ")"
        ir = self.IR()
        if cs == "SPHERICAL":
            return (
                f"({self.D2(c, 0)} + (2.0 * {self.D1(c, 0)} + ({self.D2(c, 2)} "
                f"+ ({self.COS()} * {self.D1(c, 2)} + {self.D2(c, 1)} * "
                f"{self.IS()}) * {self.IS()}) * {ir}) * {ir})"
            )
        out = (
            f"({self.D2(c, 0)} + ({self.D2(c, 1)} * {ir} + {self.D1(c, 0)}) "
            f"* {ir})"
        )
        if cs == "CYLINDRICAL":
            out = f"({out} + {self.D2(c, 2)})"
        return out

    def vector_laplacian(self, comps, ind):
        if len(comps) != self.nd:
            raise ValueError("vector Laplacian needs one component per axis")
        if not 0 <= ind < self.nd:
            raise ValueError(f"vector Laplacian index {ind} out of range")
        v = comps
        lap = self.laplacian(v[ind])
        cs = self.cs
        if cs == "CARTESIAN":
            return lap
        ir2 = f"({self.IR()} * {self.IR()})"
        if cs == "SPHERICAL":
            if ind == 1:
                return (
                    f"({lap} - 2.0 * ({self.Y(v[0])} + {self.D1(v[2], 2)} + "
                    f"({self.COS()} * {self.Y(v[2])} + {self.D1(v[1], 1)}) * "
                    f"{self.IS()}) * {ir2})"
                )
            if ind == 2:
                return (
                    f"({lap} + 2.0 * ({self.D1(v[0], 1)} + ({self.COS()} * "
                    f"{self.D1(v[2], 1)} - {self.Y(v[1])} / 2.0) * {self.IS()}) "
                    f"* ({self.IS()} * {ir2}))"
                )
            return (
                f"({lap} + 2.0 * ({self.D1(v[0], 2)} - ({self.Y(v[2])} / 2.0 + "
                f"{self.COS()} * {self.D1(v[1], 1)}) * ({self.IS()} * "
                f"{self.IS()})) * {ir2})"
            )
        if ind == 0:
            return (
                f"({lap} - ({self.Y(v[0])} + 2.0 * {self.D1(v[1], 1)}) * {ir2})"
            )
        if ind == 1:
            return (
                f"({lap} - ({self.Y(v[1])} - 2.0 * {self.D1(v[0], 1)}) * {ir2})"
            )
        return lap

    # symbol name -> expression --------------------------------------------
    def leaf(self, name: str) -> str:
        tokens = name.split("_")
        kind = tokens[0]
        idx = [int(s) for s in tokens[1:]]
        if kind == "t":
            return "t"
        if kind == "y":
            return self.Y(idx[0])
        if kind == "x":
            self._check_axis(idx[0])
            return self.X(idx[0])
        if kind == "y-gradient":
            return self.gradient(*idx)
        if kind == "y-hessian":
            return self.hessian(*idx)
        if kind == "y-laplacian":
            return self.laplacian(idx[0])
        if kind == "y-divergence":
            return self.divergence(idx)
        if kind == "y-curl":
            if self.nd == 2:
                return self.curl(idx, 0)
            return self.curl(idx[:-1], idx[-1])
        if kind == "y-vector-laplacian":
            return self.vector_laplacian(idx[:-1], idx[-1])
        raise KeyError(name)


def _emit_body(spec: ProblemSpec, exprs, fold_first_derivatives: bool) -> str:
    """Statements that evaluate ``exprs`` into ``out[]``.  With
    ``fold_first_derivatives`` (only valid where no boundary handling applies)
    a leaf that is a plain first derivative ``(hi - lo) / (2 h)`` enters the
    expressions as the product of the mesh constant and the raw difference,
    and SymPy's common-subexpression pass then shares products such as
    ``u_a / (2 h_a)`` between the equations (Burgers: 6 fp64 operations fewer
    per cell; rounding differs from the unfolded form by ~1 ulp per product)."""
    builder = _LeafBuilder(spec)
    symbols = sorted(
        set().union(*[e.free_symbols for e in exprs]) if exprs else set(),
        key=lambda s: s.name,
    )
    names, leaf_lines, raw_lines, subs = {}, [], [], {}
    for n, s in enumerate(symbols):
        ident = f"L{n}"
        leaf = builder.leaf(s.name)
        m = re.fullmatch(r"D1_(\d+)_(\d+)", leaf) if fold_first_derivatives else None
        if m:
            comp, axis = int(m.group(1)), int(m.group(2))
            raw, const = f"RAW_{comp}_{axis}", f"PML_INV2H{axis}"
            raw_lines.append(
                f"  const double {raw} = pml_d1raw_at<{axis}>(S, {comp}, c);"
                f"  // {s.name} * 2 h"
            )
            names[raw], names[const] = raw, const
            subs[s] = sp.Symbol(const) * sp.Symbol(raw)
            continue
        names[s.name] = ident
        leaf_lines.append(f"  const double {ident} = {leaf};  // {s.name}")
    cse_lines = []
    if subs:
        folded = [e.xreplace(subs) for e in exprs]
        temps, exprs = sp.cse(folded, symbols=sp.numbered_symbols("X"))
        for t, _ in temps:
            names[t.name] = t.name
        printer = _CudaPrinter(names)
        cse_lines = [
            f"  const double {t.name} = {printer.doprint(e)};" for t, e in temps
        ]
    printer = _CudaPrinter(names)
    out_lines = [
        f"  out[{j}] = {printer.doprint(e)};" for j, e in enumerate(exprs)
    ]
    # only the primitives something still refers to (a folded first derivative
    # leaves its D1 primitive unused)
    text = "\n".join(leaf_lines + cse_lines + out_lines)
    prim_lines, needed = [], set()
    for p_name in reversed(builder.order):
        code = builder.prims[p_name]
        if re.search(rf"\b{p_name}\b", text) or p_name in needed:
            prim_lines.append(f"  const double {p_name} = {code};")
            needed.update(re.findall(r"\b[A-Z][A-Z0-9]*(?:_\d+)*\b", code))
    prim_lines.reverse()
    return "\n".join(prim_lines + raw_lines + leaf_lines + cse_lines + out_lines)


def _emit_function(fn_name: str, spec: ProblemSpec, eq_indices: Sequence[int]):
    exprs = [sp.sympify(spec.rhs[i]) for i in eq_indices]
    body = _emit_body(spec, exprs, False)
    if os.environ.get("PML_FOLD_D1", "1") != "0" and spec.shape:
        fast = _emit_body(spec, exprs, True)
        if "pml_d1raw_at<" in fast:
            # the all-interior instantiation has no boundary handling: its
            # first derivatives are plain differences
            body = (
                "  if constexpr (IM == PML_IM_ALL) {\n" + fast
                + "\n  } else {\n" + body + "\n  }"
            )
    return (
        "template <int IM, class SRC>\n"
        f"__device__ __forceinline__ void {fn_name}(const PmlArgs& a, "
        "const SRC& S, const PmlCell& c, double t, double* out) {\n"
        "  (void)a; (void)S; (void)c; (void)t; (void)out;\n"
        f"{body}\n}}\n"
    )


def _int_list(values) -> str:
    values = list(values) or [0]
    return "{" + ", ".join(str(int(v)) for v in values) + "}"


def generate_source(spec: ProblemSpec) -> str:
    nd = len(spec.shape)
    if nd > 3:
        raise NotImplementedError(
            "the B200 FDM kernels support at most 3 spatial dimensions"
        )
    if any(n < 3 for n in spec.shape):
        # numerical_differentiator.py:1021-1024
        raise ValueError("y must contain at least 3 points along every x-axis")
    n = list(spec.shape) + [1] * (3 - nd)
    h = list(spec.d_x) + [1.0] * (3 - nd)
    kinds = [KIND_CODES[k] for k in spec.kinds]
    dt_idx = [i for i, k in enumerate(kinds) if k == 0]
    alg_idx = [i for i, k in enumerate(kinds) if k == 1]
    lap_idx = [i for i, k in enumerate(kinds) if k == 2]
    block = spec.block or default_block(spec.shape)
    if os.environ.get("PML_BLOCK"):
        block = tuple(int(v) for v in os.environ["PML_BLOCK"].split(","))

    threads = block[0] * block[1] * block[2]
    # cap registers at 64 per thread (16 warps per SM sub-partition budget):
    # measured +12 % on the 512^3 Burgers RK4 stages versus the uncapped 70
    min_blocks = int(
        os.environ.get("PML_MIN_BLOCKS", str(max(1, 1024 // threads)))
    )
    fused = spec.fused
    lines = [
        "// generated by pararealml_b200/operators/fdm/codegen.py",
        f"#define PML_NDIM {nd}",
        f"#define PML_C {spec.y_dim}",
        f"#define PML_N0 {n[0]}",
        f"#define PML_N1 {n[1]}",
        f"#define PML_N2 {n[2]}",
        f"#define PML_COORD {COORD_CODES[spec.coord]}",
        f"#define PML_NEU_MASK {spec.neu_mask}",
        f"#define PML_DIR_MASK {spec.dir_mask}",
        f"#define PML_NEU_ZERO_MASK {spec.neu_zero_mask & spec.neu_mask}",
        f"#define PML_PASSTHROUGH {int(spec.passthrough)}",
        f"#define PML_COHERENT_LOADS {int(spec.coherent_loads or spec.small_threads > 0)}",
        f"#define PML_SMALL {int(spec.small_threads > 0)}",
        f"#define PML_SMALL_THREADS {max(spec.small_threads, 32)}",
        f"#define PML_NDT {len(dt_idx)}",
        f"#define PML_NALG {len(alg_idx)}",
        f"#define PML_NLAP {len(lap_idx)}",
        f"#define PML_STREAMING {int(os.environ.get('PML_STREAM', '0'))}",
        f"#define PML_MIN_BLOCKS {min_blocks}",
        f"#define PML_FUSED {fused.variant if fused else 0}",
        f"#define PML_FTX {fused.tx if fused else 32}",
        f"#define PML_FTY {fused.ty if fused else 1}",
        f"#define PML_FZC {fused.zc if fused else 1}",
        f"#define PML_FDEPTH {fused.depth if fused else 1}",
        f"#define PML_F_THREADS {fused.threads if fused else 32}",
        f"#define PML_FMIN_BLOCKS {fused.min_blocks if fused else 1}",
        f"#define PML_FROWS {fused.rows if fused else 1}",
        f"#define PML_FSYNC {fused.sync if fused else 0}",
        f"#define PML_F_PATH1 {int(os.environ.get('PML_FPATH1', '1'))}",
        f"#define PML_ZREP {max(1, int(spec.zrep))}",
        f"#define PML_JREP {JACOBI_REP}",
        f"#define PML_BX {block[0]}",
        f"#define PML_BY {block[1]}",
        f"#define PML_BZ {block[2]}",
    ]
    for a in range(3):
        lines.append(f"#define PML_H{a} {c_double(h[a])}")
        lines.append(f"#define PML_INV2H{a} {c_double(1.0 / (2.0 * h[a]))}")
        lines.append(f"#define PML_INVHH{a} {c_double(1.0 / (h[a] * h[a]))}")
    inv_diag = 1.0 / float((2.0 / np.square(np.array(spec.d_x))).sum()) if nd else 1.0
    lines.append(f"#define PML_JAC_INV_DIAG {c_double(inv_diag)}")
    lap_diag = -2.0 * float(sum(1.0 / (float(v) * float(v)) for v in spec.d_x)) if nd else 0.0
    lines.append(f"#define PML_LAP_DIAG {c_double(lap_diag)}")
    lines.append(
        f"static __device__ constexpr int PML_KIND[] = {_int_list(kinds)};"
    )
    lines.append(
        f"static __device__ constexpr int PML_DT_IDX[] = {_int_list(dt_idx)};"
    )
    lines.append(
        f"static __device__ constexpr int PML_ALG_IDX[] = {_int_list(alg_idx)};"
    )
    lines.append(
        f"static __device__ constexpr int PML_LAP_IDX[] = {_int_list(lap_idx)};"
    )
    prelude = "\n".join(lines) + "\n"

    rhs_code = _emit_function("pml_rhs_dt", spec, dt_idx) + _emit_function(
        "pml_rhs_aux", spec, alg_idx + lap_idx
    )
    with open(TEMPLATE_PATH) as fh:
        template = fh.read()
    # mixed second derivatives read the in-plane neighbours of the planes
    # above and below (the marching stage-pair kernels wait accordingly)
    prelude = prelude.replace(
        "#define PML_FUSED ",
        f"#define PML_MIXED {int('pml_d2m_at<' in rhs_code)}\n#define PML_FUSED ", 1)
    return prelude + template.replace("PML_GENERATED_RHS", rhs_code)


def source_key(source: str) -> str:
    return hashlib.sha256(source.encode()).hexdigest()[:24]
