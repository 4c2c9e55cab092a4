"""Device plans: generated kernels + tables + scratch, driven through the
C ABI (include/pararealml_b200.h).  PyTorch only provides device memory and
streams here.
"""
import ctypes
import os
import warnings
from typing import Dict, Optional

import numpy as np
import torch

from pararealml_b200 import _native
from pararealml_b200.operators.fdm import codegen
from pararealml_b200.operators.fdm.lowering import LoweredProblem

CACHE_DIR = os.environ.get(
    "PML_KERNEL_CACHE",
    os.path.join(
        os.path.dirname(os.path.abspath(__file__)), "..", "..", "_kcache"
    ),
)

_PLANS: Dict[str, "DevicePlan"] = {}
_TOTAL_LAUNCHES = 0


def _from_numpy(array: np.ndarray) -> torch.Tensor:
    """torch view of a (possibly read-only) host array; only ever read."""
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", UserWarning)
        return torch.from_numpy(np.ascontiguousarray(array))


def require_cuda() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError(
            "pararealml_b200 operators need a CUDA device (sm_100a); there is "
            "no CPU fallback"
        )
    return torch.device("cuda", torch.cuda.current_device())


def stream_ptr() -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def cubin_path(source: str) -> str:
    os.makedirs(CACHE_DIR, exist_ok=True)
    return os.path.join(CACHE_DIR, codegen.source_key(source) + ".cubin")


def count_launch(n: int = 1):
    global _TOTAL_LAUNCHES
    _TOTAL_LAUNCHES += n


def to_device(array: np.ndarray, device) -> torch.Tensor:
    return _from_numpy(np.asarray(array, dtype=np.float64)).to(device)


def ic_mesh(mesh, device):
    """``pml_ic_mesh`` of a mesh (axis coordinate vectors and, for curvilinear
    meshes, the host-evaluated trigonometric factors of
    ``mesh.py:to_cartesian_coordinates``) plus the tensors it points at."""
    from pararealml_b200.operators.fdm.codegen import COORD_CODES

    out = _native.IcMesh()
    axes = mesh.vertex_axis_coordinates
    out.n_dims = len(axes)
    out.coord = COORD_CODES[mesh.coordinate_system_type.name]
    keep = []
    for a in range(3):
        out.shape[a] = len(axes[a]) if a < len(axes) else 1
        if a < len(axes):
            t = to_device(axes[a], device)
            keep.append(t)
            out.axis_dev[a] = t.data_ptr()
    if out.coord != 0:
        trig = [np.cos(axes[1]), np.sin(axes[1])]
        if out.coord == 3:
            trig += [np.sin(axes[2]), np.cos(axes[2])]
        for i, v in enumerate(trig):
            t = to_device(v, device)
            keep.append(t)
            out.trig_dev[i] = t.data_ptr()
    return ctypes.byref(out), (out, keep)


def total_launches() -> int:
    """Kernels launched by this process through plans (bench accounting)."""
    return _TOTAL_LAUNCHES + sum(p.launches for p in _PLANS.values())


class DevicePlan:
    """One compiled kernel set for a (problem spec) on the current device."""

    def __init__(self, low: LoweredProblem, spec: codegen.ProblemSpec):
        self.device = require_cuda()
        self.low = low
        self.spec = spec
        self.source = codegen.generate_source(spec)
        block = spec.block or codegen.default_block(spec.shape)
        if os.environ.get("PML_BLOCK"):
            block = tuple(int(v) for v in os.environ["PML_BLOCK"].split(","))
        desc = _native.PlanDesc()
        desc.n_dims = low.n_dims
        shape3 = list(low.shape) + [1] * (3 - low.n_dims)
        for i in range(3):
            desc.shape[i] = shape3[i]
            desc.block[i] = block[i]
        fused = spec.fused
        desc.fused = fused.variant if fused is not None else 0
        if fused is not None:
            desc.fused_tile[0], desc.fused_tile[1] = fused.tx, fused.ty
            desc.fused_zc = fused.zc
            desc.fused_threads = fused.threads
            desc.fused_smem[0] = fused.smem_first
            desc.fused_smem[1] = fused.smem_pointwise
        self.fused = fused
        desc.small_threads = int(spec.small_threads)
        desc.zrep = max(1, int(spec.zrep))
        desc.y_dim = low.y_dim
        desc.n_dt = len(low.kind_indices("D_Y_OVER_D_T"))
        desc.n_alg = len(low.kind_indices("Y"))
        desc.n_lap = len(low.kind_indices("Y_LAPLACIAN"))
        self.n_lap = desc.n_lap
        self.n_cells = low.n_cells
        self.y_dim = low.y_dim
        grid = [-(-n // b) for n, b in zip(reversed(shape3[: max(low.n_dims, 1)]), block)]
        self.n_blocks = int(np.prod(grid))
        handle = ctypes.c_void_p()
        _native.check(
            _native.lib().pml_plan_create(
                self.source.encode(),
                ctypes.byref(desc),
                cubin_path(self.source).encode(),
                ctypes.byref(handle),
            )
        )
        self.handle = handle
        self._keep = []  # tensors referenced by the current table set
        self._ws = None
        self._tables = _native.Tables()

    # -- bookkeeping --------------------------------------------------------
    @property
    def launches(self) -> int:
        return int(_native.lib().pml_plan_launches(self.handle))

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                _native.lib().pml_plan_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    def _dev(self, array: np.ndarray) -> torch.Tensor:
        return _from_numpy(array).to(self.device)

    def workspace(self) -> _native.Workspace:
        if self._ws is None:
            n = self.n_cells * self.y_dim
            f64 = dict(dtype=torch.float64, device=self.device)
            bufs = {
                "u_a": torch.empty(n, **f64),
                "u_b": torch.empty(n, **f64),
                "acc": torch.empty(n, **f64),
                "partials": torch.empty(max(self.n_blocks, 2), **f64),
                "flags": torch.zeros(4, dtype=torch.int32, device=self.device),
            }
            nl = max(self.n_lap, 0) * self.n_cells
            for k in ("lap_rhs", "jac_a", "jac_b"):
                bufs[k] = torch.empty(max(nl, 1), **f64)
            t_capacity = 4096 if self.spec.small_threads else 0
            bufs["t_dev"] = torch.empty(max(t_capacity, 1), **f64)
            ws = _native.Workspace()
            for k, t in bufs.items():
                setattr(ws, k, t.data_ptr())
            ws.t_capacity = t_capacity
            self._ws_bufs = bufs
            self._ws = ws
        return self._ws

    # -- tables -------------------------------------------------------------
    def bind_tables(
        self,
        low: LoweredProblem,
        dyn_neu: Optional[Dict[int, np.ndarray]] = None,
        dyn_dir: Optional[Dict[int, np.ndarray]] = None,
    ):
        """Points the plan at the boundary tables and coordinate vectors of
        ``low``.  Static tables are uploaded once per lowered problem; the
        dynamic ones carry one row per time slot and are uploaded per call."""
        static = getattr(low, "_device_static", None)
        if static is None or static[0] != self.device:
            static = (
                self.device,
                {f: self._dev(a) for f, a in low.static_neu.items()},
                {f: self._dev(a) for f, a in low.static_dir.items()},
                [self._dev(a) for a in low.coords],
                [None if a is None else self._dev(a) for a in low.aux],
            )
            low._device_static = static
        _, s_neu, s_dir, coords, aux = static
        keep = [static]
        tabs = _native.Tables()
        for f in range(6):
            for stat, dyn, ptrs, strides in (
                (s_neu, dyn_neu, tabs.neu, tabs.neu_stride),
                (s_dir, dyn_dir, tabs.dir, tabs.dir_stride),
            ):
                if f in stat:
                    ptrs[f] = stat[f].data_ptr()
                    strides[f] = 0
                elif dyn is not None and f in dyn:
                    t = self._dev(dyn[f])
                    keep.append(t)
                    ptrs[f] = t.data_ptr()
                    strides[f] = t.shape[1]
                else:
                    ptrs[f] = None
                    strides[f] = 0
        for i, c in enumerate(coords):
            tabs.coord[i] = c.data_ptr()
        for i, a in enumerate(aux):
            tabs.aux[i] = None if a is None else a.data_ptr()
        self._keep = keep
        self._tables = tabs
        _native.check(
            _native.lib().pml_plan_set_tables(self.handle, ctypes.byref(tabs))
        )

    # -- execution ------------------------------------------------------------
    def run(
        self,
        integrator: str,
        y0: torch.Tensor,
        traj: torch.Tensor,
        t_starts: np.ndarray,
        d_t: float,
        slot0: int = 0,
        jacobi_init: Optional[torch.Tensor] = None,
        jacobi_tol: float = 1e-3,
        max_sweeps: int = 0,
    ) -> Optional[np.ndarray]:
        """``len(t_starts)`` steps from the planes ``y0`` into ``traj``
        (steps, y_dim * n_cells).  Asynchronous unless the system has
        Y_LAPLACIAN equations; returns the Jacobi sweep counts in that case."""
        n_steps = len(t_starts)
        assert traj.shape[0] >= n_steps and traj.is_contiguous()
        t_host = np.ascontiguousarray(t_starts, dtype=np.float64)
        sweeps = (ctypes.c_int * max(n_steps, 1))()
        _native.check(
            _native.lib().pml_fdm_run(
                self.handle,
                _native.INTEGRATOR_CODES[integrator],
                ctypes.byref(self.workspace()),
                y0.data_ptr(),
                traj.data_ptr(),
                traj.stride(0),
                t_host.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
                n_steps,
                float(d_t),
                slot0,
                ptr(jacobi_init),
                float(jacobi_tol),
                int(max_sweeps),
                sweeps,
                stream_ptr(),
            )
        )
        return np.array(sweeps[:n_steps]) if self.n_lap else None

    def run_batch(
        self,
        integrator: str,
        y0: torch.Tensor,
        traj: torch.Tensor,
        t_starts: np.ndarray,
        d_t: float,
        slot0: int = 0,
    ):
        """``len(t_starts)`` steps of every member of a batch: ``y0`` is
        (batch, y_dim * n_cells) planes, ``traj`` (batch, steps, y_dim *
        n_cells).  Small meshes: one launch, one thread block per member."""
        batch, n_steps = y0.shape[0], len(t_starts)
        assert traj.shape[0] == batch and traj.shape[1] >= n_steps
        assert y0.is_contiguous() and traj.is_contiguous()
        ws = self.workspace()
        keep = None
        ws_stride = 0
        if self.spec.small_threads and batch > 1:
            # every member needs its own stage buffers
            n = self.n_cells * self.y_dim
            keep = torch.empty((3, batch, n), dtype=torch.float64, device=self.device)
            batched = _native.Workspace()
            ctypes.memmove(
                ctypes.byref(batched), ctypes.byref(ws), ctypes.sizeof(ws)
            )
            batched.u_a, batched.u_b, batched.acc = (
                keep[0].data_ptr(), keep[1].data_ptr(), keep[2].data_ptr()
            )
            ws, ws_stride = batched, n
        t_host = np.ascontiguousarray(t_starts, dtype=np.float64)
        _native.check(
            _native.lib().pml_fdm_run_batch(
                self.handle, _native.INTEGRATOR_CODES[integrator],
                ctypes.byref(ws), y0.data_ptr(), traj.data_ptr(),
                traj.stride(1), batch, y0.stride(0), traj.stride(0), ws_stride,
                t_host.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
                n_steps, float(d_t), slot0, stream_ptr(),
            )
        )
        if keep is not None:
            keep.record_stream(torch.cuda.current_stream())

    def apply_static_dirichlet(self, planes: torch.Tensor):
        """Writes the static Dirichlet values of the problem into component
        planes (what ``DiscreteInitialCondition`` does on the host,
        initial_condition.py:86-89)."""
        if self.low.dir_mask == 0:
            return
        if not self.low.all_static:
            raise ValueError("static Dirichlet values of a dynamic problem")
        self.bind_tables(self.low)
        _native.check(
            _native.lib().pml_apply_dirichlet(
                self.handle, planes.data_ptr(), 0, stream_ptr()
            )
        )

    def eval_rhs(self, u: torch.Tensor, out: torch.Tensor, t: float = 0.0):
        _native.check(
            _native.lib().pml_eval_rhs(
                self.handle, u.data_ptr(), out.data_ptr(), float(t), 0,
                stream_ptr(),
            )
        )

    def jacobi(
        self, rhs: torch.Tensor, y_init: torch.Tensor, out: torch.Tensor,
        tol: float, max_sweeps: int = 0,
    ) -> int:
        sweeps = ctypes.c_int(0)
        _native.check(
            _native.lib().pml_jacobi_run(
                self.handle, ctypes.byref(self.workspace()), rhs.data_ptr(),
                y_init.data_ptr(), out.data_ptr(), 0, float(tol),
                int(max_sweeps), ctypes.byref(sweeps), stream_ptr(),
            )
        )
        return int(sweeps.value)


def get_plan(low: LoweredProblem, **spec_overrides) -> DevicePlan:
    """Plans are cached per generated source and device."""
    spec = low.spec(**spec_overrides)
    source = codegen.generate_source(spec)
    key = f"{torch.cuda.current_device() if torch.cuda.is_available() else -1}:" + (
        codegen.source_key(source)
    )
    plan = _PLANS.get(key)
    if plan is None:
        plan = DevicePlan(low, spec)
        _PLANS[key] = plan
    plan.low = low
    return plan


# -- layout conversion and host transfers -------------------------------------
def aos_to_soa(aos: torch.Tensor, n_cells: int, y_dim: int, n_states: int = 1):
    out = torch.empty_like(aos)
    if y_dim == 1 or n_cells == 1:
        out.copy_(aos)
        return out
    _native.check(
        _native.lib().pml_aos_to_soa(
            aos.data_ptr(), out.data_ptr(), n_cells, y_dim, n_states,
            stream_ptr(),
        )
    )
    global _TOTAL_LAUNCHES
    _TOTAL_LAUNCHES += 1
    return out


def soa_to_aos(soa: torch.Tensor, n_cells: int, y_dim: int, n_states: int = 1,
               out: Optional[torch.Tensor] = None):
    if out is None:
        out = torch.empty_like(soa)
    if y_dim == 1 or n_cells == 1:
        out.copy_(soa)
        return out
    _native.check(
        _native.lib().pml_soa_to_aos(
            soa.data_ptr(), out.data_ptr(), n_cells, y_dim, n_states,
            stream_ptr(),
        )
    )
    global _TOTAL_LAUNCHES
    _TOTAL_LAUNCHES += 1
    return out


#: pinned staging buffers for large host -> device uploads (two, reused)
_UPLOAD_STAGE = []
_UPLOAD_CHUNK = 1 << 25  # doubles (256 MiB)


def _to_device_staged(flat: torch.Tensor, dev: torch.device) -> torch.Tensor:
    """Pageable host tensor -> device through two pinned staging buffers: the
    (multi-threaded) host copy of one chunk overlaps the DMA of the previous
    one.  A plain ``.to(device)`` of pageable memory runs at ~11 GB/s on the
    B200 boxes, this at the speed of the host copy (PCIe gen5 x16 does 55)."""
    n = flat.numel()
    if n <= _UPLOAD_CHUNK:
        return flat.to(dev, non_blocking=True)
    while len(_UPLOAD_STAGE) < 2:
        _UPLOAD_STAGE.append(
            (torch.empty(_UPLOAD_CHUNK, dtype=torch.float64, pin_memory=True),
             torch.cuda.Event())
        )
    out = torch.empty(n, dtype=torch.float64, device=dev)
    stream = torch.cuda.current_stream()
    for k, first in enumerate(range(0, n, _UPLOAD_CHUNK)):
        stage, done = _UPLOAD_STAGE[k % 2]
        if k >= 2:
            done.synchronize()  # the DMA out of this buffer has finished
        m = min(_UPLOAD_CHUNK, n - first)
        stage[:m].copy_(flat[first : first + m])
        out[first : first + m].copy_(stage[:m], non_blocking=True)
        done.record(stream)
    for _, done in _UPLOAD_STAGE:
        done.synchronize()  # the buffers may be reused by the next upload
    return out


def upload_state(y: np.ndarray, n_cells: int, y_dim: int) -> torch.Tensor:
    """Channels-last host state -> component planes on the device."""
    dev = require_cuda()
    flat = _from_numpy(np.ascontiguousarray(y, dtype=np.float64).reshape(-1))
    return aos_to_soa(_to_device_staged(flat, dev), n_cells, y_dim)
