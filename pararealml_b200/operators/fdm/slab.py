"""Spatial slab decomposition of one FDM solve across GPUs (extension; the
reference has no spatial decomposition).

The mesh is cut along axis 0 into one slab of planes per rank.  A rank
stores its slab plus ``HALO`` planes of each neighbour and runs the ordinary
generated kernels on that *local sub-mesh*: they treat its outermost planes as
mesh faces, which spoils one plane per fused stage from each end -- exactly
the halo planes, which are overwritten by the neighbours' values after every
launch (``pml_fdm_phase``).  The planes a rank owns therefore hold the same
values as in the undecomposed solve.  True mesh faces (first / last rank) keep
their boundary tables; tables of the other axes and the axis-0 coordinate
vectors are sliced to the slab.

One rank per GPU with ``torch.distributed``: NCCL moves the halo planes GPU to
GPU over NVLink; a gloo group (tests, several ranks on one GPU) stages them
through the host.  With the stage-pair kernels a launch is split by planes
(``pml_fdm_phase_planes``): the ``EDGE_PLANES`` next to each neighbour are
computed first and their exchange -- pack, send / receive on a dedicated
high-priority NCCL communicator, unpack, all on a high-priority side stream --
runs while the launch of the remaining planes computes.
"""
import ctypes
from dataclasses import replace
from typing import Tuple

import numpy as np
import torch
import torch.distributed as dist

from pararealml_b200 import _native
from pararealml_b200.operators.fdm import codegen
from pararealml_b200.operators.fdm import device as dv
from pararealml_b200.operators.fdm.lowering import LoweredProblem

#: halo planes per side: a fused stage pair spoils two planes from a local face
HALO = 2
#: planes next to a neighbouring slab that are computed ahead of the rest
EDGE_PLANES = 8


def slab_bounds(n_planes: int, size: int, rank: int) -> Tuple[int, int]:
    """Planes [z0, z1) of axis 0 owned by ``rank`` (balanced split)."""
    base, extra = divmod(n_planes, size)
    z0 = rank * base + min(rank, extra)
    return z0, z0 + base + (1 if rank < extra else 0)


def slab_lowered(low: LoweredProblem, lo: int, hi: int) -> LoweredProblem:
    """The lowered problem of the sub-mesh of planes [lo, hi) of axis 0."""
    n0 = low.shape[0]
    assert 0 <= lo < hi <= n0
    shape = (hi - lo,) + tuple(low.shape[1:])
    sub = replace(
        low,
        shape=shape,
        face_static=dict(low.face_static),
        face_cells={},
        static_neu={},
        static_dir={},
        coords=[c[lo:hi] if a == 0 else c for a, c in enumerate(low.coords)],
        aux=[
            (x[lo:hi] if (i == 0 and x is not None) else x)
            for i, x in enumerate(low.aux)
        ],
    )
    # a face of axis 0 exists only where the slab touches the mesh face
    drop = (0 if lo == 0 else 1) | (0 if hi == n0 else 2)
    sub.neu_mask = low.neu_mask & ~drop
    sub.neu_zero_mask = low.neu_zero_mask & ~drop
    sub.dir_mask = low.dir_mask & ~drop
    cells = int(np.prod(shape))
    for f in range(2 * len(shape)):
        axis = f // 2
        sub.face_cells[f] = cells // shape[axis]
        for src, dst, mask in (
            (low.static_neu, sub.static_neu, sub.neu_mask),
            (low.static_dir, sub.static_dir, sub.dir_mask),
        ):
            if f not in src or not (mask >> f) & 1:
                continue
            if axis == 0:
                dst[f] = src[f]
            else:
                # face cells are ordered (axis 0 index, other index): rows
                # [lo, hi) of the table
                row = low.face_cells[f] // n0 * low.y_dim
                dst[f] = np.ascontiguousarray(src[f][lo * row : hi * row])
    if hasattr(sub, "_device_static"):
        del sub._device_static
    return sub


_HALO_GROUPS = {}


def _halo_group():
    """A second NCCL communicator over all ranks whose kernels run on a
    high-priority stream (made once, kept for reuse)."""
    key = None
    if key not in _HALO_GROUPS:
        ranks = list(range(dist.get_world_size()))
        try:
            opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
            _HALO_GROUPS[key] = dist.new_group(ranks, backend="nccl", pg_options=opts)
        except (AttributeError, TypeError):  # option not available: plain group
            _HALO_GROUPS[key] = dist.new_group(ranks, backend="nccl")
    return _HALO_GROUPS[key]


class SlabSolver:
    """Per-rank state of a slab-decomposed solve of a (static boundary
    condition, fully time-stepped) problem."""

    def __init__(self, low: LoweredProblem, family: str, group=None):
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("slab decomposition needs torch.distributed")
        if low.n_dims < 2:
            raise ValueError("slab decomposition needs a mesh with >= 2 axes")
        if not low.all_static:
            raise NotImplementedError(
                "slab decomposition with dynamic boundary conditions"
            )
        if len(low.kind_indices("D_Y_OVER_D_T")) != low.y_dim:
            raise NotImplementedError(
                "slab decomposition of systems with algebraic or Poisson equations"
            )
        self.group = group
        self.size = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.on_nccl = dist.get_backend(group) == "nccl"
        # NCCL, all ranks of the job: the halo transfers get their own
        # communicator on a high-priority stream (new_group is collective over
        # the default group, so a caller-made subgroup is used as it is)
        self.halo_group = (
            _halo_group() if (self.on_nccl and group is None) else group
        )
        self.low = low
        self.family = family
        n0 = low.shape[0]
        self.z0, self.z1 = slab_bounds(n0, self.size, self.rank)
        if self.z1 - self.z0 < HALO:
            raise ValueError(
                f"{n0} planes over {self.size} ranks leave fewer than {HALO} "
                "planes per rank"
            )
        self.lo = max(self.z0 - HALO, 0)
        self.hi = min(self.z1 + HALO, n0)
        self.local = slab_lowered(low, self.lo, self.hi)
        self.plane = int(np.prod(low.shape[1:]))
        self.n_loc = self.hi - self.lo
        self.state = low.y_dim * self.n_loc * self.plane
        overrides = {
            "passthrough": False,
            "fused": codegen.default_fused(self.local.shape, low.y_dim, low.y_dim, False),
            "small_threads": 0,  # phases are separate launches
            "zrep": codegen.default_zrep(self.local.shape),
        }
        self.plan = dv.get_plan(self.local, **overrides)
        self.plan.bind_tables(self.local)
        code = _native.INTEGRATOR_CODES[family]
        self.code = code
        self.n_phases = int(_native.lib().pml_fdm_phase_count(self.plan.handle, code))
        if self.n_phases <= 0:
            _native.check(-1)
        ws = self.plan.workspace()
        self._ws_by_ptr = {
            t.data_ptr(): t for t in self.plan._ws_bufs.values() if t.numel() >= self.state
        }
        self._ws = ws
        f64 = dict(dtype=torch.float64, device=self.plan.device)
        n_send = low.y_dim * HALO * self.plane
        self._pack = [torch.empty(n_send, **f64) for _ in range(2)]
        self._unpack = [torch.empty(n_send, **f64) for _ in range(2)]
        self._ops = None  # NCCL: the four transfers, built once
        self._comm = None  # side stream of the overlapped halo exchange

    # -- layout ---------------------------------------------------------------
    def local_planes(self, y: np.ndarray) -> torch.Tensor:
        """Channels-last host state of the whole mesh -> component planes of
        this rank's slab (with halo) on the device."""
        part = np.ascontiguousarray(y[self.lo : self.hi])
        return dv.upload_state(part, self.n_loc * self.plane, self.low.y_dim)

    def owned(self, planes: torch.Tensor) -> torch.Tensor:
        """(..., y_dim * n_loc * plane) -> view (..., y_dim, owned planes, plane)"""
        lead = planes.shape[:-1]
        v = planes.view(*lead, self.low.y_dim, self.n_loc, self.plane)
        return v[..., self.z0 - self.lo : self.z1 - self.lo, :]

    # -- halo exchange ----------------------------------------------------------
    def _neighbours(self, v: torch.Tensor):
        """(side, peer, planes to send, halo planes to fill) of ``v``
        (y_dim, n_loc, plane)."""
        a = self.z0 - self.lo  # first owned local plane
        b = self.z1 - self.lo  # one past the last owned local plane
        out = []
        if self.rank > 0:
            out.append((0, self.rank - 1, v[:, a : a + HALO], v[:, a - HALO : a]))
        if self.rank + 1 < self.size:
            out.append((1, self.rank + 1, v[:, b - HALO : b], v[:, b : b + HALO]))
        return out

    def _global(self, peer: int) -> int:
        return peer if self.group is None else dist.get_global_rank(self.group, peer)

    def exchange(self, buf: torch.Tensor):
        """Overwrites the halo planes of ``buf`` (y_dim * n_loc * plane
        doubles) with the neighbours' outermost owned planes."""
        v = buf.view(self.low.y_dim, self.n_loc, self.plane)
        shape = (self.low.y_dim, HALO, self.plane)
        sides = self._neighbours(v)
        if not sides:
            return
        for side, _, send_view, _ in sides:
            self._pack[side].view(shape).copy_(send_view)
        if self.on_nccl:
            # GPU to GPU over NVLink, one grouped launch for all four transfers
            if self._ops is None:
                self._ops = []
                for side, peer, _, _ in sides:
                    g = self._global(peer)
                    self._ops.append(dist.P2POp(dist.isend, self._pack[side], g, self.halo_group))
                    self._ops.append(dist.P2POp(dist.irecv, self._unpack[side], g, self.halo_group))
            for work in dist.batch_isend_irecv(self._ops):
                work.wait()
            for side, _, _, halo_view in sides:
                halo_view.copy_(self._unpack[side].view(shape))
            return
        # gloo (tests, ranks sharing a GPU): staged through the host
        ops, staged = [], []
        for side, peer, _, halo_view in sides:
            g = self._global(peer)
            recv = torch.empty(self._unpack[side].shape, dtype=torch.float64)
            ops.append(dist.P2POp(dist.isend, self._pack[side].cpu(), g, self.group))
            ops.append(dist.P2POp(dist.irecv, recv, g, self.group))
            staged.append((halo_view, recv))
        for work in dist.batch_isend_irecv(ops):
            work.wait()
        for halo_view, recv in staged:
            halo_view.copy_(recv.view(shape).to(buf.device))

    # -- time stepping ----------------------------------------------------------
    def _launch(self, y, y_next, t, d_t, phase, z_begin, z_end, fresh):
        _native.check(
            _native.lib().pml_fdm_phase_planes(
                self.plan.handle, self.code, ctypes.byref(self._ws),
                y.data_ptr(), y_next.data_ptr(), float(t), float(d_t),
                0, phase, z_begin, z_end, ctypes.byref(fresh), dv.stream_ptr(),
            )
        )

    def edge_ranges(self):
        """Plane ranges of one phase: the planes next to a neighbouring slab
        (launched first, their outermost owned planes are what the neighbour
        waits for) and the remaining planes, whose launch the halo exchange
        overlaps.  Without the stage-pair kernels, or on slabs too thin to
        split, one launch covers everything."""
        n = self.n_loc
        edge = min(EDGE_PLANES, n // 4)
        # (forward Euler is a single stage: it has no stage-pair kernel)
        if (not self.plan.spec.fused or self.family == "forward_euler"
                or self.size == 1 or edge < 2 * HALO):
            return [], (0, n)
        edges = []
        lo, hi = 0, n
        if self.rank > 0:
            edges.append((0, edge))
            lo = edge
        if self.rank + 1 < self.size:
            edges.append((n - edge, n))
            hi = n - edge
        return edges, (lo, hi)

    def integrate(self, y0_planes: torch.Tensor, t: np.ndarray, d_t: float,
                  traj: torch.Tensor):
        """Steps starting at ``t[:-1]`` from the local planes ``y0_planes``
        (halo planes valid) into ``traj[j]`` (local planes, halo planes valid
        on return).  Per launch ("phase") the planes next to the neighbours
        are computed first; their exchange (side stream) runs while the rest
        of the slab is computed."""
        n_steps = len(t) - 1
        assert traj.shape == (n_steps, self.state) and traj.is_contiguous()
        fresh = ctypes.c_void_p()
        edges, rest = self.edge_ranges()
        overlap = bool(edges) and self.on_nccl
        compute = torch.cuda.current_stream()
        if overlap and self._comm is None:
            # high priority: the pack / unpack copies (and, through the halo
            # group below, NCCL's transfer kernels) take the first SMs the
            # interior launch frees instead of queueing behind all its blocks
            self._comm = torch.cuda.Stream(priority=-1)
        y = y0_planes
        for j in range(n_steps):
            y_next = traj[j]
            for phase in range(self.n_phases):
                for z_begin, z_end in edges:
                    self._launch(y, y_next, t[j], d_t, phase, z_begin, z_end, fresh)
                if overlap:
                    ready = torch.cuda.Event()
                    ready.record(compute)
                self._launch(y, y_next, t[j], d_t, phase, rest[0], rest[1], fresh)
                buf = y_next if fresh.value == y_next.data_ptr() else self._ws_by_ptr[fresh.value]
                if overlap:
                    # the edge planes are final: exchange them on the side
                    # stream while the launch above computes the interior
                    with torch.cuda.stream(self._comm):
                        self._comm.wait_event(ready)
                        self.exchange(buf[: self.state])
                        done = torch.cuda.Event()
                        done.record(self._comm)
                    compute.wait_event(done)
                else:
                    self.exchange(buf[: self.state])
            y = y_next

    # -- gathering ----------------------------------------------------------------
    def gather(self, traj: torch.Tensor) -> np.ndarray:
        """The trajectory of the whole mesh, channels-last, on the host of
        every rank (all-gather of the owned planes)."""
        n_steps = traj.shape[0]
        c, n0 = self.low.y_dim, self.low.shape[0]
        mine = self.owned(traj).contiguous()  # (steps, C, owned, plane)
        counts = [
            slab_bounds(n0, self.size, r)[1] - slab_bounds(n0, self.size, r)[0]
            for r in range(self.size)
        ]
        widest = max(counts)
        padded = torch.zeros((n_steps, c, widest, self.plane), dtype=torch.float64,
                             device=mine.device if self.on_nccl else "cpu")
        padded[:, :, : mine.shape[2]] = mine if self.on_nccl else mine.cpu()
        out = torch.empty((self.size,) + tuple(padded.shape), dtype=torch.float64,
                          device=padded.device)
        dist.all_gather_into_tensor(out.view(-1), padded.view(-1), group=self.group)
        out = out.cpu().numpy()
        full = np.empty((n_steps, n0, self.plane, c))
        for r in range(self.size):
            z0, z1 = slab_bounds(n0, self.size, r)
            # (steps, C, planes, plane) -> (steps, planes, plane, C)
            full[:, z0:z1] = np.moveaxis(out[r][:, :, : z1 - z0], 1, -1)
        return full.reshape((n_steps,) + tuple(self.low.shape) + (c,))
