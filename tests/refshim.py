"""Import shim that lets the UNMODIFIED reference (``/root/reference``) be
imported in the build container: matplotlib and mpi4py are absent and the
reference calls ``np.product`` (removed in NumPy 2).  Used only by the golden
fixture generator and by CPU cross-check tests; never on the GPU box and never
by the product path.
"""
import os
import sys
import threading
import types
from unittest.mock import MagicMock

import numpy as np

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _find_reference_root():
    """``PML_REFERENCE_ROOT``, else the offline install under
    ``baseline/_ref`` (made by ``__graft_entry__.build()``; it travels to the
    GPU box with the snapshot), else the read-only checkout of the build
    container."""
    candidates = [
        os.environ.get("PML_REFERENCE_ROOT"),
        os.path.join(_REPO, "baseline", "_ref"),
        "/root/reference",
    ]
    for c in candidates:
        if c and os.path.isdir(os.path.join(c, "pararealml")):
            return c
    return candidates[0] or candidates[-1]


REFERENCE_ROOT = _find_reference_root()


class FakeComm:
    """Thread-per-rank stand-in for ``mpi4py.MPI.COMM_WORLD`` (only the calls
    ``parareal_operator.py:108-193`` and ``utils/time.py`` make)."""

    _local = threading.local()

    def __init__(self):
        self._size = 1
        self._barrier = threading.Barrier(1)
        self._slots = [None]

    def configure(self, size):
        self._size = size
        self._barrier = threading.Barrier(size)
        self._slots = [None] * size

    def set_rank(self, rank):
        FakeComm._local.rank = rank

    @property
    def size(self):
        return self._size

    @property
    def rank(self):
        return getattr(FakeComm._local, "rank", 0)

    def Get_size(self):
        return self.size

    def Get_rank(self):
        return self.rank

    def barrier(self):
        self._barrier.wait()

    Barrier = barrier

    def Allgather(self, send, recv):
        send_buf = send[0] if isinstance(send, (list, tuple)) else send
        recv_buf = recv[0] if isinstance(recv, (list, tuple)) else recv
        self._slots[self.rank] = np.array(send_buf, copy=True)
        self._barrier.wait()
        flat = recv_buf.reshape(self._size, -1)
        for r in range(self._size):
            flat[r, :] = self._slots[r].reshape(-1)
        self._barrier.wait()


class GlooComm:
    """Process-per-rank stand-in for ``MPI.COMM_WORLD`` on a
    ``torch.distributed`` gloo group (timed host baselines: true parallelism,
    which the thread-per-rank communicator cannot give under the GIL)."""

    def __init__(self):
        import torch.distributed as dist

        self._dist = dist
        self.size = dist.get_world_size()
        self.rank = dist.get_rank()

    def Get_size(self):
        return self.size

    def Get_rank(self):
        return self.rank

    def barrier(self):
        self._dist.barrier()

    Barrier = barrier

    def Allgather(self, send, recv):
        import torch

        send_buf = send[0] if isinstance(send, (list, tuple)) else send
        recv_buf = recv[0] if isinstance(recv, (list, tuple)) else recv
        out = torch.from_numpy(recv_buf).view(-1)
        src = torch.from_numpy(np.ascontiguousarray(send_buf)).view(-1)
        self._dist.all_gather_into_tensor(out, src)


COMM = FakeComm()


def set_comm(comm):
    """Replaces ``mpi4py.MPI.COMM_WORLD`` of the (already installed) shim."""
    sys.modules["mpi4py.MPI"].COMM_WORLD = comm


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "pararealml"))


def install():
    """Makes ``import pararealml`` resolve to the reference. Idempotent."""
    if "pararealml" in sys.modules and getattr(
        sys.modules["pararealml"], "__file__", ""
    ).startswith(REFERENCE_ROOT):
        return sys.modules["pararealml"]
    if not available():
        raise ImportError(f"reference not found under {REFERENCE_ROOT}")

    for name in (
        "matplotlib",
        "matplotlib.pyplot",
        "matplotlib.cm",
        "matplotlib.animation",
        "matplotlib.collections",
        "matplotlib.colors",
        "matplotlib.contour",
        "matplotlib.figure",
        "matplotlib.lines",
        "matplotlib.quiver",
        "matplotlib.streamplot",
        "mpl_toolkits",
        "mpl_toolkits.mplot3d",
        "mpl_toolkits.mplot3d.art3d",
    ):
        sys.modules.setdefault(name, MagicMock())

    if not hasattr(np, "product"):
        np.product = np.prod

    if "mpi4py" not in sys.modules:
        mpi4py = types.ModuleType("mpi4py")
        mpi = types.ModuleType("mpi4py.MPI")
        mpi.COMM_WORLD = COMM
        mpi.DOUBLE = "DOUBLE"
        import time as _time

        mpi.Wtime = _time.perf_counter
        mpi4py.MPI = mpi
        sys.modules["mpi4py"] = mpi4py
        sys.modules["mpi4py.MPI"] = mpi

    # appended, not prepended: the reference checkout has its own ``tests``
    # package, which must not shadow this repo's in spawned worker processes
    if REFERENCE_ROOT not in sys.path:
        sys.path.append(REFERENCE_ROOT)
    import pararealml  # noqa: F401

    return pararealml


def run_ranks(size, fn):
    """Runs ``fn(rank)`` on ``size`` threads sharing the fake communicator and
    returns the per-rank results."""
    COMM.configure(size)
    results = [None] * size
    errors = []

    def target(rank):
        COMM.set_rank(rank)
        try:
            results[rank] = fn(rank)
        except BaseException as e:  # pragma: no cover
            errors.append(e)
            COMM._barrier.abort()

    threads = [threading.Thread(target=target, args=(r,)) for r in range(size)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    COMM.configure(1)
    if errors:
        raise errors[0]
    return results
