"""Launches world_size local processes joined in a torch.distributed group."""
import os
import socket
import sys
import traceback

import torch.multiprocessing as mp


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _entry(rank, world_size, port, backend, fn, args, errors):
    import torch.distributed as dist

    here = os.path.dirname(os.path.abspath(__file__))
    for p in (here, os.path.dirname(here)):
        if p not in sys.path:
            sys.path.insert(0, p)
    try:
        if backend == "nccl":
            import torch

            torch.cuda.set_device(rank % torch.cuda.device_count())
        dist.init_process_group(
            backend, init_method=f"tcp://127.0.0.1:{port}", rank=rank,
            world_size=world_size,
        )
        fn(rank, world_size, *args)
        dist.barrier()
    except Exception:
        errors.put((rank, traceback.format_exc()))
        raise
    finally:
        if dist.is_initialized():
            dist.destroy_process_group()


def run_distributed(fn, world_size, args=(), backend="gloo", attempts=2):
    """Runs ``fn(rank, world_size, *args)`` on ``world_size`` processes.  A
    rendezvous failure (port stolen between probing and binding) is retried
    once on a fresh port; assertion failures of ``fn`` are not."""
    ctx = mp.get_context("spawn")
    for attempt in range(attempts):
        errors = ctx.SimpleQueue()
        port = free_port()
        try:
            mp.spawn(
                _entry, args=(world_size, port, backend, fn, args, errors),
                nprocs=world_size, join=True,
            )
            return
        except Exception as e:
            msgs = []
            while not errors.empty():
                msgs.append("rank %d:\n%s" % errors.get())
            text = "\n".join(msgs) or str(e)
            rendezvous = "AssertionError" not in text and (
                "address already in use" in text.lower()
                or "connect" in text.lower()
                or not msgs
            )
            if rendezvous and attempt + 1 < attempts:
                continue
            raise AssertionError(text)
