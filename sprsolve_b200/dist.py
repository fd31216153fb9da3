"""One-process-per-GPU plumbing around the C ABI (host logic only; the data path -- halo exchange
and scalar all-reduces -- lives in csrc/dist.cu and runs over NVLink).

The reference has no multi-process path (single process, rayon / MKL threads: src/mat.rs:85-107);
this is the launch-side glue a `torchrun` job needs: rendezvous, distribution of the 128-byte
communicator id, the row-block partition, max-over-ranks timing.  `torch.distributed` is plumbing
only: "nccl" on GPU boxes, "gloo" in the CPU tests (tests/test_dist_gloo.py, world_size 2).
"""
from __future__ import annotations

import ctypes as C
import os

from . import _ffi as F


def env_world():
    """(world, rank, local_rank) from the torchrun environment (1, 0, 0 when not launched by it)."""
    return (int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")))


def init_process_group(backend: str | None = None, device=None):
    """Join the job's process group (no-op for world 1).  Rendezvous on 127.0.0.1 by default:
    the container hostname may not resolve."""
    import torch
    import torch.distributed as dist

    world, rank, _ = env_world()
    if world == 1 or dist.is_initialized():
        return world, rank
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29533")
    backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
    kw = {"device_id": device} if (backend == "nccl" and device is not None) else {}
    dist.init_process_group(backend, rank=rank, world_size=world, **kw)
    return world, rank


def broadcast_bytes(payload: bytes | None, src: int = 0) -> bytes:
    """`payload` of rank `src` on every rank."""
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size() == 1:
        return bytes(payload)
    box = [bytes(payload) if dist.get_rank() == src else None]
    dist.broadcast_object_list(box, src=src)
    return box[0]


def all_gather_bytes(payload: bytes) -> list[bytes]:
    """Every rank's `payload`, ordered by rank."""
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size() == 1:
        return [bytes(payload)]
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, bytes(payload))
    return out


def attach_communicator(ctx, make_id=None):
    """Give `ctx` its rank in the job: rank 0 creates the communicator id, everyone receives it,
    every rank calls spb_comm_init (collective).  `make_id` is injectable for CPU tests."""
    world, rank, _ = env_world()
    if world == 1:
        return world, rank
    make_id = make_id or type(ctx).comm_unique_id
    uid = broadcast_bytes(make_id() if rank == 0 else None, src=0)
    if len(uid) != 128:
        raise ValueError("communicator id must be 128 bytes")
    ctx.comm_init(world, rank, uid)
    return world, rank


def row_block(kind: int, nx: int, ny: int, nz: int, world: int, rank: int):
    """[begin, end) of the rows spb_csr_create_stencil gives `rank`: contiguous, cut at plane
    boundaries.  Host arithmetic inside the C ABI (spb_stencil_partition), no GPU needed."""
    b, e = C.c_int64(0), C.c_int64(0)
    st = F.lib().spb_stencil_partition(kind, nx, ny, nz, world, rank, C.byref(b), C.byref(e))
    if st != F.OK:
        raise ValueError(F.last_error())
    return int(b.value), int(e.value)


def max_over_ranks(v: float, device=None) -> float:
    """Timing rule: a multi-GPU number is the max over ranks."""
    import torch
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(v)
    t = torch.tensor([float(v)], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier(device=None):
    import torch
    import torch.distributed as dist

    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
    if device is not None and torch.cuda.is_available():
        torch.cuda.synchronize()
