"""One process per GPU: the rank plumbing that replaces Julia `Distributed` workers.

The reference farms contiguous chunks of supersources to workers (`sschunks`, src/fdtd/fdtd.jl:246-267)
and stacks per-worker gradients through a host `SharedArray` (src/fdtd/gradient.jl:2-11,
src/fdtd/propagate.jl:110-117).  Here a worker is a rank started by `torchrun` (RANK / LOCAL_RANK /
WORLD_SIZE in the environment); `torch.distributed` (NCCL on GPUs, gloo in CPU tests) is only the
control plane -- it carries the 128-byte `ncclUniqueId` to every rank and gathers records -- while the
data-path collective (one sum all-reduce of the FWI gradient) runs inside the engine on its own NCCL
communicator over NVLink (`gpi_nccl_init`, `gpi_allreduce_gradients`).
"""
from __future__ import annotations

import os
from typing import List, Optional

import numpy as np


def env_ranks():
    """(rank, local_rank, world_size) from the torchrun environment; (0, 0, 1) when launched plainly."""
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def init_process_group(backend: Optional[str] = None):
    """Initialise torch.distributed from the environment (no-op for a single process).
    Returns the module or None."""
    rank, local_rank, world = env_ranks()
    if world == 1:
        return None
    import torch
    import torch.distributed as dist
    if dist.is_initialized():
        return dist
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    if backend == "nccl":
        torch.cuda.set_device(local_rank)
        dist.init_process_group(backend, device_id=torch.device("cuda", local_rank))
    else:
        dist.init_process_group(backend)
    return dist


def share_unique_id(make_uid, dist=None, src: int = 0) -> Optional[bytes]:
    """Rank `src` creates the ncclUniqueId (`make_uid()` -> 128 bytes); everyone receives it."""
    if dist is None:
        return None
    box = [make_uid() if dist.get_rank() == src else None]
    dist.broadcast_object_list(box, src=src)
    uid = box[0]
    assert isinstance(uid, (bytes, bytearray)) and len(uid) == 128
    return bytes(uid)


def attach_nccl(pa, dist=None):
    """Give the experiment's engine its own NCCL communicator (one per handle)."""
    if dist is None:
        return
    uid = share_unique_id(pa.engine.nccl_unique_id, dist)
    pa.init_nccl(uid, dist.get_world_size())


def gather_records(pa, dist=None, dst: int = 0):
    """`update_datamat!` across workers (receiver.jl:17-34): every rank sends the records of its local
    supersources; rank `dst` ends up with the complete `pa.c.data` like the reference's master."""
    if dist is None:
        return pa.c.data
    mine = {iss: {f: pa.c.data[0][iss].d[f] for f in pa.c.rfields} for iss in pa.local}
    out: List = [None] * dist.get_world_size() if dist.get_rank() == dst else None
    dist.gather_object(mine, out, dst=dst)
    if dist.get_rank() == dst:
        for part in out:
            for iss, recs in part.items():
                for f, a in recs.items():
                    pa.c.data[0][iss].d[f][...] = a
    return pa.c.data


def allreduce_host(arrays, dist=None):
    """Sum numpy arrays over ranks through torch.distributed (control-plane fallback used by the CPU
    tests; the product's gradient all-reduce is `gpi_allreduce_gradients`)."""
    if dist is None:
        return arrays
    import torch
    out = []
    for a in arrays:
        t = torch.from_numpy(np.ascontiguousarray(a))
        dist.all_reduce(t)
        out.append(t.numpy().reshape(a.shape))
    return out


def stack_illum(pa, dist=None):
    """`stack_illums!` across workers (fdtd.jl:556-565, propagate.jl:110-117: upstream adds every worker's shots into one SharedArray):
    after `update!`, `pa.c.illum_stack` holds this rank's supersources; one sum over the ranks completes it on every rank.  A
    preconditioner read once per inversion (Float64, medium grid): control plane, not the data path."""
    if dist is None or getattr(pa.c, "illum_stack", None) is None:
        return getattr(pa.c, "illum_stack", None)
    import torch
    t = torch.from_numpy(np.ascontiguousarray(pa.c.illum_stack))
    if dist.get_backend() == "nccl":
        t = t.cuda()
        dist.all_reduce(t)
        pa.c.illum_stack[...] = t.cpu().numpy()
    else:
        dist.all_reduce(t)
        pa.c.illum_stack[...] = t.numpy()
    return pa.c.illum_stack
