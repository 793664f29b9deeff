"""Multi-GPU plumbing: one process per GPU, ensembles sharded by contiguous member block.

Members are independent given the forcing (the reference's member loop has no cross-iteration
dependence, ``rrmpg/models/hbvedu.py:199-209``), so the path shards with no data-path collective:
rank r simulates columns ``[lo_r, hi_r)`` of the ``[T, N]`` result and the only communication is
ONE broadcast of the (small, member-independent) forcing arrays from rank 0 before the launch.
``torch.distributed`` (NCCL over NVLink on GPUs, gloo in the CPU tests) carries that broadcast.
"""
import os

import numpy as np


def env_world():
    """(rank, local_rank, world_size) from the torchrun environment (1 process = 1 GPU)."""
    return (int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)),
            int(os.environ.get("WORLD_SIZE", 1)))


def member_block(n_members, rank, world_size):
    """Contiguous member block ``[lo, hi)`` of ``rank``; blocks differ in size by at most one."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    base, rem = divmod(int(n_members), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def init_process_group(backend=None):
    """Initialise torch.distributed from the environment; returns (rank, local_rank, world_size)."""
    import torch
    import torch.distributed as dist
    rank, local_rank, world = env_world()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        kwargs = {}
        if backend == "nccl":
            kwargs["device_id"] = torch.device("cuda", local_rank)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kwargs)
    return rank, local_rank, world


def shutdown():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.destroy_process_group()


def pack_forcing(series):
    """dict of equal-length float64 [T] (or [T, L]) arrays -> one [rows, T] matrix + the layout."""
    names = sorted(series)
    rows, layout = [], []
    for n in names:
        a = np.asarray(series[n], dtype=np.float64)
        a2 = a.reshape(a.shape[0], -1).T  # [cols, T]
        layout.append((n, a.shape))
        rows.append(a2)
    return np.ascontiguousarray(np.concatenate(rows, axis=0)), layout


def unpack_forcing(mat, layout):
    out, r = {}, 0
    for name, shape in layout:
        cols = int(np.prod(shape[1:])) if len(shape) > 1 else 1
        block = mat[r:r + cols]
        out[name] = (block.T.reshape(shape) if len(shape) > 1 else block[0])
        r += cols
    return out


def broadcast_forcing(tensor, src=0):
    """One collective for the whole forcing block (in place).  No-op for a single process."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(tensor, src=src)
    return tensor


def max_over_ranks(value, device=None):
    """Max of a python float over all ranks (timings are reported as the slowest rank)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()


_HOST_GROUP = None


def host_group():
    """A gloo group over all ranks (collective call: every rank must make it at the same point)."""
    global _HOST_GROUP
    import torch.distributed as dist
    if _HOST_GROUP is None and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        _HOST_GROUP = dist.group.WORLD if dist.get_backend() == "gloo" else dist.new_group(backend="gloo")
    return _HOST_GROUP


def host_barrier():
    """Barrier on the HOST side (gloo): the waiting ranks launch nothing on their GPUs.  An NCCL barrier parks a
    spinning kernel on every waiting rank's device -- fatal for timing when another process (rank 0 driving all the
    GPUs of the node through the library's own member sharding) uses those devices in the meantime."""
    import torch.distributed as dist
    g = host_group()
    if g is not None:
        dist.barrier(group=g)


def broadcast_series(series, src=0, device=None):
    """Broadcast a dict of forcing arrays from ``src`` to every rank: ONE object broadcast of the layout
    (names and shapes, a few hundred bytes) and ONE tensor broadcast of the packed block.  Ranks other than
    ``src`` may pass ``None``.  Returns the dict (numpy arrays) on every rank."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
        return {k: np.asarray(v, dtype=np.float64) for k, v in series.items()}
    rank = dist.get_rank()
    meta = [None]
    mat = None
    if rank == src:
        mat, layout = pack_forcing(series)
        meta = [(layout, mat.shape)]
    dist.broadcast_object_list(meta, src=src)
    layout, shape = meta[0]
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.as_tensor(mat, device=device) if rank == src else torch.empty(shape, dtype=torch.float64, device=device)
    broadcast_forcing(t, src=src)
    return unpack_forcing(t.cpu().numpy(), layout)


def simulate_sharded(model, params, series=None, src=0, **kwargs):
    """``model.simulate`` for this rank's contiguous member block of a global ensemble.

    ``params``: the GLOBAL record array (every rank draws or holds the same one, e.g. from the same seed);
    ``series``: dict of the time-series arguments of ``model.simulate`` (``prec``, ``temp`` ...), needed on
    ``src`` only -- it is broadcast once; everything else (scalars, ``altitudes`` ...) goes through ``kwargs``
    and must be equal on all ranks.  Returns ``(lo, hi, result)`` where ``result`` is what
    ``model.simulate(params=params[lo:hi])`` returns: columns ``[lo, hi)`` of the global ``[T, N]`` arrays.
    No collective touches the result; use :func:`gather_members` for small per-member vectors."""
    rank, _, world = env_world()
    forcing = broadcast_series(series, src=src) if (series is not None or world > 1) else {}
    lo, hi = member_block(len(params), rank, world)
    out = model.simulate(params=params[lo:hi], **forcing, **kwargs)
    return lo, hi, out


def gather_members(local, n_members):
    """All-gather a per-member vector (e.g. the fused MSE of this rank's block) into the global ``[N]`` order."""
    import torch
    import torch.distributed as dist
    local = np.ascontiguousarray(local, dtype=np.float64)
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
        return local
    world = dist.get_world_size()
    sizes = [member_block(n_members, r, world) for r in range(world)]
    width = max(hi - lo for lo, hi in sizes)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    buf = torch.zeros(width, dtype=torch.float64, device=dev)
    buf[:local.size] = torch.as_tensor(local, device=dev)
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf)
    return np.concatenate([p[:hi - lo].cpu().numpy() for p, (lo, hi) in zip(parts, sizes)])
