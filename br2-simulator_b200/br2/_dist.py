"""Batch partition across the GPUs of one box and the gather of results.

Assembly instances never interact, so the batch is split contiguously, every
rank integrates its share with its own GPU and **no collective sits on the step path**.  Collectives
(NCCL over NVLink / NVSwitch on GPUs, gloo in the CPU tests) are used only to bring trajectories and
per-instance metrics together -- what the reference's single process would have held in its callback
dictionaries (/root/reference/br2/free_simulator.py:56-72) and ``TerminalInfo``
(/root/reference/br2/environment.py:72-131).
"""
import numpy as np


def partition(total, world_size, rank):
    """Contiguous share ``[start, stop)`` of ``total`` instances for ``rank``; the first
    ``total % world_size`` ranks hold one instance more."""
    if world_size < 1 or not 0 <= rank < world_size:
        raise ValueError("bad rank / world_size")
    base, extra = divmod(int(total), int(world_size))
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_action(action, total, world_size, rank):
    """This rank's slice of a global action dict (scalars are broadcast, arrays are sliced)."""
    start, stop = partition(total, world_size, rank)
    out = {}
    for name, value in action.items():
        arr = np.asarray(value)
        if arr.ndim == 0:
            out[name] = value
        else:
            if arr.shape[0] != total:
                raise ValueError("action %r has %d entries, batch is %d" % (name, arr.shape[0], total))
            out[name] = arr[start:stop]
    return out


def _dist():
    import torch.distributed as dist

    return dist if dist.is_available() and dist.is_initialized() else None


def gather_batch(local, total, dst=None, device=None):
    """Concatenate per-rank ``[local_batch, ...]`` arrays (numpy or torch) along the batch axis in rank
    order.  ``dst=None`` -> every rank gets the full array (all-gather); otherwise only ``dst`` does
    (others return ``None``).  Uneven shares are padded to the largest one for the collective."""
    import torch

    dist = _dist()
    as_numpy = isinstance(local, np.ndarray)
    tensor = torch.from_numpy(np.ascontiguousarray(local)) if as_numpy else local.contiguous()
    if dist is None or dist.get_world_size() == 1:
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    if dist.get_backend() == "nccl":
        tensor = tensor.to(device if device is not None else torch.device("cuda", torch.cuda.current_device()))
    shares = [partition(total, world, r) for r in range(world)]
    widest = max(stop - start for start, stop in shares)
    padded = tensor.new_zeros((widest,) + tuple(tensor.shape[1:]))
    padded[: tensor.shape[0]] = tensor
    if dst is None:
        parts = [torch.empty_like(padded) for _ in range(world)]
        dist.all_gather(parts, padded)
    else:
        parts = [torch.empty_like(padded) for _ in range(world)] if rank == dst else None
        dist.gather(padded, parts, dst=dst)
        if rank != dst:
            return None
    full = torch.cat([part[: stop - start] for part, (start, stop) in zip(parts, shares)], dim=0)
    return full.cpu().numpy() if as_numpy else full


def gather_history(sink, total, dst=0):
    """Gather one rod's callback dictionary (lists of per-capture arrays with a leading local-batch axis)
    into the global-batch layout on ``dst``.  ``time`` / ``step`` are identical on every rank."""
    out = {}
    for key, series in sink.items():
        if key in ("time", "step"):
            out[key] = list(series)
            continue
        stacked = np.stack(series, axis=1) if len(series) else np.zeros((0, 0))  # [local_batch, T, ...]
        merged = gather_batch(stacked, total, dst=dst)
        if merged is not None:
            out[key] = [merged[:, t] for t in range(merged.shape[1])]
    rank = _dist().get_rank() if _dist() is not None else 0
    return out if rank == dst else None


def reduce_flags(local_flags):
    """Logical OR of per-rank booleans (NaN flags and the like)."""
    import torch

    dist = _dist()
    if dist is None or dist.get_world_size() == 1:
        return bool(local_flags)
    device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([1 if local_flags else 0], dtype=torch.int32, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return bool(t.item())
