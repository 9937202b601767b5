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


def gather_block(block, total, dst=0):
    """``[local_batch, ...]`` torch tensor of every rank -> ``[total, ...]`` on ``dst`` (``None`` elsewhere), in rank
    order, on the device the blocks live on: point-to-point sends straight into the right rows of the output (NCCL over
    NVLink on GPUs, gloo on CPU tensors), so uneven shares need no padding and nothing is staged through the host."""
    dist = _dist()
    if dist is None or dist.get_world_size() == 1:
        return block
    world, rank = dist.get_world_size(), dist.get_rank()
    shares = [partition(total, world, r) for r in range(world)]
    block = block.contiguous()
    if block.shape[0] != shares[rank][1] - shares[rank][0]:
        raise ValueError("rank %d holds %d instances, its share of %d is %d" % (rank, block.shape[0], total, shares[rank][1] - shares[rank][0]))
    if rank != dst:
        for req in dist.batch_isend_irecv([dist.P2POp(dist.isend, block, dst)]):
            req.wait()
        return None
    out = block.new_empty((total,) + tuple(block.shape[1:]))
    ops = []
    for r, (start, stop) in enumerate(shares):
        if r == rank:
            out[start:stop].copy_(block)
        else:
            ops.append(dist.P2POp(dist.irecv, out[start:stop], r))
    for req in dist.batch_isend_irecv(ops):
        req.wait()
    return out


def gather_history(sink, total, dst=0, local_batch=None, gathered_captures=None):
    """Gather one rod's callback dictionary into the global-batch layout on ``dst`` (``None`` elsewhere).  ``time`` /
    ``step`` are identical on every rank.  Entries that are still references into a device-side capture
    (br2/_capture.py) are re-pointed at the gathered capture -- ``gathered_captures`` maps ``id(local capture)`` to the
    global one, filled by ``gather_captures`` with ONE device-to-device gather per capture for all rods and fields --
    and stay lazy; entries that already are host arrays (callback classes on the reference's host protocol) are gathered
    per field.  ``local_batch``: this rank's instance count (a rank that holds ONE instance stores its arrays without a
    batch axis, like the reference; the axis is put back before gathering)."""
    from ._capture import CaptureSeries, _Lazy

    rank = _dist().get_rank() if _dist() is not None else 0
    out = {}
    for key, series in sink.items():
        if key in ("time", "step"):
            out[key] = list(series)
            continue
        items = list(series._items) if isinstance(series, CaptureSeries) else list(series)
        if items and all(isinstance(v, _Lazy) for v in items) and gathered_captures is not None:
            if rank == dst:
                merged = CaptureSeries()
                for v in items:
                    merged.append(v.rebind(gathered_captures[id(v.capture)]))
                out[key] = merged
            continue
        arrays = [v.make() if isinstance(v, _Lazy) else np.asarray(v) for v in items]
        stacked = np.stack(arrays, axis=0) if arrays else np.zeros((0,))  # [T, (local_batch,) ...]
        if local_batch == 1 and arrays:
            stacked = stacked[:, None]
        if arrays:
            stacked = np.ascontiguousarray(np.moveaxis(stacked, 1, 0))      # [local_batch, T, ...]
        merged = gather_batch(stacked, total, dst=dst)
        if merged is not None:
            out[key] = [merged[:, t] for t in range(merged.shape[1])]
    return out if rank == dst else None


def gather_captures(sinks, engine, total, dst=0):
    """Every device-side capture referenced by the callback dictionaries ``sinks``, gathered once: returns
    ``{id(local capture): global capture}`` on ``dst`` (values ``None`` elsewhere).  A capture that is still a device
    clone travels device to device (``gather_block``); one that was already read moves from its host copy."""
    from ._capture import CaptureSeries, StateCapture, _Lazy

    import torch

    dist = _dist()
    rank = dist.get_rank() if dist is not None else 0
    seen, order = set(), []
    for sink in sinks:
        for key, series in sink.items():
            if isinstance(series, CaptureSeries):
                for v in series._items:
                    if isinstance(v, _Lazy) and id(v.capture) not in seen:
                        seen.add(id(v.capture))
                        order.append(v.capture)
    on_gpu = dist is not None and dist.get_backend() == "nccl"
    out = {}
    for capture in order:
        if capture.on_device:
            block = capture._tensor
        else:
            block = torch.from_numpy(capture.host())
            if on_gpu:
                block = block.to(engine.device)
        merged = gather_block(block, total, dst=dst)
        if rank != dst:
            out[id(capture)] = None
        elif merged.is_cuda:
            out[id(capture)] = StateCapture(engine, tensor=merged, batch=total)
        else:
            out[id(capture)] = StateCapture(engine, host=merged.numpy(), batch=total)
    return out


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
