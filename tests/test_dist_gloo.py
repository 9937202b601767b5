"""Multi-rank host logic on CPU: world_size-2 gloo processes on 127.0.0.1.  The batch partition and the
result gathers are what runs between GPUs (NCCL) on the box; the arithmetic under test here is only the
bookkeeping, since stepping itself needs a GPU."""
import os
import socket

import numpy as np
import pytest

from br2 import _dist


def test_partition_is_contiguous_and_complete():
    for total, world in ((4096, 8), (10, 3), (7, 7), (65536, 8), (5, 2)):
        shares = [_dist.partition(total, world, r) for r in range(world)]
        assert shares[0][0] == 0 and shares[-1][1] == total
        assert all(a[1] == b[0] for a, b in zip(shares, shares[1:]))
        sizes = [b - a for a, b in shares]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        _dist.partition(4, 2, 2)


def test_shard_action_slices_arrays_and_broadcasts_scalars():
    action = {"action1": np.arange(10.0), "action2": 3.5}
    a0 = _dist.shard_action(action, 10, 3, 0)
    a2 = _dist.shard_action(action, 10, 3, 2)
    np.testing.assert_array_equal(a0["action1"], [0, 1, 2, 3])
    np.testing.assert_array_equal(a2["action1"], [7, 8, 9])
    assert a0["action2"] == 3.5
    with pytest.raises(ValueError):
        _dist.shard_action({"a": np.arange(9.0)}, 10, 3, 0)


def _free_port():
    with socket.socket() as sock:
        sock.bind(("127.0.0.1", 0))
        return sock.getsockname()[1]


class _FakeProgram:
    layout = [{"x": (0, 6), "len": (6, 1)}, {"x": (7, 6), "len": (13, 1)}]  # two rods of one element each
    n_elems = [1, 1]


class _FakeEngine:
    """What StateCapture reads of an Engine (the batch state itself is a CPU tensor here)."""
    program = _FakeProgram()
    device = "cpu"

    def __init__(self, batch):
        self.batch = batch


def _worker(rank, world, port, total, out_dir):
    import torch
    import torch.distributed as dist

    from br2._capture import CaptureSink, StateCapture, _Lazy

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        start, stop = _dist.partition(total, world, rank)
        n_local = stop - start
        # every instance carries its global index, so the gathered order is checkable
        local = np.arange(start, stop, dtype=np.float64)[:, None, None] * np.ones((1, 3, 4))
        everywhere = _dist.gather_batch(local, total)
        on_root = _dist.gather_batch(local, total, dst=0)
        # a rank that holds ONE instance records arrays without a batch axis, like the reference does
        rec = (lambda a: a[0]) if n_local == 1 else (lambda a: a)
        sink = {"time": [0.04, 0.08], "step": [2000, 4000],
                "position": [rec(local + 0.25), rec(local + 0.5)], "lengths": [rec(local[:, 0]), rec(2 * local[:, 0])]}
        history = _dist.gather_history(sink, total, dst=0, local_batch=n_local)
        flag = _dist.reduce_flags(rank == 1)
        # device-side captures (CPU tensors under gloo): one block gather per capture, entries stay lazy on the root
        engine = _FakeEngine(n_local)
        block = torch.arange(start, stop, dtype=torch.float64)[:, None] * 100.0 + torch.arange(14, dtype=torch.float64)[None, :]
        merged_block = _dist.gather_block(block, total, dst=0)
        captures = [StateCapture(engine, tensor=block.clone()), StateCapture(engine, tensor=block.clone() + 0.5)]
        sinks = [CaptureSink(), CaptureSink()]
        for r, lazy_sink in enumerate(sinks):
            for t, capture in enumerate(captures):
                lazy_sink["time"].append(0.04 * (t + 1))
                lazy_sink["step"].append(2000 * (t + 1))
                lazy_sink["position"].append(_Lazy(capture, r, "x", (3, 2)))
                lazy_sink["com"].append(_Lazy(capture, r, "x", (3, 2), post=lambda x: x.mean(axis=-1)))
        gathered = _dist.gather_captures(sinks, engine, total, dst=0)
        lazy_hist = [_dist.gather_history(sk, total, dst=0, local_batch=n_local, gathered_captures=gathered) for sk in sinks]
        extra = {}
        if rank == 0:
            extra = dict(block=merged_block.numpy(), lazy_pos=np.array(lazy_hist[1]["position"]),
                         lazy_com=np.array(lazy_hist[0]["com"]), lazy_repr=repr(lazy_hist[1]["position"]))
        else:
            assert merged_block is None and lazy_hist == [None, None]
        np.savez(os.path.join(out_dir, "rank%d.npz" % rank), everywhere=everywhere,
                 on_root=np.zeros(0) if on_root is None else on_root,
                 hist_pos1=np.zeros(0) if history is None else history["position"][1],
                 hist_len0=np.zeros(0) if history is None else history["lengths"][0],
                 hist_steps=np.zeros(0) if history is None else np.array(history["step"]),
                 flag=flag, root_is_none=on_root is None, hist_is_none=history is None, **extra)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("total", [8, 7, 2, 3])
def test_gather_over_gloo_world_size_2(tmp_path, total):
    """total = 2 and 3: a rank holds ONE instance, whose recorded arrays carry no batch axis (_dist.py)."""
    import torch.multiprocessing as mp

    port = _free_port()
    mp.spawn(_worker, args=(2, port, total, str(tmp_path)), nprocs=2, join=True)
    want = np.arange(total, dtype=np.float64)[:, None, None] * np.ones((1, 3, 4))
    r0, r1 = np.load(tmp_path / "rank0.npz"), np.load(tmp_path / "rank1.npz")
    for r in (r0, r1):  # all-gather: every rank holds the global batch in rank order
        np.testing.assert_array_equal(r["everywhere"], want)
        assert bool(r["flag"]) is True  # OR over ranks
    np.testing.assert_array_equal(r0["on_root"], want)
    assert bool(r1["root_is_none"]) and bool(r1["hist_is_none"])
    np.testing.assert_array_equal(r0["hist_pos1"], want + 0.5)
    np.testing.assert_array_equal(r0["hist_len0"], want[:, 0])
    np.testing.assert_array_equal(r0["hist_steps"], [2000, 4000])
    # block gather and the lazy histories that ride on it
    block = np.arange(total, dtype=np.float64)[:, None] * 100.0 + np.arange(14, dtype=np.float64)[None, :]
    np.testing.assert_array_equal(r0["block"], block)
    assert "still on the device" in str(r0["lazy_repr"])
    want_pos = np.stack([block[:, 7:13].reshape(total, 3, 2), block[:, 7:13].reshape(total, 3, 2) + 0.5])  # rod 1, two captures
    np.testing.assert_array_equal(r0["lazy_pos"], want_pos)
    want_com = np.stack([block[:, 0:6].reshape(total, 3, 2).mean(axis=-1), (block[:, 0:6] + 0.5).reshape(total, 3, 2).mean(axis=-1)])
    np.testing.assert_array_equal(r0["lazy_com"], want_com)
