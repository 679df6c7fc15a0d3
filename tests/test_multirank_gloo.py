"""N > 1 path on CPU: world_size 2 over gloo.

The multi-GPU design is "shard the instances, no data-path collective": rank r owns a contiguous block of
instance indices, builds its own context from the broadcast shared tables and ticks it independently. What the
host side must get right is therefore (1) the partition, (2) that instance i gets the same inputs whichever rank
owns it (RNG stream = global instance index), (3) the three uses of torch.distributed (broadcast of the shared
tables, max-over-ranks timing, gather of the heads). Here two gloo ranks each drive their block through the CPU
oracle (test infrastructure; the GPU ranks drive libsmpc_cuda.so with the very same call sequence, bench.py) and
the gathered result must equal a single process solving the whole batch, bit for bit.
"""
import os
import socket
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "simple-mpc_b200", "python"))
sys.path.insert(0, HERE)
import smpc_dist  # noqa: E402

TOTAL, TICKS = 5, 3   # 5 instances over 2 ranks: ragged blocks (3 + 2)


def test_shard_range_partitions_every_instance_once():
    for total in (0, 1, 5, 7, 4096, 4097):
        for world in (1, 2, 3, 4, 8):
            owned = []
            for r in range(world):
                first, count = smpc_dist.shard_range(total, world, r)
                owned += list(range(first, first + count))
                # the closed form of SURVEY §8e: instance i lives on rank floor(i * world / total)
                assert all(i * world // total == r for i in range(first, first + count))
            assert owned == list(range(total))
            sizes = [smpc_dist.shard_range(total, world, r)[1] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    assert smpc_dist.shard_range(4096, 8, 3) == (1536, 512)
    with pytest.raises(ValueError):
        smpc_dist.shard_range(8, 2, 2)


def _drive(first, count, ticks):
    """constructor solve + `ticks` MPC ticks of instances [first, first+count) through the oracle"""
    import ctypes as C
    import parity_utils as pu
    from smpc_ctypes import SolverOpts
    from oracle_libs import load_oracle
    orc = load_oracle()
    sc = pu.Go2Scenario(count, seed=5, first=first)
    opts = SolverOpts()
    orc.orc_default_solver_opts(C.byref(opts))
    b = pu.Backend(orc, "orc", sc.model, sc.ocpb, opts, count)
    try:
        b.set_slots(sc.initial_slots())
        b.set_terminal(sc.terminals())
        b.set_warmstart(*sc.warmstart())
        b.solve(np.tile(sc.xref, (count, 1)), 100)
        x = sc.x_start()
        for t in range(ticks):
            b.cycle(sc.tails(t))
            b.shift()
            b.solve(x, 1)
            x = sc.disturb(b.head()[0], t + 1)
        x1, u0, u1, K0 = b.head()
        return bytes(sc.model), x1, u0, K0
    finally:
        b.close()


def _worker(rank, world, port, out_path):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["OMP_NUM_THREADS"] = "2"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        first, count = smpc_dist.shard_range(TOTAL, world, rank)
        model_bytes, x1, u0, K0 = _drive(first, count, TICKS)
        ref, same = smpc_dist.broadcast_bytes(dist, model_bytes)
        assert same and ref == model_bytes
        # a rank holding a different table must be told so
        _, same_bad = smpc_dist.broadcast_bytes(dist, model_bytes if rank == 0 else model_bytes[:-1] + b"\x01")
        assert same_bad == (rank == 0)
        tmax = smpc_dist.max_over_ranks(dist, [float(rank + 1), 10.0 - rank])
        assert tmax == [float(world), 10.0]
        g = [smpc_dist.gather_blocks(dist, a, TOTAL) for a in (x1, u0, K0)]
        if rank == 0:
            np.savez(out_path, x1=g[0], u0=g[1], K0=g[2])
        else:
            assert all(a is None for a in g)
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_ranks_equal_one_process(tmp_path):
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "gathered.npz")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = np.load(out)
    _, x1, u0, K0 = _drive(0, TOTAL, TICKS)
    assert got["x1"].shape == x1.shape and got["K0"].shape == K0.shape
    # every instance sees identical inputs whichever rank owns it and instances never interact: bitwise equal
    assert np.array_equal(got["x1"], x1)
    assert np.array_equal(got["u0"], u0)
    assert np.array_equal(got["K0"], K0)
