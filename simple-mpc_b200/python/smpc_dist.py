"""Multi-GPU partitioning of a batch of MPC instances.

Instances are independent: they share only read-only tables (robot model, OCP weights, gait stage table,
solver options) and never exchange data inside a solve. Rank r of G therefore owns a contiguous block of
instance indices and its own CUDA context; `torch.distributed` is used for exactly three things:

  * `broadcast_bytes`   - rank 0's shared tables to every rank, once at set-up (tens of KB);
  * `max_over_ranks`    - the timing reduction the bench contract asks for;
  * `gather_blocks`     - optional gather of the per-instance heads (x1, u0, K0) to rank 0.

There is deliberately no collective on the data path. Works over NCCL (device tensors) and gloo (CPU tensors,
used by tests/test_multirank_gloo.py with world_size 2).
"""
import numpy as np
import torch


def shard_range(total, world, rank):
    """Contiguous block [first, first + count) of `total` instances owned by `rank`: instance i lives on rank
    floor(i * world / total), i.e. block sizes differ by at most one."""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world of %d" % (rank, world))
    first = -(-rank * total // world)          # ceil(rank * total / world)
    last = -(-(rank + 1) * total // world)
    return first, last - first


def _dev(dist):
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")


def broadcast_bytes(dist, payload, src=0):
    """Every rank passes its own copy of the shared tables; returns rank `src`'s bytes and whether the local copy
    was identical (ranks that built a different robot table must not silently carry on)."""
    mine = torch.frombuffer(bytearray(payload), dtype=torch.uint8).to(_dev(dist))
    n = torch.tensor([mine.numel()], dtype=torch.int64, device=mine.device)
    dist.broadcast(n, src)
    ref = mine.clone() if int(n.item()) == mine.numel() else torch.empty(int(n.item()), dtype=torch.uint8, device=mine.device)
    dist.broadcast(ref, src)
    same = ref.numel() == mine.numel() and bool(torch.equal(ref, mine))
    return bytes(ref.cpu().numpy().tobytes()), same


def max_over_ranks(dist, values):
    t = torch.tensor(list(values), dtype=torch.float64, device=_dev(dist))
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t.cpu()]


def gather_blocks(dist, block, total, dst=0):
    """Gather per-rank blocks `[count_r, ...]` (ragged in the first dimension, sharded with `shard_range`) into one
    `[total, ...]` array on rank `dst`; other ranks get None."""
    world, rank = dist.get_world_size(), dist.get_rank()
    counts = [shard_range(total, world, r)[1] for r in range(world)]
    pad = max(counts)
    block = np.ascontiguousarray(block)
    assert block.shape[0] == counts[rank], (block.shape, counts, rank)
    buf = np.zeros((pad,) + block.shape[1:], dtype=block.dtype)
    buf[: counts[rank]] = block
    t = torch.from_numpy(buf).to(_dev(dist))
    outs = [torch.empty_like(t) for _ in range(world)] if rank == dst else None
    dist.gather(t, outs, dst=dst)
    if rank != dst:
        return None
    return np.concatenate([o.cpu().numpy()[:c] for o, c in zip(outs, counts)], axis=0)
