"""The path's collectives: the batch is sharded by rank with no data-path exchange;
the only cross-rank data are the int64 PCK/AUC/EPE counters and the four f64 loss sums.

Mirrors the reference's helpers (all of which are unused call sites there):
  train/spawn_dist.py:68-80          all_reduce(values)   list -> f32 tensor -> barrier -> SUM
  train/distributed_utils.py:65-76   reduce_value(value, average=True)
Backend: NCCL over NVLink/NVSwitch on the GPU box, gloo in the CPU tests.  No barrier is issued
before the all-reduce (the reference's dist.barrier() only adds latency).
"""
import torch
import torch.distributed as dist


def _active():
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def shard_bounds(n, rank, world):
    """Contiguous batch shard [lo, hi) of rank `rank`."""
    return (n * rank) // world, (n * (rank + 1)) // world


def all_reduce(values, device=None):
    """spawn_dist.py:68-80: list of python numbers -> summed f32 tensor (on every rank)."""
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else "cpu"
    t = torch.tensor(values, dtype=torch.float32, device=device)
    if _active():
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def reduce_value(value, average=True):
    """distributed_utils.py:65-76: mean (or sum) of a tensor over the ranks (DDP loss logging)."""
    if not _active():
        return value
    with torch.no_grad():
        value = value.clone()
        dist.all_reduce(value)
        if average:
            value /= dist.get_world_size()
    return value


def all_reduce_loss_sums(sums):
    """Global-batch semantics of the balanced loss: SUM the f64 (S_pos, S_neg, N_pos, numel) of every
    shard, then finalise — equals the single-process reference on the concatenated batch."""
    if _active():
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    return sums


def all_reduce_counters(counters):
    """int64 SUM of the metric counters: bit-exact for any sharding."""
    if _active():
        dist.all_reduce(counters, op=dist.ReduceOp.SUM)
    return counters
