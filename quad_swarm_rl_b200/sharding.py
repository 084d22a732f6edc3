"""Multi-GPU layout: envs are independent units, so the job shards them in contiguous blocks, one process per GPU,
with NO step-time collective.  The only cross-rank traffic is an optional metrics reduction."""
import torch.distributed as dist


def shard_range(total_envs, world_size, rank):
    """Contiguous block [lo, hi) of global env ids owned by `rank`; sizes differ by at most one."""
    base, rem = divmod(int(total_envs), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def reduce_metrics(vec, op='sum'):
    """All-reduce a small metrics vector (episode counters, reward sums) across ranks; identity when not distributed.
    Works on CUDA tensors with NCCL and on CPU tensors with gloo."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(vec, op=dist.ReduceOp.SUM if op == 'sum' else dist.ReduceOp.MAX)
    return vec
