"""Multi-GPU: environments are independent, so a batch of N envs shards contiguously over R ranks with
no data-path collective.  The Philox key of every draw includes the GLOBAL env id, so the
union of the shards is bit-identical to one N-env batch on one GPU, for any R.

The only collective is one all-gather of a 16-byte record per rank (env-steps, elapsed ns) for the
aggregated throughput counter -- NCCL over NVLink on GPUs, gloo in the CPU tests.
"""
import torch
import torch.distributed as dist


def shard_range(n_envs, rank, world_size):
    """Rank r owns global env ids [r*N/R, (r+1)*N/R)."""
    lo = (rank * n_envs) // world_size
    hi = ((rank + 1) * n_envs) // world_size
    return lo, hi


def gather_throughput(env_steps, elapsed_ns, device=None):
    """All-gather {env_steps:int64, elapsed_ns:int64}; returns (total_env_steps, max_elapsed_ns, per_rank)."""
    rec = torch.tensor([int(env_steps), int(elapsed_ns)], dtype=torch.int64, device=device)
    if not (dist.is_available() and dist.is_initialized()):
        return int(rec[0]), int(rec[1]), [rec.tolist()]
    out = [torch.zeros_like(rec) for _ in range(dist.get_world_size())]
    dist.all_gather(out, rec)
    per_rank = [o.tolist() for o in out]
    return sum(p[0] for p in per_rank), max(p[1] for p in per_rank), per_rank
