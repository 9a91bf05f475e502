"""Rank sharding of environments and the episode-statistics reduce.

Environments are independent, so rank r of R owns the contiguous global ids
[r*E/R, (r+1)*E/R) and the data path has no collective.  The only exchange is one all-reduce of
<= 8 doubles per episode (the sums src/train.py:181-192 accumulates): NCCL over NVLink on GPUs,
gloo in the CPU tests.
"""
import torch
import torch.distributed as dist

_SUM_KEYS = ("rewards", "target_tracking_reward", "boundary_punishment", "duplicate_tracking_punishment",
             "covered_sum", "env_steps")


def shard_envs(total_envs, rank, world_size):
    """-> (n_local, global id of the first local env); sizes differ by at most one."""
    lo = total_envs * rank // world_size
    hi = total_envs * (rank + 1) // world_size
    return hi - lo, lo


def reduce_episode_stats(stats, device=None, group=None):
    """All-reduce the dict returned by BatchedEnvironment.episode_stats(): SUM for the sums, MAX for
    covered_max.  Returns a new dict; a no-op when torch.distributed is not initialised."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return dict(stats)
    dev = device if device is not None else ("cuda" if dist.get_backend(group) == "nccl" else "cpu")
    s = torch.tensor([stats[k] for k in _SUM_KEYS], dtype=torch.float64, device=dev)
    mx = torch.tensor([stats["covered_max"]], dtype=torch.float64, device=dev)
    dist.all_reduce(s, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=group)
    out = {k: float(v) for k, v in zip(_SUM_KEYS, s.tolist())}
    out["covered_max"] = float(mx.item())
    return out


def episode_summary(stats, n_uav):
    """The per-episode scalars src/train.py:187-192 logs, from (reduced) statistics."""
    agent_steps = max(stats["env_steps"] * n_uav, 1.0)
    return {"return": stats["rewards"] / agent_steps,
            "target_tracking_return": stats["target_tracking_reward"] / agent_steps,
            "boundary_punishment_return": stats["boundary_punishment"] / agent_steps,
            "duplicate_tracking_punishment_return": stats["duplicate_tracking_punishment"] / agent_steps,
            "average_covered_targets": stats["covered_sum"] / max(stats["env_steps"], 1.0),
            "max_covered_targets": stats["covered_max"]}
