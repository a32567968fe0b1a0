"""Env-batch sharding over the GPUs of one box.

The reference splits its E emulators over W worker processes with ``np.split``
(fed_gym/agents/paac/runners.py:18-19,65-66: contiguous, equal slices, E % W == 0 required).
Here rank r of g owns the contiguous slice of GLOBAL env ids [r*E/g, (r+1)*E/g): that offset goes
into ``SwarmParams.env_id_offset`` so every env's Philox stream is keyed by its global id and a
trajectory does not depend on g.  The step path has no collective; the helpers below are the
only cross-rank traffic of a benchmark / rollout (a timing MAX and scalar sums for logging).
"""
import torch
import torch.distributed as dist


def shard_envs(total_envs, world_size, rank):
    """-> (first_global_env_id, n_local_envs).  Like np.split, refuses uneven splits."""
    total_envs, world_size, rank = int(total_envs), int(world_size), int(rank)
    if world_size < 1 or not 0 <= rank < world_size:
        raise ValueError("bad rank %d / world %d" % (rank, world_size))
    if total_envs % world_size:
        raise ValueError("array split does not result in an equal division: %d envs over %d ranks"
                         % (total_envs, world_size))
    per = total_envs // world_size
    return rank * per, per


def global_env_ids(total_envs, world_size, rank, device=None):
    first, n = shard_envs(total_envs, world_size, rank)
    return torch.arange(first, first + n, dtype=torch.int64, device=device)


def _active():
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def max_over_ranks(value, device=None):
    """Device-timed milliseconds -> the slowest rank's figure (every rank gets it)."""
    if not _active():
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value, device=None):
    if not _active():
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def gather_shards(local, device=None):
    """Concatenate per-rank env-major tensors in rank order (= global env id order)."""
    if not _active():
        return local
    out = [torch.empty_like(local) for _ in range(dist.get_world_size())]
    dist.all_gather(out, local.contiguous())
    return torch.cat(out, dim=0)


def allreduce_mean_(flat):
    """In-place mean over ranks of a flat tensor (the PAAC learner's 8.84 MB FP32 gradient): one all-reduce."""
    if _active():
        dist.all_reduce(flat)
        flat /= dist.get_world_size()
    return flat
