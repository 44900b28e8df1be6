"""Multi-GPU Monte-Carlo job: one process per GPU, filters sharded by global id, statistics all-reduced.

Filters never interact (the reference has exactly one estimator instance,
quad_state_estimation/include/relative_pose_EKF_node.hpp:31), so the batch partitions into contiguous
blocks of global filter ids, one block per rank.  A filter's noise realisation depends on its GLOBAL id only
(csrc/ekf_synth.cuh), hence the per-filter results do not depend on the number of ranks.  The only exchange
step of the whole path is one sum all-reduce of the RMSE / NEES statistics [n_bins][STAT_DIM] per pass
(NCCL on the device buffer qekf_copy_stats_device fills; gloo in the CPU test tier).
"""
from __future__ import annotations

import copy


def shard_range(n_total: int, rank: int, world: int):
    """(first global id, count) of the contiguous block rank `rank` owns: ceil(n_total / world) filters per
    rank, the last ranks taking the remainder (possibly nothing)."""
    if world <= 0 or not (0 <= rank < world) or n_total < 0:
        raise ValueError("bad shard request: n_total=%r rank=%r world=%r" % (n_total, rank, world))
    per = -(-n_total // world)
    first = min(rank * per, n_total)
    return first, min(per, n_total - first)


def all_reduce_stats(stats, group=None):
    """Sum the statistics tensor over all ranks, in place (no-op outside torch.distributed)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)
    return stats


class ShardedMonteCarlo:
    """This rank's shard of an `n_total`-filter Monte-Carlo job.

    `make_batch(count)` builds the compute object for `count` filters: `BatchEKF` on a GPU box; the CPU test
    tier passes a stand-in with the same four methods (stats_configure, run_monte_carlo, state, stats_tensor)
    so that the sharding and reduction logic runs under gloo without a GPU.
    """

    def __init__(self, make_batch, n_total: int, rank: int, world: int, noise, n_bins: int, stride: int, empty_device=None):
        self.rank, self.world, self.n_total = rank, world, n_total
        self.empty_device = empty_device     # where an empty shard's zero statistics live (default: by backend)
        self.first, self.count = shard_range(n_total, rank, world)
        self.noise = copy.copy(noise)
        self.noise.first_global_id = noise.first_global_id + self.first
        self.batch = make_batch(self.count) if self.count > 0 else None
        self.n_bins, self.stride = n_bins, stride
        if self.batch is not None:
            self.batch.stats_configure(n_bins, stride)

    def run(self, scn, k0=0, n_steps=None, group=None):
        """Replay the scenario on the local shard and return the job-wide statistics tensor."""
        import torch
        if self.batch is not None:
            self.batch.run_monte_carlo(scn, self.noise, k0, n_steps)
            stats = self.batch.stats_tensor()
        else:
            # an empty shard (more ranks than filters) still takes part in the collective, with zeros on the device the
            # backend reduces on: NCCL needs a CUDA tensor on this rank's GPU, gloo a CPU one
            import torch.distributed as dist
            from ._native import STAT_DIM
            device = self.empty_device
            if device is None:
                on_nccl = dist.is_available() and dist.is_initialized() and dist.get_backend(group) == "nccl"
                device = torch.device("cuda", torch.cuda.current_device()) if on_nccl else torch.device("cpu")
            stats = torch.zeros((self.n_bins, STAT_DIM), dtype=torch.float64, device=device)
        return all_reduce_stats(stats, group)
