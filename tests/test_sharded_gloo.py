"""CPU tier, world_size 2 over gloo: the multi-GPU path's host logic -- contiguous sharding by global filter
id and the all-reduce of the Monte-Carlo statistics (quadrotor_landing_b200/sharded.py) -- with the product's
per-filter code instantiated for the host standing in for the GPU batch.  On a GPU box the same class drives
BatchEKF and NCCL (tests/test_gpu_sharded.py, bench.py)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class HostStandIn:
    """BatchEKF's Monte-Carlo surface on top of tests/host_core (same templates as the CUDA kernel)."""

    def __init__(self, params, count):
        import host_core as hc
        self.hb = hc.HostBatch(params, count)
        self.acc = None

    def stats_configure(self, n_bins, stride):
        self.acc = np.zeros((32, n_bins, 20))
        self.stride = stride

    def run_monte_carlo(self, scn, noise, k0=0, n_steps=None):
        self.hb.run_mc(scn, noise, k0, n_steps, stats=self.acc, stride=self.stride)

    def stats_tensor(self):
        return torch.from_numpy(self.acc.sum(axis=0).copy())

    def state(self):
        return self.hb.state()


def _job(p, scn, noise, n_total, rank, world, n_bins, stride):
    from quadrotor_landing_b200.sharded import ShardedMonteCarlo
    job = ShardedMonteCarlo(lambda c: HostStandIn(p, c), n_total, rank, world, noise, n_bins, stride)
    stats = job.run(scn)
    return job, stats


def _scenario():
    import quadrotor_landing_b200 as q
    from quadrotor_landing_b200 import scenario
    from streams_np import rotors_params
    p = rotors_params(q.default_params())
    spec = scenario.default_spec()
    spec.duration_s, spec.hover_s = 4.0, 1.0
    scn = scenario.generate(p, spec)
    noise = q.default_noise()
    noise.first_global_id = 1000
    noise.dropout_k0, noise.dropout_k1 = 300, 380
    noise.rand_dropout_len, noise.rand_dropout_lo, noise.rand_dropout_hi = 60, 50, 600
    return p, scn, noise


def _worker(rank, world, port, n_total, out_dir):
    for pth in (ROOT, HERE):
        if pth not in sys.path:
            sys.path.insert(0, pth)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        p, scn, noise = _scenario()
        stride = 200
        job, stats = _job(p, scn, noise, n_total, rank, world, scn.T // stride, stride)
        x = job.batch.state() if job.batch is not None else np.zeros((16, 0))
        np.savez(os.path.join(out_dir, "rank%d.npz" % rank), stats=stats.numpy(), x=x, first=job.first, count=job.count)
    finally:
        dist.destroy_process_group()


def test_shard_range_partitions_every_size():
    from quadrotor_landing_b200.sharded import shard_range
    for n in (0, 1, 7, 8, 9, 1000, 1 << 20):
        for world in (1, 2, 3, 8):
            blocks = [shard_range(n, r, world) for r in range(world)]
            assert sum(c for _, c in blocks) == n
            nxt = 0
            for first, count in blocks:
                assert first == nxt or count == 0
                nxt = first + count
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def test_two_ranks_reproduce_the_single_rank_job(tmp_path):
    """Filters [0, 11) on one rank == [0, 6) + [6, 11) on two ranks: per-filter states bit-identical, the
    all-reduced statistics identical on both ranks and equal to the single-rank ones."""
    n_total, world = 11, 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n_total, str(tmp_path)), nprocs=world, join=True)
    parts = [np.load(os.path.join(str(tmp_path), "rank%d.npz" % r)) for r in range(world)]
    assert [int(pp["first"]) for pp in parts] == [0, 6] and [int(pp["count"]) for pp in parts] == [6, 5]
    assert np.array_equal(parts[0]["stats"], parts[1]["stats"])
    p, scn, noise = _scenario()
    stride = 200
    job, stats = _job(p, scn, noise, n_total, 0, 1, scn.T // stride, stride)
    assert np.array_equal(np.concatenate([pp["x"] for pp in parts], axis=1), job.batch.state())
    ref = stats.numpy()
    assert np.array_equal(parts[0]["stats"][:, 16:19], ref[:, 16:19])          # sample / hit / divergence counts
    assert np.max(np.abs(parts[0]["stats"] - ref)) <= 1e-12 * np.max(np.abs(ref))
    assert ref[-1, 16] == n_total


def test_empty_shard_takes_part_in_the_reduction(tmp_path):
    """More ranks than filters: the rank that owns nothing contributes zeros (on the backend's device) and every rank
    still ends with the job-wide statistics."""
    n_total, world = 1, 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n_total, str(tmp_path)), nprocs=world, join=True)
    parts = [np.load(os.path.join(str(tmp_path), "rank%d.npz" % r)) for r in range(world)]
    assert [int(pp["count"]) for pp in parts] == [1, 0]
    assert np.array_equal(parts[0]["stats"], parts[1]["stats"]) and parts[1]["stats"][-1, 16] == 1
