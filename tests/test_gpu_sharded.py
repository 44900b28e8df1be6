"""GPU tier: the sharded Monte-Carlo job (quadrotor_landing_b200/sharded.py) on CUDA handles.  One GPU emulates
the ranks one after the other (a real multi-rank run is `torchrun bench.py --gpus N`); the CPU tier runs the
same class under gloo with world_size 2 (tests/test_sharded_gloo.py)."""
import numpy as np
import pytest

import quadrotor_landing_b200 as q
from quadrotor_landing_b200.sharded import ShardedMonteCarlo
from test_sharded_gloo import _scenario

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world", [2, 3])
def test_shards_reproduce_the_whole_job(world):
    p, scn, noise = _scenario()
    n_total, stride = 301, 200
    nb = scn.T // stride
    whole = ShardedMonteCarlo(lambda c: q.BatchEKF(p, c), n_total, 0, 1, noise, nb, stride)
    ref = whole.run(scn).cpu().numpy()
    xs, tot = [], np.zeros_like(ref)
    for r in range(world):
        job = ShardedMonteCarlo(lambda c: q.BatchEKF(p, c), n_total, r, world, noise, nb, stride)
        tot += job.run(scn).cpu().numpy()
        xs.append(job.batch.state())
        job.batch.close()
    assert np.array_equal(np.concatenate(xs, axis=1), whole.batch.state())
    assert np.array_equal(tot[:, 16:19], ref[:, 16:19])
    assert np.max(np.abs(tot - ref)) <= 1e-12 * np.max(np.abs(ref))
    whole.batch.close()
