"""GPU tier: edge cases of the replay path through the C ABI -- no detections at all, empty launches, ragged
batch sizes around the CTA size, a filter fed non-finite inputs next to healthy ones, a detection on tick 0."""
import numpy as np
import pytest

import quadrotor_landing_b200 as q
from oracle import ekf_oracle as orc
from quadrotor_landing_b200 import scenario
from streams_np import noisy_streams, norm_rel, rotors_params

pytestmark = pytest.mark.gpu
TOL = 1e-9


def short(p, seconds=3.0):
    spec = scenario.default_spec()
    spec.duration_s, spec.hover_s = seconds, 1.0
    return scenario.generate(p, spec)


@pytest.mark.parametrize("multirate", [0, 1])
def test_no_detections_means_no_filter(multirate):
    """filter_update returns before touching anything until a tag has initialised the state (cpp:129-130)."""
    p = rotors_params(q.default_params(), multirate=bool(multirate))
    scn = short(p)
    N = 40
    st = noisy_streams(scn, N, seed=1)
    b = q.BatchEKF(p, N)
    x0, P0 = b.state(), b.cov()
    none = np.zeros((0,), dtype=np.int32)
    b.run(0, scn.T, st["imu"], none, np.zeros((0, 7, N)), np.zeros((0,)))
    assert np.array_equal(b.state(), x0) and np.array_equal(b.cov(), P0)
    assert not b.flags().any()
    # every arrival invalid: same thing
    b.run(0, scn.T, st["imu"], st["tag_step"], st["tag_pose"], st["tag_stamp"], np.zeros_like(st["tag_valid"]))
    assert np.array_equal(b.state(), x0) and not b.flags().any()
    # an empty launch is a no-op
    b.run(5, 0, st["imu"], st["tag_step"], st["tag_pose"], st["tag_stamp"], st["tag_valid"])
    assert np.array_equal(b.state(), x0)
    b.close()


@pytest.mark.parametrize("N", [1, 31, 224, 225, 449])
def test_ragged_batch_sizes(N):
    p = rotors_params(q.default_params())
    scn = short(p)
    st = noisy_streams(scn, N, seed=N, dropout=(200, 260))
    b = q.BatchEKF(p, N)
    ob = orc.Batch(orc.params_from(p), N)
    b.run(0, scn.T, st["imu"], st["tag_step"], st["tag_pose"], st["tag_stamp"], st["tag_valid"])
    ob.run(0, scn.T, st["imu"], st["tag_step"], st["tag_pose"], st["tag_stamp"], st["tag_valid"])
    assert norm_rel(b.state(), ob.state()) < TOL and norm_rel(b.cov(), ob.cov()) < TOL
    assert np.array_equal(b.flags()[0:5], ob.flags()[0:5])
    b.close()


def test_a_poisoned_filter_does_not_touch_its_neighbours():
    """Non-finite inputs in one filter: that filter's state goes non-finite and the statistics count it as
    diverged; every other filter is bit-identical to the clean run (filters never interact)."""
    p = rotors_params(q.default_params())
    scn = short(p, 4.0)
    N, bad = 70, 33
    st = noisy_streams(scn, N, seed=3)
    clean = q.BatchEKF(p, N)
    clean.run(0, scn.T, st["imu"], st["tag_step"], st["tag_pose"], st["tag_stamp"], st["tag_valid"])
    imu = st["imu"].copy()
    imu[300:, 0, bad] = np.nan
    b = q.BatchEKF(p, N)
    b.run(0, scn.T, imu, st["tag_step"], st["tag_pose"], st["tag_stamp"], st["tag_valid"])
    keep = np.arange(N) != bad
    assert np.array_equal(b.state()[:, keep], clean.state()[:, keep])
    assert np.array_equal(b.cov()[:, :, keep], clean.cov()[:, :, keep])
    assert not np.isfinite(b.state()[:, bad]).all()
    b.close(); clean.close()


def test_detection_on_the_first_tick_and_back_to_back_arrivals():
    p = rotors_params(q.default_params())
    p.limit_measurement_freq = 0                      # a correction on every arrival
    scn = short(p, 2.0)
    N = 33
    st = noisy_streams(scn, N, seed=8)
    steps = np.array([0, 1, 2, 3, 10, 11, 399], dtype=np.int32)
    pose = st["tag_pose"][: len(steps)]
    stamp = steps / p.update_freq
    b = q.BatchEKF(p, N)
    ob = orc.Batch(orc.params_from(p), N)
    for k0, n in ((0, 1), (1, 3), (4, 396)):
        b.run(k0, n, st["imu"], steps, pose, stamp)
        ob.run(k0, n, st["imu"], steps, pose, stamp)
        assert norm_rel(b.state(), ob.state()) < TOL and norm_rel(b.cov(), ob.cov()) < TOL
        assert np.array_equal(b.flags()[0:5], ob.flags()[0:5])
    assert ob.counts()[1] == 7 * N
    b.close()


def test_statistics_count_divergence():
    p = rotors_params(q.default_params())
    scn = short(p, 4.0)
    noise = q.default_noise()
    noise.sigma_accel = float("inf")                  # every synthetic IMU sample is non-finite
    N, stride = 64, 200
    b = q.BatchEKF(p, N)
    b.stats_configure(scn.T // stride, stride)
    b.run_monte_carlo(scn, noise)
    s = b.stats()
    assert s[-1, 18] == N and s[-1, 16] == 0          # all diverged, none sampled
    b.close()


def test_rejected_per_filter_override_leaves_the_handle_untouched():
    """A first qekf_set_filter_params call that fails validation must not switch the per-filter tables on (they would be
    uninitialised): the next replay still uses the handle-wide parameters."""
    p = rotors_params(q.default_params())
    p.multirate_ekf = 1
    scn = scenario.generate(p)
    N, T = 40, 600
    st = noisy_streams(scn, N, seed=4, T=T)
    ref = q.BatchEKF(p, N)
    ref.run(0, T, st["imu"], st["tag_step"], st["tag_pose"], st["tag_stamp"], st["tag_valid"])
    b = q.BatchEKF(p, N)
    bad = np.zeros((2, N)); bad[0] = -1.0                      # negative measurement_delay
    with pytest.raises(q.QekfError):
        b.set_filter_params(q.PF_DELAY, bad)
    b.run(0, T, st["imu"], st["tag_step"], st["tag_pose"], st["tag_stamp"], st["tag_valid"])
    assert norm_rel(b.state(), ref.state()) == 0 and norm_rel(b.cov(), ref.cov()) == 0
    ref.close(); b.close()


def test_set_params_reaches_the_fields_that_were_never_overridden():
    """Q is overridden per filter, R is not: a later qekf_set_params with a new R must be used by every filter
    (include/qekf.h: fields never overridden keep the handle-wide value)."""
    p = rotors_params(q.default_params())
    scn = scenario.generate(p)
    N, T = 33, 800
    st = noisy_streams(scn, N, seed=6, T=T)
    Q = np.array(list(p.Q_a) + list(p.Q_w) + list(p.Q_ab) + list(p.Q_wb))[:, None] * np.linspace(0.5, 2.0, N)[None, :]
    p2 = rotors_params(q.default_params())
    for i in range(3):
        p2.R_r[i] *= 4.0
        p2.R_ang[i] *= 0.25
    b = q.BatchEKF(p, N)
    b.set_filter_params(q.PF_Q, Q)
    b.set_params(p2)
    b.run(0, T, st["imu"], st["tag_step"], st["tag_pose"], st["tag_stamp"], st["tag_valid"])
    ob = orc.Batch(orc.params_from(p2), N)
    ob.set_filter_params(orc.PF_Q, Q)
    ob.run(0, T, st["imu"], st["tag_step"], st["tag_pose"], st["tag_stamp"], st["tag_valid"])
    assert norm_rel(b.state(), ob.state()) < 1e-9 and norm_rel(b.cov(), ob.cov()) < 1e-9
    b.close()


def test_forced_initialisation_before_any_tag_uses_the_identity_orientation():
    """initialize_state with no detection latched yet: apriltag_orien is the constructor's identity (cpp:14), so the
    nominal attitude becomes conj(q_vc) instead of a zero quaternion."""
    p = rotors_params(q.default_params())
    b = q.BatchEKF(p, 3)
    b.initialize_state(False)
    f = orc.Filter(orc.params_from(p))
    f.initialize_state(False)
    x = b.state()
    assert np.all(np.isfinite(x)) and norm_rel(x[:, 0], f.state()) < 1e-12
    b.close()
