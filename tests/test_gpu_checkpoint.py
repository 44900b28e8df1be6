"""GPU tier: exact checkpoint / resume through the C ABI (qekf_export_state / qekf_import_state).

Run T/2 ticks, export, destroy the handle, create a fresh one, import, run the rest: every array the accessors
expose must be BIT-identical to one uninterrupted run -- single-rate, delayed fusion with fixed and with dynamic
delay (where the lagged checkpoint, the IMU ring, nh / hpos / hlen are part of the state; the reference's
x_hist / u_hist / P_hist vectors, relative_pose_EKF.cpp:196-264), the per-tick interface (latched IMU sample and
tag pose), and a Monte-Carlo run with statistics accumulators."""
import numpy as np
import pytest

import quadrotor_landing_b200 as q
from quadrotor_landing_b200 import scenario
from streams_np import noisy_streams, rotors_params
from test_multirate_host import delayed_scenario

pytestmark = pytest.mark.gpu


def snapshot(b):
    return [b.state(), b.cov(), b.aux(), b.flags()]


def same(a, b):
    return all(np.array_equal(x, y, equal_nan=True) for x, y in zip(a, b))


@pytest.mark.parametrize("multirate,dynamic,precision", [(0, 0, 64), (1, 0, 64), (1, 1, 64), (1, 1, 32)])
def test_resume_is_bit_identical(multirate, dynamic, precision):
    p = rotors_params(q.default_params(), multirate=bool(multirate), dynamic_delay=bool(dynamic))
    scn = delayed_scenario(p, 0.042 if dynamic else 0.030) if multirate else scenario.generate(p)
    N, T, cut = 300, 2400, 1187          # the cut falls between two corrections, with history entries pending
    st = noisy_streams(scn, N, seed=31, T=T, dropout=(700, 950), random_dropout_ticks=150)
    args = (st["imu"], st["tag_step"], st["tag_pose"], st["tag_stamp"], st["tag_valid"])

    ref = q.BatchEKF(p, N, precision=precision)
    ref.run(0, cut, *args)
    ref.run(cut, T - cut, *args)
    want = snapshot(ref)
    want_counts = ref.step_counts()
    ref.close()

    a = q.BatchEKF(p, N, precision=precision)
    a.run(0, cut, *args)
    blob = a.export_state()
    mid = snapshot(a)
    a.close()

    b = q.BatchEKF(p, N, precision=precision)
    b.import_state(blob)
    assert same(snapshot(b), mid)
    b.run(cut, T - cut, *args)
    assert same(snapshot(b), want)
    assert b.step_counts() == want_counts
    if multirate:
        assert (want[3][5] > 1).any()     # x_hist.size() is part of what resumed
    b.close()


def test_import_rejects_a_blob_that_does_not_fit():
    p = rotors_params(q.default_params())
    a = q.BatchEKF(p, 64)
    blob = a.export_state()
    a.close()
    for other in (q.BatchEKF(p, 96), q.BatchEKF(p, 64, precision=q.QEKF_FP32),
                  q.BatchEKF(rotors_params(q.default_params(), multirate=True), 64)):
        with pytest.raises(q.QekfError):
            other.import_state(blob)
        other.close()
    c = q.BatchEKF(p, 64)
    with pytest.raises(q.QekfError):
        c.import_state(blob[:1000])
    bad = blob.copy()
    bad[0] ^= 0xFF
    with pytest.raises(q.QekfError):
        c.import_state(bad)
    c.import_state(blob)
    c.close()


def test_per_tick_interface_resumes_with_latched_inputs():
    """The N = 1 drop-in: the latched IMU sample (zero-order hold, relative_pose_EKF_node.cpp:144-151) and the
    latched, not yet fused tag pose survive the checkpoint."""
    p = rotors_params(q.default_params(), multirate=True, dynamic_delay=True)
    scn = delayed_scenario(p, 0.042)
    st = noisy_streams(scn, 1, seed=3, T=600)
    arrivals = {int(s): m for m, s in enumerate(st["tag_step"])}

    def ticks(b, k0, k1, feed_imu_from=0):
        for k in range(k0, k1):
            if k in arrivals:
                m = arrivals[k]
                b.set_tag(st["tag_pose"][m, 0:3, 0], st["tag_pose"][m, 3:7, 0], float(st["tag_stamp"][m]))
            if k >= feed_imu_from:
                b.set_imu(st["imu"][k, 0:3, 0], st["imu"][k, 3:6, 0])
            b.filter_update(scn.spec.t_start + k / p.update_freq)

    cut = next(k for k in sorted(arrivals) if k > 300) + 1     # a tag is latched but the frequency gate has not fused it
    ref = q.BatchEKF(p, 1)
    ticks(ref, 0, 600, 0)
    a = q.BatchEKF(p, 1)
    ticks(a, 0, cut)
    blob = a.export_state()
    a.close()
    b = q.BatchEKF(p, 1)
    b.import_state(blob)
    # the first resumed tick gets no new IMU sample: it must reuse the imported latch, as the uninterrupted run
    # would if the IMU callback had not fired in between
    ref2 = q.BatchEKF(p, 1)
    ticks(ref2, 0, cut)
    ticks(ref2, cut, cut + 1, feed_imu_from=cut + 1)
    ticks(b, cut, cut + 1, feed_imu_from=cut + 1)
    assert same(snapshot(b), snapshot(ref2))
    ticks(b, cut + 1, 600)
    ticks(ref2, cut + 1, 600)
    assert same(snapshot(b), snapshot(ref2))
    for h in (ref, ref2, b):
        h.close()


def test_monte_carlo_statistics_resume():
    p = rotors_params(q.default_params())
    scn = scenario.generate(p)
    noise = q.default_noise()
    noise.seed = 77
    noise.rand_dropout_len, noise.rand_dropout_lo, noise.rand_dropout_hi = 200, 300, 1500
    N, T, cut = 1000, 2000, 900
    ref = q.BatchEKF(p, N)
    ref.stats_configure(T // 100, 100)
    ref.run_monte_carlo(scn, noise, 0, cut)
    ref.run_monte_carlo(scn, noise, cut, T - cut)
    want, want_stats = snapshot(ref), ref.stats()
    ref.close()

    a = q.BatchEKF(p, N)
    a.stats_configure(T // 100, 100)
    a.run_monte_carlo(scn, noise, 0, cut)
    blob = a.export_state()
    a.close()
    b = q.BatchEKF(p, N)                  # statistics are configured by the import
    b.import_state(blob)
    b.run_monte_carlo(scn, noise, cut, T - cut)
    assert same(snapshot(b), want)
    # floating-point atomics accumulate in launch order: sums agree to rounding, counts exactly
    got = b.stats()
    assert np.array_equal(got[:, 16:19], want_stats[:, 16:19])
    assert np.allclose(got, want_stats, rtol=1e-12, atol=0)
    b.close()
