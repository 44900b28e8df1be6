"""The synthetic tag-detection front-end of the Monte-Carlo generator (qekf_noise_spec.edge_loss / range_*):
detections are lost when the landing bundle leaves the image and the pose noise grows with the camera-to-tag
range.  CPU tier: the product's generator instantiated for the host against the independent numpy restatement
(oracle/noise_np.py); GPU tier: the device generator against the same, and a filter replay on the result."""
import numpy as np
import pytest

import quadrotor_landing_b200 as q
from oracle import ekf_oracle as orc
from oracle import noise_np
from quadrotor_landing_b200 import scenario
from streams_np import norm_rel, rotors_params


def low_pass_scenario(p):
    """A descent to 0.45 m with 0.35 m of lateral sway: near the ground the 0.8 m tag no longer fits the image."""
    spec = scenario.default_spec()
    spec.duration_s, spec.hover_s = 10.0, 2.0
    spec.z_start, spec.z_end = 2.5, 0.45
    spec.sway_ax, spec.sway_ay = 0.35, 0.25
    return scenario.generate(p, spec)


def frontend_noise(first=0):
    n = q.default_noise()
    n.first_global_id = first
    n.edge_loss = 1
    n.range_ref, n.range_exp_pos, n.range_exp_ang = 1.5, 2.0, 1.0
    return n


def check_against_numpy(st, p, scn, noise, ids):
    ref = noise_np.synthesize(noise, scn.imu_clean, scn.tag_step, scn.tag_pose_clean, ids, params=p)
    assert np.array_equal(st["tag_valid"], ref["tag_valid"])
    assert np.max(np.abs(st["tag_pose"] - ref["tag_pose"])) < 2e-5 * noise.sigma_tag_pos * 6
    lost = 1.0 - st["tag_valid"].mean()
    assert 0.05 < lost < 0.9, lost                     # the scenario really loses detections near the ground
    # range dependence: the position noise near the ground (range < range_ref) is smaller than at altitude
    rng_ = np.linalg.norm(scn.tag_pose_clean[:, 0:3], axis=1)
    d = st["tag_pose"][:, 0:3] - scn.tag_pose_clean[:, 0:3, None]
    far, near = rng_ > 2.0, rng_ < 1.0
    assert d[far].std() > 2.5 * d[near].std()
    return ref


def test_host_generator_frontend_matches_numpy():
    import host_core as hc
    p = rotors_params(q.default_params())
    scn = low_pass_scenario(p)
    noise = frontend_noise(first=77)
    st = hc.synthesize(scn, noise, 2, 24, params=p)
    check_against_numpy(st, p, scn, noise, 77 + 2 + np.arange(24))
    # the hardware bundle: small inner tags keep the bundle detectable closer to the ground than one big tag
    ph = q.params_from_yaml("hardware_bundle")
    ph.update_freq, ph.multirate_ekf = 200.0, 0
    scn_h = low_pass_scenario(ph)
    one = q.params_from_yaml("hardware_bundle")
    one.n_tags = 1
    one.tag_widths[0] = 0.8
    assert noise_np.bundle_in_image(scn_h.tag_pose_clean, ph).mean() > noise_np.bundle_in_image(scn_h.tag_pose_clean, one).mean()


@pytest.mark.gpu
def test_device_generator_frontend_and_replay():
    p = rotors_params(q.default_params())
    scn = low_pass_scenario(p)
    noise = frontend_noise(first=1 << 33)
    N = 96
    b = q.BatchEKF(p, N)
    st = b.synthesize_streams(scn, noise, 0, N)
    check_against_numpy(st, p, scn, noise, noise.first_global_id + np.arange(N))
    # the fused Monte-Carlo kernel sees exactly the realisation it dumps; the oracle replays the dump
    b.run_monte_carlo(scn, noise)
    ob = orc.Batch(orc.params_from(p), N)
    ob.run(0, scn.T, st["imu"], st["tag_step"], st["tag_pose"], st["tag_stamp"], st["tag_valid"])
    assert norm_rel(b.state(), ob.state()) < 1e-9 and norm_rel(b.cov(), ob.cov()) < 1e-9
    assert np.array_equal(b.flags()[0:5], ob.flags()[0:5])
    # long predictions while the tag is out of view
    assert b.flags()[4].max() > 50
    b.close()
