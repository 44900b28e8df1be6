"""CPU tier: the node's parameter files through the product's loader (csrc/preset.hpp, C ABI
qekf_params_from_yaml), against the values the reference's constructor would end up with
(relative_pose_EKF_node.cpp:35-136; presets restate config/relative_pose_EKF_{rotors,hardware}.yaml)."""
import numpy as np
import pytest

import quadrotor_landing_b200 as q
from oracle import ekf_oracle as orc


def test_rotors_preset():
    p = q.params_from_yaml("rotors_sim")
    assert (p.update_freq, p.measurement_freq) == (100.0, 15.0)
    assert (p.measurement_delay, p.measurement_delay_max, p.dyn_measurement_delay_offset) == (0.030, 0.200, 0.005)
    assert list(p.Q_a) == [0.0005] * 3 and list(p.Q_w) == [0.00005] * 3
    assert list(p.Q_ab) == [5.0e-5] * 3 and list(p.Q_wb) == [5.0e-6] * 3
    assert list(p.R_r) == [0.015, 0.015, 0.020] and list(p.R_ang) == [0.0015, 0.0015, 0.04]
    assert list(p.r_v_cv) == [0, 0, -0.073] and list(p.q_vc) == [0.70711, -0.70711, 0, 0]
    assert list(p.camera_K) == [241.4268, 0, 376.5, 0, 241.4268, 240.5, 0, 0, 1]
    assert (p.camera_width, p.camera_height, p.n_tags) == (752, 480, 1)
    assert p.tag_widths[0] == 0.8 and p.tag_in_view_margin == 0.02
    assert [p.limit_measurement_freq, p.corner_margin_enbl, p.est_bias, p.direct_orien_method, p.multirate_ekf,
            p.dynamic_meas_delay] == [1] * 6
    # covariance initialisation is not in the file: the node's param<> defaults (node.cpp:89-93)
    assert (p.r_cov_init, p.v_cov_init, p.ang_cov_init, p.ab_cov_init, p.wb_cov_init) == (0.1, 0.1, 0.15, 0.5, 0.1)


def test_hardware_preset_bundle_geometry():
    p = q.params_from_yaml("hardware_bundle")
    assert p.n_tags == 13 and p.limit_measurement_freq == 0 and p.tag_in_view_margin == 0.0
    assert (p.camera_width, p.camera_height) == (640, 480)                 # "640.0" in the file
    w = np.array(list(p.tag_widths))[:13]
    assert w[0] == 0.08382 and np.all(w[1:5] == 0.16764) and np.all(w[5:9] == 0.33528) and np.all(w[9:13] == 0.16764)
    pos = np.array(list(p.tag_positions))[:39].reshape(13, 3)
    assert np.all(pos[:, 2] == 0) and list(pos[5]) == [-0.244475, 0.244475, 0] and list(pos[12]) == [-0.314325, 0, 0]
    assert list(p.ab_static) == [0.20, -0.09, -0.03] and list(p.wb_static) == [-0.02, -0.01, 0.0]
    assert list(p.q_vc) == [-0.7035177, 0.7106742, 0.0014521, -0.0017207]
    assert p.measurement_delay_max == 0.350
    # the oracle accepts the struct as is (same layout) and derives its parameters from it
    f = orc.Filter(orc.params_from(p))
    assert f.n == 15


def test_missing_keys_take_the_node_defaults_and_errors_are_reported():
    p = q.params_from_yaml_text("update_freq: 250   # only this\nQ_a_diag: [1, 2,\n   3]\nest_bias: false\n")
    assert p.update_freq == 250.0 and p.measurement_freq == 10.0 and p.measurement_delay == 0.010
    assert list(p.Q_a) == [1, 2, 3] and p.est_bias == 0
    d = q.default_params()
    assert list(p.R_r) == list(d.R_r) and list(p.camera_K) == list(d.camera_K)     # getParam: member unchanged
    assert p.direct_orien_method == 0 and p.multirate_ekf == 0                      # node.cpp:62-64 defaults
    for bad in ("Q_a_diag: [1, 2]\n", "update_freq: fast\n", "n_tags: 99\n", "tag_widths: [0.1, 0.2\n", "est_bias: maybe\n",
                "just a line\n", "update_freq: -5\n"):
        with pytest.raises(q.QekfError):
            q.params_from_yaml_text(bad)
    with pytest.raises(q.QekfError):
        q.params_from_yaml("/nonexistent/file.yaml")
