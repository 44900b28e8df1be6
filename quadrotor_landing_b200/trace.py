"""Trace export for a sampled subset of a Monte-Carlo batch.

The reference's offline-evaluation workflow records the node's topics with rosbag
(quad_state_estimation/launch/start_EKF_Cpp_rosbag_record.launch).  For one filter the ROS-free replay driver
(tools/replay_driver.cpp) writes those topics tick by tick; for a batch of a million filters this module samples
a few of them: the fused replay is cut into launches of `every` ticks (state, time skew and pending measurements
survive a cut, so the results are identical to a single launch) and after each launch the accessors of the
chosen filters are read back and appended to a CSV with the replay driver's column layout.
"""
from __future__ import annotations

import numpy as np

COLUMNS = (["filter", "tick", "t", "active", "px", "py", "pz", "qx", "qy", "qz", "qw"] + ["cov%d" % i for i in range(36)] +
           ["bias_ax", "bias_ay", "bias_az", "bias_wx", "bias_wy", "bias_wz", "vx", "vy", "vz", "ax", "ay", "az",
            "pred_length", "corrected", "obs_px", "obs_py", "obs_pz", "obs_qx", "obs_qy", "obs_qz", "obs_qw", "meas_delay"])
_POSE_IDX = [0, 1, 2, 6, 7, 8]          # rows / cols of cov_pert the node publishes (node.cpp:203-210)


def trace_monte_carlo(batch, scn, noise, filters, every: int, path: str, k0: int = 0, n_steps: int | None = None):
    """Run `batch.run_monte_carlo` over ticks [k0, k0 + n_steps) in launches of `every` ticks and write the topics of
    the filters in `filters` (local indices) after every launch to `path`.  Returns the number of rows written."""
    filters = [int(f) for f in filters]
    n_steps = scn.T - k0 if n_steps is None else n_steps
    p = batch.params
    ab_s, wb_s = np.array(list(p.ab_static)), np.array(list(p.wb_static))
    lo, hi = min(filters), max(filters) + 1
    rows = 0
    with open(path, "w") as out:
        out.write(",".join(COLUMNS) + "\n")
        k = k0
        while k < k0 + n_steps:
            n = min(every, k0 + n_steps - k)
            batch.run_monte_carlo(scn, noise, k, n)
            k += n
            x, P = batch.state(lo, hi - lo), batch.cov(lo, hi - lo)
            aux, fl = batch.aux(lo, hi - lo), batch.flags(lo, hi - lo)
            t = scn.spec.t_start + (k - 1) / p.update_freq
            for f in filters:
                j = f - lo
                vals = [f, k - 1, repr(float(t)), int(fl[3, j])]
                if fl[3, j]:
                    vals += [repr(float(v)) for v in x[0:3, j]] + [repr(float(v)) for v in x[6:10, j]]
                    vals += [repr(float(v)) for v in P[np.ix_(_POSE_IDX, _POSE_IDX)][:, :, j].reshape(-1)]
                    vals += [repr(float(v)) for v in x[10:13, j] + ab_s] + [repr(float(v)) for v in x[13:16, j] + wb_s]
                    vals += [repr(float(v)) for v in x[3:6, j]] + [repr(float(v)) for v in aux[0:3, j]]
                    vals += [int(fl[4, j]), int(fl[2, j])]
                    vals += [repr(float(v)) for v in aux[3:11, j]]
                out.write(",".join(str(v) for v in vals) + "\n")
                rows += 1
    return rows
