"""GPU pipeline with depth refinement of in-state candidates (`use_depth_opt`, on in cfg/tumvi_cam1.json and cfg/phab_calibration.json of the
reference; Feature::RefineDepth feature.cpp:299-420, call sites manager.cpp:387-395, :431-440, :504-536) against the REFERENCE'S OWN ESTIMATOR on the
`dopt_*` point-cloud sequences of tests/test_reference_pin.py.  The refinement is host work (a 3x3 Gauss-Newton per candidate); what is checked here is the
CUDA pipeline fed by it — covariance blocks set from the refined Hessian, slot edits, gating, update — with the tolerances of tests/test_gpu_estimator.py
(the all-view cases, whose Gauss-Newton amplifies rounding by up to 1e7, at 1e-5 m).
(File name sorts after the other GPU suites on purpose: this row was added after the last GPU minute of round 1.)"""
import numpy as np
import pytest

import test_reference_pin as RP
from xivo_b200 import pyxivo, sim

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", [c[0] for c in RP.CASES if c[0].startswith("dopt_")])
def test_pcw_trajectory_with_depth_refinement_matches_the_reference_estimator(case, tmp_path):
    name, G, F, duration, seed, sim_depths, over, offset = next(c for c in RP.CASES if c[0] == case)
    cfg = sim.load_cfg(RP.CFG)
    cfg.update(over)
    ref, how = RP.reference_result(name, over, G, F, duration, seed, sim_depths, offset, tmp_path)
    msgs, _ = RP.stream(cfg, duration, seed, offset)
    b = pyxivo.Batch(cfg, n_seq=1, max_groups=G, max_features=F)
    b.init_with_sim_depths()
    tol_t, tol_r = (1e-5, 1e-6) if "all" in case else (1e-7, 1e-8)
    k = 0
    for kind, ts, p in msgs:
        if kind == "imu":
            b.inertial_meas(ts, p[0], p[1])
        else:
            b.visual_meas_pointcloud(ts, p[0], p[1])
            g = b.gsb(0)
            assert sorted(b.instate_features(0)["ids"].tolist()) == [int(x) for x in ref["ids"][k] if x >= 0], f"{how}: frame {k}"
            assert b.counters(0)["gauge_group"] == int(ref["gauge"][k])
            assert np.abs(g[:, 3] - ref["gsb"][k][:, 3]).max() <= tol_t and np.abs(g[:, :3] - ref["gsb"][k][:, :3]).max() <= tol_r, f"{how}: frame {k}"
            k += 1
    assert k == len(ref["gsb"])
    P = b.P(0)
    assert np.abs(P - ref["P"]).max() <= (1e-4 if "all" in case else 1e-7) * np.abs(ref["P"]).max()
    b.close()
