"""GPU pipeline with depth triangulation before the sub-filter (`triangulate_pre_subfilter`, on in 7 of the reference's 8
estimator configs; Feature::Triangulate feature.cpp:686-751, helpers.cpp:103-372) against the REFERENCE'S OWN ESTIMATOR on the
point-cloud streams of tests/test_reference_pin.py (`tri_*` cases: all five triangulation methods, N = 89 and 203).  The
triangulation itself is host work (csrc/triangulate.h, once per feature); what is checked here is that the CUDA pipeline fed
by it — sub-filter kernel on the triangulated states, bad-triangulation priors, selection, gating, update — reproduces the
reference's id tables exactly and its trajectory / covariance to the tolerances of tests/test_gpu_estimator.py.
(File name sorts after the other GPU suites on purpose: this row was added after the last GPU minute of round 1.)"""
import numpy as np
import pytest

import test_gpu_estimator as TG
import test_reference_pin as RP

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", [c[0] for c in RP.CASES if c[0].startswith("tri_")])
def test_pcw_trajectory_with_triangulation_matches_the_reference_estimator(case, tmp_path):
    TG.test_pcw_trajectory_matches_the_reference_estimator(case, tmp_path)


def test_unknown_triangulation_method_is_an_error():
    from xivo_b200 import pyxivo, sim

    cfg = sim.load_cfg(RP.CFG)
    cfg.update(RP._tri("no_such_method"))
    with pytest.raises(pyxivo.XivoError, match="Triangulation"):
        pyxivo.Batch(cfg, n_seq=1, max_groups=4, max_features=14)
