"""bench.py pieces that do not need a GPU: the M2 summary (potrf + K-build against their rooflines)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def test_m2_summary_arithmetic_and_robustness():
    ph = {"kbuild": 0.2, "potrf": 10.0}
    r = bench.m2_summary(8192, ph, 6552.6, 1677.8, 37.0, 8)
    assert abs(r["potrf_tflops"] - (8192 ** 3 / 3) / 10e-3 / 1e12) < 1e-9
    assert abs(r["kbuild_gbs"] - 8 * 8192 ** 2 / 2 / 0.2e-3 / 1e9) < 1e-6
    assert abs(r["potrf_frac_of_dmma_peak"] - r["potrf_tflops"] / 37.0) < 1e-12
    assert abs(r["peaks"]["int8_equiv_fp64_tflops"] - 2 * 1677.8 / 36) < 1e-9
    assert abs(r["kbuild_frac_of_hbm"] - r["kbuild_gbs"] / 6552.6) < 1e-12
    # never raises, whatever is missing
    assert "error" in bench.m2_summary(8192, {}, None, 1590.0, 0.0, 8)
    r = bench.m2_summary(8192, ph, None, 1590.0, 0.0, 0)
    assert r["potrf_frac_of_dmma_peak"] is None and r["kbuild_frac_of_hbm"] is None


def test_workload_inputs_are_the_survey_ones():
    X, y, params = bench.make_inputs("c2")
    assert X.shape == (8192, 8) and y.shape == (8192, 1) and abs(float(y.mean())) < 1e-12
    assert list(params) == [1.0 / 8, 1.0, 0.01]
    assert X.flags["F_CONTIGUOUS"] and y.flags["F_CONTIGUOUS"]
