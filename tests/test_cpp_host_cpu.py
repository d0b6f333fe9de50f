"""The C++ host side of the drop-in (gpc_b200/cpp: CGpB200 : CGp, CGplvmB200 : CGplvm) without a GPU: what must hold on
any machine.  The driver oracle/_ref/cgp_b200_check (tests/cpp/cgp_b200_check.cpp, built by oracle/build_ref.sh against
the unmodified reference) runs the reference class and the drop-in class on the same data in one process."""
import json
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CHECK = os.path.join(ROOT, "oracle", "_ref", "cgp_b200_check")


def _run(*args, expect_ok=True):
    if not os.path.exists(CHECK):
        pytest.skip("oracle/_ref/cgp_b200_check not built (python __graft_entry__.py in the build container)")
    out = subprocess.run([CHECK] + [str(a) for a in args], capture_output=True, text=True, timeout=600)
    if expect_ok:
        assert out.returncode == 0, out.stderr[-2000:]
        lines = [l for l in out.stdout.splitlines() if not l.startswith("Warning:")]
        return json.loads("\n".join(lines))
    return out


@pytest.mark.parametrize("mode,N,D,d", [("gp", 60, 2, 1), ("gp", 45, 3, 2), ("gplvm", 50, 2, 4)])
def test_component_outside_the_device_path_falls_through_to_the_reference_host_code(mode, N, D, d):
    """ratquad is not a device kernel: every call of the drop-in class must land in the inherited reference
    implementation -- bit-identical results, no device evaluation, no CUDA needed."""
    r = _run(mode, N, D, d, 3, "ratquad,bias,white", 1 if (d == 1 or mode == "gplvm") else 0, 0, 12)
    assert r["on_device"] == 0 and r["device_evals"] == 0
    assert r["ll_ref"] == r["ll_dev"] == r["ll_dev_again"]
    for k in ("g", "out", "std", "opt"):
        assert r[k + "_ref"] == r[k + "_dev"], k
    assert r["opt_ll_ref"] == r["opt_ll_dev"]
    assert r["opt_ref"] != r["g_ref"] and r["opt_ll_ref"] > r["ll_ref"]  # the optimiser did move


def test_device_path_fails_loudly_without_a_gpu():
    """No CPU fallback for a model the device path covers: without a CUDA device the drop-in class throws
    ndlexceptions::Error carrying gpc_last_error() (it never silently computes on the host)."""
    import ctypes
    try:
        ndev = ctypes.CDLL(os.path.join(ROOT, "gpc_b200", "libgpc_b200.so")).gpc_device_count()
    except OSError:
        pytest.skip("libgpc_b200.so not built")
    if ndev > 0:
        pytest.skip("a CUDA device is present")
    out = _run("gp", 40, 2, 1, 3, "rbf,bias,white", 0, 0, 0, expect_ok=False)
    assert out.returncode == 1
    assert "gpc_b200:" in out.stderr
