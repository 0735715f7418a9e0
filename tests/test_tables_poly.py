"""Host-side checks of the f32 path's ingredients (no GPU): the merged-axis look-up tables against the DLL's
look2_binlx / look1 semantics (dll@0x1000), and the generated polynomial header."""
import ctypes
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_merged_axis_tables_reproduce_look2():
    from b747_rl_ctrl_b200 import _lib
    L = _lib.load()
    out = (ctypes.c_double * 5)()
    # inside the breakpoint ranges: the only difference is the float32 rounding of the table entries
    assert L.b747_selftest_tables(400000, 0, out) == 0
    assert max(out) < 1e-6, list(out)
    # far outside (end cells extrapolate linearly, like look2_binlx): still the same function
    assert L.b747_selftest_tables(400000, 1, out) == 0
    assert max(out) < 1e-5, list(out)


def test_poly_header_is_current_and_accurate():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_poly.py"), "--check"], capture_output=True, text=True)
    assert r.returncode == 0, "b747_poly.h is stale: run tools/gen_poly.py\n" + r.stdout + r.stderr
    rep = eval(r.stdout.strip().splitlines()[-1])
    assert rep["sin_abs"] < 2e-7 and rep["cos_abs"] < 2e-7 and rep["atan_abs"] < 1e-7 and rep["rho_rel"] < 4e-7


def test_kernel_uses_every_fitted_coefficient():
    """Regression: the kernel once dropped the last sin/cos coefficient (1e-5 error at 90 deg of pitch)."""
    import re
    hdr = open(os.path.join(ROOT, "b747_rl_ctrl_b200", "csrc", "b747_poly.h")).read()
    src = open(os.path.join(ROOT, "b747_rl_ctrl_b200", "csrc", "b747_model_mx.cuh")).read()
    for name in re.findall(r"constexpr float ((?:SIN|COS|ATAN|RHO)\d)\b", hdr):
        assert "poly::" + name in src, name
