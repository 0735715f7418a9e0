"""The acceptance procedure pinned to the reference's OWN Python: tests/golden/transfer_refpy.json holds what
ControlTestCallback's episode loop, `Controller.stepinfo_SS()` (the reference's `calc_stepinfo`), `Controller.quality()`
and the error properties return when /root/reference's env/ctrl_env.py + core/controller.py + tools/general.py run over the
DLL (tests/golden/make_transfer_refpy.py).  Checked: the DLL-backed C layer's goldens are bit-identical, the host
helpers of the product reproduce the reference's `calc_stepinfo` / `calc_err` / `calc_exp_k`, the oracle's recorder
reproduces the random-action episodes, and (GPU) so does the in-kernel tracker."""
import json
import math
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
DEG = math.pi / 180
R = json.load(open(os.path.join(HERE, "golden", "transfer_refpy.json")))
G = json.load(open(os.path.join(HERE, "golden", "transfer_golden.json")))
DEGS = (5, -5, 10, -10)
KEYS = ("overshoot", "rise_time", "settling_time", "static_error")


def test_provenance():
    assert "tools/general.py" in R["_provenance"] and "unmodified" in R["_provenance"]


def test_dll_backed_goldens_are_bit_identical_to_reference_python():
    """transfer_golden.json (C env layer + recorder over the DLL) == the reference's own stepinfo_SS / quality."""
    for d in DEGS:
        a, b = R["episodes"][f"ADD_PROC_CONTROL/{d}"], G[str(d)]
        assert a["stepinfo"] == b["stepinfo"] and a["quality"] == b["quality"], d
        assert a["length"] == 400 and a["records"] == 2000


def test_host_helpers_reproduce_reference_functions():
    from b747_rl_ctrl_b200.tools.general import calc_err, calc_exp_k, calc_stepinfo
    for c in R["calc_stepinfo"]:
        assert calc_stepinfo(c["ys"], c["y_base"], ts=c["ts"]) == c["info"]
    for a, b, e in R["calc_err"]:
        assert calc_err(a, b) == e
    assert [calc_exp_k(0.8, 10), calc_exp_k(0.75, 0.15)] == R["calc_exp_k"]


def _actions(deg):
    rng = np.random.default_rng(deg + 100)
    return np.array([float(np.float32(rng.uniform(-0.3, 0.3))) for _ in range(400)])


def test_oracle_recorder_reproduces_random_action_episodes(oracle):
    """DIRECT_CONTROL episodes with random elevator commands: the restatement's recorder + stepinfo + quality."""
    cfg = oracle.make_cfg(reset_ref_mode=oracle.RESET_NONE)
    for d in DEGS:
        ref = R["episodes"][f"DIRECT_CONTROL/{d}"]
        ob = oracle.OracleBatch(cfg, 1)
        ob.reset_to([oracle.episode([0, 11000, 250, 0, 0, 0], vref=d * DEG)])
        v = ob.env(0)
        v.enable_storage(2001)
        ret = 0.0
        for a in _actions(d):
            _, r, done, _ = ob.step([a], auto_reset=False)
            ret += r[0]
        assert done[0] and ret == pytest.approx(ref["return"], rel=1e-9)
        info = v.stepinfo_SS()
        for k in KEYS:
            assert (info[k] is None) == (ref["stepinfo"][k] is None), (d, k)
            if info[k] is not None:
                assert info[k] == pytest.approx(ref["stepinfo"][k], rel=1e-8, abs=1e-9), (d, k)
        st = v.storage
        for k, vals in ref["samples"].items():
            assert np.allclose(st[k][49::250], vals, rtol=1e-8, atol=1e-9), (d, k)
        m = ob.model(0)
        assert math.exp(-6 * m.get("ITSE") / (20 * (d * DEG) ** 2)) == pytest.approx(ref["quality"], rel=1e-7, abs=1e-300)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype_name", ["F64", "F32"])
def test_gpu_tracker_reproduces_reference_python(dtype_name):
    """run_control_test (in-kernel recorder + online calc_stepinfo) on the same episodes."""
    from b747_rl_ctrl_b200 import engine as E
    from b747_rl_ctrl_b200.control_test import run_control_test
    acts = np.stack([_actions(d) for d in DEGS], axis=1)   # [400, 4]
    k = {"i": 0}

    def policy(obs):
        a = acts[min(k["i"], 399)]
        k["i"] += 1
        return a
    f32 = dtype_name == "F32"
    r = run_control_test(policy, [d * DEG for d in DEGS], ctrl_mode=E.MODE_DIRECT, dtype=getattr(E, dtype_name), record=True)
    for j, d in enumerate(DEGS):
        ref = R["episodes"][f"DIRECT_CONTROL/{d}"]
        assert r["length"][j] == 400
        assert r["return"][j] == pytest.approx(ref["return"], rel=2e-4 if f32 else 1e-9)
        for key in KEYS:
            want = ref["stepinfo"][key]
            got = r[key][j]
            assert (want is None) == bool(np.isnan(got)), (d, key)
            if want is not None:
                tol = dict(rel=2e-3, abs=0.0100001) if f32 else dict(rel=1e-8, abs=1e-9)
                assert got == pytest.approx(want, **tol), (d, key)
        # quality = exp(-6 ITSE / (tk vref^2)) spans 80 decades on these episodes: compare the exponent
        assert math.log(r["quality"][j]) == pytest.approx(math.log(ref["quality"]), rel=1e-4 if f32 else 1e-8)
        st = r["storage"][j]
        for key, vals in ref["samples"].items():
            assert np.allclose(st[key][49::250], vals, rtol=2e-4 if f32 else 1e-8, atol=2e-4 if f32 else 1e-9), (d, key)
