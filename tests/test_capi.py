"""The C-ABI libraries load and export every symbol the headers declare; without a GPU the entry
points fail loudly instead of falling back (no compute calls here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def libs():
    from b747_rl_ctrl_b200 import _lib, build
    build.build_native()
    return ctypes.CDLL(_lib.LIB_PATH), ctypes.CDLL(_lib.SCALAR_LIB_PATH)


def _declared_functions(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(b747_[a-z0-9_]+|model_simple_[a-z]+)\s*\(", txt)))


def test_batched_abi_symbols(libs):
    lib, scal = libs
    names = _declared_functions("b747.h")
    assert len(names) >= 24
    for n in names:
        assert hasattr(lib, n), n
        assert hasattr(scal, n), n  # model_simple.so is self-contained


def test_scalar_abi_symbols(libs):
    _, scal = libs
    for n in _declared_functions("b747_scalar.h"):
        assert hasattr(scal, n), n
    # the data symbols core/model.py:129-164 binds with in_dll (+ the unbound exports)
    data = {"state": 6, "sim_time": 1, "vartheta_zh": 1, "U_com_PID": 1, "CXa": 1, "CYa": 1, "mz": 1, "K_alpha": 1,
            "dCm_ddeltaz": 1, "U_com": 1, "deltaz_RP": 1, "dvartheta": 1, "dvartheta_int": 1, "dvartheta_dt": 1,
            "dvartheta_dt_dt": 1, "TAE": 1, "ITAE": 1, "TSE": 1, "ITSE": 1, "AE": 1, "IAE": 1, "SE": 1, "ISE": 1,
            "state0": 6, "h_zh": 1, "use_RP": 1, "use_PID_SS": 1, "use_PID_CS": 1, "PID_SS": 4, "PID_CS": 4,
            "deltaz": 1, "vartheta": 1, "P": 1, "aero_err": 5, "Iz": 1, "S": 1, "c_": 1, "g": 1, "m0": 1, "use_RL": 1,
            "alpha": 1, "V": 1, "Mach": 1}
    # where the reference tree is mounted, the list comes from core/model.py itself: every `in_dll(self.dll, "<name>")`
    ref_model = "/root/reference/core/model.py"
    if os.path.exists(ref_model):
        src = open(ref_model, encoding="utf-8").read()
        bound = re.findall(r"""(\(real_T\*(\d+)\)|real_T)\.in_dll\(self\.dll,\s*['"](\w+)['"]\)""", src)
        assert len(bound) >= 34
        for _, n_elem, name in bound:
            assert data.get(name) == int(n_elem or 1), f"core/model.py binds {name}[{n_elem or 1}]"
        funcs = set(re.findall(r"""getattr\(self\.dll,\s*f["']\{model\}_(\w+)["']\)""", src))
        assert funcs == {"initialize", "step", "terminate"}
    for n, k in data.items():
        v = (ctypes.c_double * k).in_dll(scal, n)
        assert len(v) == k
    # .data defaults of the DLL
    assert list((ctypes.c_double * 6).in_dll(scal, "state0")) == [0, 11000, 259.1667, 0, 0, 0]
    assert list((ctypes.c_double * 4).in_dll(scal, "PID_SS")) == [-5.9151, -1.2404, -6.6927, 58.0826]
    assert ctypes.c_double.in_dll(scal, "Iz").value == 67300000.0
    assert ctypes.c_double.in_dll(scal, "use_RP").value == 1.0


def test_struct_layout_matches_ctypes(libs):
    from b747_rl_ctrl_b200 import _lib
    lib, _ = libs
    assert lib.b747_abi_info(0) == _lib.ABI_VERSION
    assert lib.b747_abi_info(1) == ctypes.sizeof(_lib.Cfg)
    assert lib.b747_abi_info(2) == ctypes.sizeof(_lib.Episode)


def test_field_table(libs):
    lib, _ = libs
    lib.b747_field_name.restype = ctypes.c_char_p
    n = lib.b747_n_fields()
    names = [lib.b747_field_name(i).decode() for i in range(n)]
    assert len(set(names)) == n
    for must in ("h", "q0", "q3", "Vx", "deltaz", "vartheta", "h_zh", "state0_x", "sig_U_com_PID", "sig_ITSE", "tick"):
        assert must in names
        assert lib.b747_field_index(must.encode()) == names.index(must)
    assert lib.b747_field_index(b"nope") == -1


def test_no_cpu_fallback(libs):
    """On a box without a CUDA device creation must fail with B747_ERR_CUDA -- never a CPU path."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from b747_rl_ctrl_b200 import B747Error, engine
    with pytest.raises(B747Error, match="no CUDA device"):
        engine.BatchEngine(n_envs=4)
    from b747_rl_ctrl_b200.core.model import Model
    with pytest.raises(B747Error):
        Model()


def test_argument_validation(libs):
    from b747_rl_ctrl_b200 import _lib, engine
    lib, _ = libs
    L = _lib.load()
    h = ctypes.c_void_p()
    cfg = engine.make_cfg(n_envs=0)
    assert L.b747_create(ctypes.byref(cfg), ctypes.byref(h)) == _lib.ERR_ARG
    cfg = engine.make_cfg(n_envs=4)
    cfg.abi_version = 99
    assert L.b747_create(ctypes.byref(cfg), ctypes.byref(h)) == _lib.ERR_ARG
    cfg = engine.make_cfg(n_envs=4, ctrl_type=engine.CTRL_AUTO)  # random reset without an NN in the loop
    assert L.b747_create(ctypes.byref(cfg), ctypes.byref(h)) == _lib.ERR_ARG
    assert b"random reset" in L.b747_last_error()
    cfg = engine.make_cfg(n_envs=4, dtype=_lib.F32, env_layer=False)
    assert L.b747_create(ctypes.byref(cfg), ctypes.byref(h)) == _lib.ERR_ARG


def test_headers_are_plain_c(tmp_path):
    """The boundary is a C ABI: every header under include/ compiles as C11 on its own (no torch, no C++)."""
    import subprocess
    for h in ("b747.h", "b747_scalar.h", "b747_scalar_legacy.h", "b747_params.h"):
        src = tmp_path / f"use_{h}.c"
        src.write_text(f'#include "{os.path.join(ROOT, "include", h)}"\nint main(void) {{ return 0; }}\n')
        r = subprocess.run(["gcc", "-std=c11", "-Wall", "-Werror", "-fsyntax-only", str(src)], capture_output=True, text=True)
        assert r.returncode == 0, (h, r.stderr)


def test_shipped_library_is_sm100a_native():
    """Static evidence in the shipped .so: sm_100a cubins only, TMA bulk copies (UBLKCP) + mbarrier (SYNCS) in the staged
    step kernel, packed FP32 (FFMA2) in the model step, and -- by design, the path is not a contraction -- no tensor-core
    opcodes (profiles/r2_sass_histogram.md is the full histogram)."""
    import shutil
    import subprocess
    from b747_rl_ctrl_b200 import _lib
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not on PATH")
    sass = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"arch = (sm_\w+)", sass))
    assert archs == {"sm_100a"}, archs
    ops = set(re.findall(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*)", sass, flags=re.M))
    assert {"UBLKCP", "SYNCS", "FFMA2", "DFMA", "MUFU", "SHFL", "VOTE"} <= ops
    assert not {o for o in ops if o.startswith(("HMMA", "UTCHMMA", "UTCQMMA", "HGMMA", "IMMA"))}
