"""include/b747_params.h (the aero tables / constants compiled into the kernels) must equal the
reference DLL's initialised data."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_matches_dll(dllref):
    sys.path.insert(0, ROOT)
    from oracle import extract_params as X
    fresh = X.render(X.extract())
    committed = open(os.path.join(ROOT, "include", "b747_params.h")).read()
    assert fresh == committed


def test_header_shape():
    txt = open(os.path.join(ROOT, "include", "b747_params.h")).read()
    assert "#define B747_NP 298" in txt
    body = txt.split("#define B747_P_INIT {")[1].split("}")[0].replace("\\", "")
    vals = [float(x) for x in body.split(",") if x.strip()]
    assert len(vals) == 298
    assert vals[16] == 288.15 and vals[136] == 0.03 and vals[128] == 5.255875601466713
