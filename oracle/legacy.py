"""ctypes driver of the reference's LEGACY dynamics library, core/model_win64.dll, hosted natively by
oracle/_ref/libb747_legacy.so (pe_host.c; msvcrt's asin / mem* / malloc resolved by name, everything else trapped).

TEST INFRASTRUCTURE (oracle) -- not product code.  Establishes what SURVEY.md 8f N4 needs: the March-2022 DLL is the same
dynamics as model_simple behind another symbol surface (tests/test_legacy_model.py), and pins lib/model.so to it.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LEGACY_SO = os.path.join(_HERE, "_ref", "libb747_legacy.so")

# the double data symbols of model_win64.dll's export table: name -> length
SIGNALS = {"state": 6, "sim_time": 1, "vartheta_zh": 1, "deltaz_ref": 1, "deltaz_com": 1, "deltaz_real": 1, "CXa": 1,
           "CYa": 1, "mz": 1, "K_alpha": 1, "dCm_ddeltaz": 1, "dvartheta": 1, "dvartheta_int": 1, "dvartheta_dt": 1,
           "dvartheta_dt_dt": 1, "TAE": 1, "ITAE": 1, "TSE": 1, "ITSE": 1, "AE": 1, "IAE": 1, "SE": 1, "ISE": 1}
PARAMS = {"state0": 6, "h_zh": 1, "use_PID_SS": 1, "use_PID_CS": 1, "use_RL": 1, "PID_SS": 4, "PID_CS": 4, "deltaz": 1,
          "vartheta": 1, "P": 1, "aero_err": 5, "I": 3, "S": 1, "c_": 1, "g": 1, "m0": 1}
FUNCTIONS = ("model_initialize", "model_step", "model_terminate")


def available():
    from . import dllref
    return os.path.exists(LEGACY_SO) and not dllref.disabled()


class LegacyDll:
    """One private instance of model_win64.dll (get / set / initialize / step like dllref.DllModel)."""

    def __init__(self):
        L = ctypes.CDLL(LEGACY_SO)
        L.b747ref_open.restype = ctypes.c_void_p
        L.b747ref_sym.restype = ctypes.c_void_p
        L.b747ref_sym.argtypes = [ctypes.c_void_p, ctypes.c_char_p]
        L.b747ref_call.argtypes = [ctypes.c_void_p]
        L.b747ref_call_n.argtypes = [ctypes.c_void_p, ctypes.c_long]
        self._L, self._h = L, L.b747ref_open()
        if not self._h:
            raise RuntimeError("cannot map model_win64.dll")
        self._f = {n: L.b747ref_sym(self._h, n.encode()) for n in FUNCTIONS}
        self._v = {}
        for name, k in {**SIGNALS, **PARAMS}.items():
            addr = L.b747ref_sym(self._h, name.encode())
            if not addr:
                raise RuntimeError(f"export {name} not found")
            self._v[name] = (ctypes.c_double * k).from_address(addr)

    def get(self, name):
        a = self._v[name]
        return a[0] if len(a) == 1 else list(a)

    def set(self, name, value):
        a = self._v[name]
        if len(a) == 1:
            a[0] = float(value)
        else:
            for i, x in enumerate(value):
                a[i] = float(x)

    def initialize(self):
        self._L.b747ref_call(self._f["model_initialize"])

    def step(self, n=1):
        self._L.b747ref_call_n(self._f["model_step"], n)
