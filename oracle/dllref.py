"""ctypes driver for the reference DLL hosted by oracle/_ref/libb747_ref.so.

TEST INFRASTRUCTURE (oracle) -- not product code.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module.

`DllModel` gives the same global-variable view of the library that the
reference's `Model` builds with `in_dll` (core/model.py:124-164): three
`void(void)` entry points plus named `double` globals.  Every instance maps a
private, relocated copy of the image, mirroring the reference's
copy-the-DLL-per-Model isolation (core/model.py:99-110).
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(_HERE, "_ref", "libb747_ref.so")

# name -> length, exactly the data symbols core/model.py binds (+ the unbound exports)
SIGNALS = {
    "state": 6, "sim_time": 1, "vartheta_zh": 1, "U_com_PID": 1, "CXa": 1, "CYa": 1, "mz": 1,
    "K_alpha": 1, "dCm_ddeltaz": 1, "U_com": 1, "deltaz_RP": 1, "dvartheta": 1,
    "dvartheta_int": 1, "dvartheta_dt": 1, "dvartheta_dt_dt": 1, "TAE": 1, "ITAE": 1, "TSE": 1,
    "ITSE": 1, "AE": 1, "IAE": 1, "SE": 1, "ISE": 1, "alpha": 1, "V": 1, "Mach": 1,
}
PARAMS = {
    "state0": 6, "h_zh": 1, "use_RP": 1, "use_PID_SS": 1, "use_PID_CS": 1, "PID_SS": 4, "PID_CS": 4,
    "deltaz": 1, "vartheta": 1, "P": 1, "aero_err": 5, "Iz": 1, "S": 1, "c_": 1, "g": 1, "m0": 1,
    "use_RL": 1,
}
N_BLOCK_PARAMS = 298  # doubles in model_simple_P


def disabled() -> bool:
    """B747_NO_DLL_ORACLE=1 switches every path that executes the reference's DLL machine code off: tests that need it
    skip, bench.py's CPU legs fall back to the plain-C restatement (cpu_baseline.kind "port")."""
    return os.environ.get("B747_NO_DLL_ORACLE", "") not in ("", "0")


def available() -> bool:
    return os.path.exists(REF_SO) and not disabled()


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError(
                f"{REF_SO} not built; run `make -C oracle ref` where /root/reference is mounted")
        L = ctypes.CDLL(REF_SO)
        L.b747ref_open.restype = ctypes.c_void_p
        L.b747ref_close.argtypes = [ctypes.c_void_p]
        L.b747ref_sym.restype = ctypes.c_void_p
        L.b747ref_sym.argtypes = [ctypes.c_void_p, ctypes.c_char_p]
        L.b747ref_rva.restype = ctypes.c_void_p
        L.b747ref_rva.argtypes = [ctypes.c_void_p, ctypes.c_uint32]
        L.b747ref_call.argtypes = [ctypes.c_void_p]
        L.b747ref_call_n.argtypes = [ctypes.c_void_p, ctypes.c_long]
        _lib = L
    return _lib


def sandbox():
    """Drop what a numeric worker does not need before it runs the DLL in bulk (pe_host.c b747ref_sandbox: seccomp-bpf
    deny-list -- sockets, exec, ptrace, module / mount calls, file writes).  Irrevocable for the calling process, so only
    worker processes call it (bench.py's reference arm, the golden generators).  Returns True if the filter is installed."""
    L = lib()
    L.b747ref_sandbox.restype = ctypes.c_int
    return L.b747ref_sandbox() == 0


class DllModel:
    """One private instance of model_simple_win64.dll."""

    def __init__(self):
        L = lib()
        self._L = L
        self._h = L.b747ref_open()
        if not self._h:
            raise RuntimeError("b747ref_open failed")
        self._f_init = L.b747ref_sym(self._h, b"model_simple_initialize")
        self._f_step = L.b747ref_sym(self._h, b"model_simple_step")
        self._f_term = L.b747ref_sym(self._h, b"model_simple_terminate")
        self._v = {}
        for name, n in {**SIGNALS, **PARAMS}.items():
            addr = L.b747ref_sym(self._h, name.encode())
            if not addr:
                raise RuntimeError(f"export {name} not found")
            self._v[name] = (ctypes.c_double * n).from_address(addr)
        self.block_params = (ctypes.c_double * N_BLOCK_PARAMS).from_address(
            L.b747ref_sym(self._h, b"model_simple_P"))

    def close(self):
        if self._h:
            self._L.b747ref_close(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- global access ------------------------------------------------------
    def get(self, name):
        v = self._v[name]
        return v[0] if len(v) == 1 else list(v)

    def set(self, name, value):
        v = self._v[name]
        if len(v) == 1:
            v[0] = float(value)
        else:
            for i, x in enumerate(value):
                v[i] = float(x)

    def addr(self, name):
        return ctypes.addressof(self._v[name])

    def rva(self, rva):
        return self._L.b747ref_rva(self._h, rva)

    # -- entry points -------------------------------------------------------
    def initialize(self):
        self._L.b747ref_call(self._f_init)

    def step(self, n=1):
        if n == 1:
            self._L.b747ref_call(self._f_step)
        else:
            self._L.b747ref_call_n(self._f_step, n)

    def terminate(self):
        self._L.b747ref_call(self._f_term)
