"""Model -- drop-in for the reference's ctypes wrapper (core/model.py:87-267), backed by the CUDA engine.

Same constructor, attribute surface and lifecycle:
    m = Model(use_PID_CS=False, initial_state=np.array([...]))
    m.hzh = 2000; m.vartheta_zh = -0.1; m.step(); m.state_dict['vartheta']; m.deltaz_real
The reference binds ~45 `double` globals of a per-instance copy of the DLL; here every Model owns a
one-environment float64 handle of libb747_b200 (state in HBM, stepped by the same sm_100a kernel
as the batched path).  The name traps of the reference are kept (core/model.py:129-164):
    Model.vartheta_ref -> signal `vartheta_zh`     Model.vartheta_zh -> parameter `vartheta`
    Model.deltaz_ref   -> signal `U_com_PID`       Model.deltaz_com  -> signal `U_com`
    Model.deltaz_real  -> signal `deltaz_RP`       Model.time        -> signal `sim_time`
There is no CPU path: constructing a Model without a CUDA device raises B747Error.
"""
import numpy as np

from .. import engine as E

_STATE_SIG = ["sig_state_x", "sig_state_y", "sig_state_Vx", "sig_state_Vy", "sig_state_vartheta", "sig_state_wz"]
_STATE0 = ["state0_x", "state0_y", "state0_Vx", "state0_Vy", "state0_vartheta", "state0_wz"]
_AERR = ["aerr0", "aerr1", "aerr2", "aerr3", "aerr4"]

# python attribute -> exported signal (core/model.py:129-151, 171-192)
_SIGNALS = {
    "time": "sim_time", "vartheta_ref": "vartheta_zh", "deltaz_ref": "U_com_PID", "deltaz_com": "U_com",
    "deltaz_real": "deltaz_RP", "CXa": "CXa", "CYa": "CYa", "mz": "mz", "Kalpha": "K_alpha",
    "dCm_ddeltaz": "dCm_ddeltaz", "dvartheta": "dvartheta", "dvartheta_int": "dvartheta_int",
    "dvartheta_dt": "dvartheta_dt", "dvartheta_dt_dt": "dvartheta_dt_dt", "TAE": "TAE", "ITAE": "ITAE",
    "TSE": "TSE", "ITSE": "ITSE", "AE": "AE", "IAE": "IAE", "SE": "SE", "ISE": "ISE",
}


class ModelView:
    """The Model attribute surface over env `index` of an existing float64 engine."""

    labels = ['x', 'y', 'Vx', 'Vy', 'vartheta', 'wz']
    dt = 0.01  # core/model.py:121

    def __init__(self, engine, index=0):
        self._e = engine
        self._i = index
        self.step_num = -1

    def _g(self, name):
        return float(self._e.get(name)[self._i])

    def _s(self, name, value):
        v = self._e.get(name)
        v[self._i] = float(value)
        self._e.set(name, v)

    # ---- signals -------------------------------------------------------------------------
    @property
    def state(self):
        # np.nan_to_num filter of core/model.py:167-168,200
        return np.nan_to_num(np.array([self._g(n) for n in _STATE_SIG]))

    @property
    def state_dict(self):
        st = self.state
        return dict((self.labels[i], st[i]) for i in range(len(self.labels)))

    # ---- parameters ----------------------------------------------------------------------
    @property
    def state0(self):
        return np.array([self._g(n) for n in _STATE0])

    @state0.setter
    def state0(self, value):
        for n, v in zip(_STATE0, value):
            self._s(n, v)

    def set_initial(self, state):
        self.state0 = state

    @property
    def hzh(self):
        return self._g("h_zh")

    @hzh.setter
    def hzh(self, v):
        self._s("h_zh", v)

    @property
    def deltaz(self):
        return self._g("deltaz")

    @deltaz.setter
    def deltaz(self, v):
        self._s("deltaz", v)

    @property
    def vartheta_zh(self):
        return self._g("vartheta")

    @vartheta_zh.setter
    def vartheta_zh(self, v):
        self._s("vartheta", v)

    @property
    def aero_err(self):
        return np.array([self._g(n) for n in _AERR])

    @aero_err.setter
    def aero_err(self, value):
        for n, v in zip(_AERR, value):
            self._s(n, v)

    @property
    def use_PID_CS(self):
        return 1.0 if (int(self._g("flags")) & 4) else 0.0

    @use_PID_CS.setter
    def use_PID_CS(self, v):
        f = int(self._g("flags"))
        self._s("flags", (f & ~4) | (4 if float(v) >= 1.0 else 0))

    # whole-batch tunables (one env per Model, so per-instance here)
    def _param(name, n=1):
        def get(self):
            return self._e.get_param(name, n)

        def set_(self, v):
            self._e.set_param(name, np.asarray(v, dtype=np.float64) if n > 1 else float(v))
        return property(get, set_)

    use_RP = _param("use_RP")
    use_PID_SS = _param("use_PID_SS")
    PID_SS = _param("PID_SS", 4)
    PID_CS = _param("PID_CS", 4)
    P = _param("P")
    del _param

    # ---- entry points ---------------------------------------------------------------------
    def initialize(self):
        """model_simple_initialize + step_num=-1, deltaz=0, vartheta_zh=0 (core/model.py:238-244)."""
        self._e.model_initialize()
        self.step_num = -1

    def step(self):
        self._e.model_step(1)
        self.step_num += 1

    def terminate(self):
        self._e.synchronize()


def _add_signal(pyname, sig):
    setattr(ModelView, pyname, property(lambda self, _s="sig_" + sig: self._g(_s)))


for _py, _sig in _SIGNALS.items():
    _add_signal(_py, _sig)


class Model(ModelView):
    """Stand-alone model instance (core/model.py:87-236)."""

    def __init__(self, model="model_simple", use_PID_SS=True, use_PID_CS=True, initial_state=None,
                 logging_path="model.log", use_RP=True, device=0):
        if model != "model_simple":
            # the reference's Model only binds symbols that exist in model_simple
            raise ValueError(f"unsupported model library: {model}")
        eng = E.BatchEngine(n_envs=1, dtype=E.F64, device=device, env_layer=False, export_signals=True,
                            reset_ref_mode=E.RESET_NONE, ctrl_type=E.CTRL_MANUAL, auto_reset=False, tk=float("inf"))
        super().__init__(eng, 0)
        self.model = model
        self._PID_initial = np.array(list(self.PID_CS) + list(self.PID_SS))
        if initial_state is not None:
            self.state0 = initial_state
        self.use_RP = float(use_RP)
        self.use_PID_CS = float(use_PID_CS)
        self.use_PID_SS = float(use_PID_SS)
        self.Pmax = self.P
        self.initialize()
