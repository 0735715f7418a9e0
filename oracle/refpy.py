"""The reference's OWN Python layers, run here (TEST INFRASTRUCTURE -- oracle, not product code).

`load()` imports /root/reference/core/model.py, core/controller.py and env/ctrl_env.py unmodified, from where they lie,
on top of the reference DLL's own machine code:

  * `core/model.py` loads `core/model_simple.so` on Linux (a per-instance copy, `cdll.LoadLibrary`, `in_dll`,
    core/model.py:99-164) -- here that file is oracle/_ref/model_simple.so, the ELF face of model_simple_win64.dll
    (ref_dll/elf_shim.c).  The module derives the library folder from its `__file__`; /root/reference is read-only, so
    `__file__` of the imported module object is pointed at a scratch directory that holds the shim (no reference source
    is copied anywhere);
  * what the modules import but this image lacks is stubbed in `sys.modules`: `win32api`, `ctypes.WinDLL` (Windows-only
    calls at import time, core/model.py:25-29), `gym` (base class + `spaces.Box`), `optuna`, `matplotlib`, `openpyxl`
    (only names, never called on the step path).

`PhiloxRandom` replaces the module-level `random` / `np.random.normal` that Controller.reset draws from
(core/controller.py:148-191) with the counter-based stream the oracle and the CUDA path use (draw j of (seed, env,
episode)), so the reference's reset code, consuming that stream in ITS OWN order, must land on exactly the episode
`b747o_env_draw_episode` produces.

Only tests/golden/make_env_golden_refpy.py and tests/ use this module; it needs /root/reference (this container only).
"""
import ctypes
import math
import os
import shutil
import sys
import tempfile
import types

import numpy as np

REFERENCE = os.environ.get("B747_REFERENCE", "/root/reference")
_HERE = os.path.dirname(os.path.abspath(__file__))
SHIM = os.path.join(_HERE, "_ref", "model_simple.so")


def available():
    from . import dllref
    return os.path.exists(os.path.join(REFERENCE, "env", "ctrl_env.py")) and os.path.exists(SHIM) and not dllref.disabled()


class _Anything:
    """Stub attribute: callable, subscriptable, usable as a base class or type annotation."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, name):
        return _Anything()


class _StubModule(types.ModuleType):
    __path__ = []  # a package: submodule imports resolve through sys.modules

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        v = type(name, (_Anything,), {})
        setattr(self, name, v)
        return v


class _Box:
    """gym.spaces.Box as far as env/ctrl_env.py:92-101 uses it."""

    def __init__(self, low, high, shape=None, dtype=np.float32):
        self.shape = tuple(shape) if shape is not None else np.shape(low)
        self.low = np.broadcast_to(np.asarray(low, dtype=dtype), self.shape).copy()
        self.high = np.broadcast_to(np.asarray(high, dtype=dtype), self.shape).copy()
        self.dtype = np.dtype(dtype)


_loaded = None


def load():
    """Returns a namespace with the reference's modules: .model, .controller, .ctrl_env (imported once per process)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError("the reference tree or oracle/_ref/model_simple.so is missing (make -C oracle ref)")
    for name in ("win32api", "optuna", "matplotlib", "matplotlib.pyplot", "openpyxl", "openpyxl.drawing",
                 "openpyxl.chart", "openpyxl.chart.axis", "openpyxl.utils", "openpyxl.utils.units", "openpyxl.drawing.line",
                 "openpyxl.chart.shapes", "openpyxl.chart.text", "openpyxl.drawing.text"):
        if name not in sys.modules:
            sys.modules[name] = _StubModule(name)
    if "gym" not in sys.modules:
        gym = _StubModule("gym")
        gym.Env = type("Env", (), {})
        spaces = _StubModule("gym.spaces")
        spaces.Box = _Box
        gym.spaces = spaces
        sys.modules["gym"], sys.modules["gym.spaces"] = gym, spaces
    had_windll = hasattr(ctypes, "WinDLL")
    if not had_windll:
        ctypes.WinDLL = lambda *a, **k: _Anything()
    # the reference's top-level packages are `core`, `env`, `tools`: import them from the reference tree itself
    sys.path.insert(0, REFERENCE)
    try:
        import core.model as r_model           # noqa: E402
        import core.controller as r_controller  # noqa: E402
        import env.ctrl_env as r_env            # noqa: E402
    finally:
        sys.path.remove(REFERENCE)
        if not had_windll:
            del ctypes.WinDLL
    assert os.path.realpath(r_env.__file__).startswith(os.path.realpath(REFERENCE)), "not the reference's module"
    # Model.__init__ copies <folder of model.py>/model_simple.so into <folder>/tmp_models/<uuid>.so and loads the copy
    scratch = tempfile.mkdtemp(prefix="b747_refpy_")
    os.makedirs(os.path.join(scratch, "core"))
    shutil.copyfile(SHIM, os.path.join(scratch, "core", "model_simple.so"))
    r_model.__file__ = os.path.join(scratch, "core", "model.py")
    # the reference attaches a FileHandler("model.log") per Model; keep the working directory clean
    r_model.logging.FileHandler = lambda *a, **k: r_model.logging.NullHandler()
    _loaded = types.SimpleNamespace(model=r_model, controller=r_controller, ctrl_env=r_env, scratch=scratch)
    return _loaded


# ---- the Philox stream of the oracle / CUDA path behind the `random` interface -----------------------------------------
def _philox4x32(ctr, key):
    c0, c1, c2, c3 = ctr
    k0, k1 = key
    M = 0xffffffff
    for _ in range(10):
        p0, p1 = 0xD2511F53 * c0, 0xCD9E8D57 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & M, p1 & M, ((p0 >> 32) ^ c3 ^ k1) & M, p0 & M
        k0, k1 = (k0 + 0x9E3779B9) & M, (k1 + 0xBB67AE85) & M
    return c0, c1, c2, c3


def uniform53(seed, env, episode, draw):
    """Draw `draw` of (seed, env, episode): b747_common.cuh uniform53 / oracle b747o_uniform53."""
    w = _philox4x32((env & 0xffffffff, env >> 32, episode, draw >> 1), (seed & 0xffffffff, seed >> 32))
    a, b = w[2 * (draw & 1)] >> 5, w[2 * (draw & 1) + 1] >> 6
    return (a * 67108864.0 + b) * (1.0 / 9007199254740992.0)


class PhiloxRandom:
    """Stands in for the `random` module and `np.random` inside core/controller.py.  Sequential draws 0, 1, 2, ... for
    uniform / choice (the order Controller.reset calls them in), draws 10 + 2 i, 11 + 2 i for the i-th Gaussian of a
    reset (Box-Muller), exactly the indices b747o_env_draw_episode uses."""

    def __init__(self, seed, env_id):
        self.seed, self.env_id = seed, env_id
        self.episode = -1
        self.begin_episode()
        self.random = self  # `np.random` look-alike: self.random.normal
        self.log = []

    def begin_episode(self):
        self.episode += 1
        self._j = 0
        self._g = 0

    def _u(self):
        u = uniform53(self.seed, self.env_id, self.episode, self._j)
        self._j += 1
        return u

    def uniform(self, a, b):
        return a + (b - a) * self._u()   # CPython's random.uniform

    def choice(self, seq):
        return seq[0] if self._u() < 0.5 else seq[1]

    def normal(self, mean, sd, size=None):
        i = self._g
        self._g += 1
        u1 = 1.0 - uniform53(self.seed, self.env_id, self.episode, 10 + 2 * i)
        u2 = uniform53(self.seed, self.env_id, self.episode, 11 + 2 * i)
        return mean + sd * (math.sqrt(-2.0 * math.log(u1)) * math.cos(2.0 * math.pi * u2))

    def seed_(self, *_):
        pass


class _NpProxy:
    """`np` as core/controller.py sees it, with `np.random` replaced."""

    def __init__(self, rng):
        self._rng = rng
        self.random = rng

    def __getattr__(self, name):
        return getattr(np, name)


def patch_rng(ref, rng):
    """Route Controller.reset's draws (module globals `random` and `np` of core/controller.py) to `rng`."""
    ref.controller.random = rng
    ref.controller.np = _NpProxy(rng)


def unpatch_rng(ref):
    import random as _random
    ref.controller.random = _random
    ref.controller.np = np


def episode_of(ctrl):
    """What the reference's Controller.reset decided, read back from the live objects: (state0, use_ctrl, vref, href,
    osc (A, f) or None, aero_err)."""
    m = ctrl.model
    s0 = np.array(m.state0, dtype=np.float64)
    aero = np.array(m.aero_err, dtype=np.float64)
    vref, href, osc = 0.0, 11000.0, None
    f = ctrl.vartheta_func
    if f is not None:
        cl = {n: c.cell_contents for n, c in zip(f.__code__.co_freevars, f.__closure__ or ())}
        if "A1" in cl:
            osc = ([cl["A1"], cl["A2"], cl["A3"]], [cl["f1"], cl["f2"], cl["f3"]])
        else:
            vref = float(f(0.0))
    if ctrl.h_func is not None:
        href = float(ctrl.h_func(0.0))
    return s0, bool(ctrl.use_ctrl), vref, href, osc, aero
