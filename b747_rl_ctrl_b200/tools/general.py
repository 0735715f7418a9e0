"""Host-side helpers the Controller facade shares with the reference's tools/general.py: the reward constants
(`calc_exp_k`, :32-33), the relative error (`calc_err`, :35-43), the step-response figures (`calc_stepinfo`, :46-61)
and the `Storage` container (:315-329).  The per-step work they sit on -- recording after every model step and the
online form of calc_stepinfo -- runs in the kernels (b747_common.cuh `trace_model_step`); these functions only
re-express the same definitions for arbitrary recorded arrays (e.g. a non-constant reference)."""
import math

stepinfo_template = {'overshoot': None, 'static_error': None, 'rise_time': None, 'settling_time': None}


def calc_exp_k(rk, xk):
    return -math.log(rk) / xk


def calc_err(x1, x2):
    err = x1 - x2
    if x2 != 0:
        err /= x2
    elif x1 != 0:
        err /= x1
    else:
        err = 0
    return abs(err)


def calc_stepinfo(ys, y_base, error_band=0.05, ts=None):
    """Overshoot [% of y_base], rise time (first sample -- the last one excluded -- whose normalised response
    reaches 1 - band), settling time (last sample outside the +-band), static error; times count from ts[0]."""
    ys = list(ys)
    n = len(ys)
    info = dict(stepinfo_template)
    if y_base != 0:
        info['overshoot'] = ((max(ys) if y_base > 0 else min(ys)) - y_base) / y_base * 100
    span = y_base - ys[0]
    if ts is not None:
        lo, hi = 1 - error_band, 1 + error_band
        for i in range(n - 1):
            if (ys[i] - ys[0]) / span >= lo:
                info['rise_time'] = ts[i] - ts[0]
                break
        for i in range(n - 1, -1, -1):
            r = (ys[i] - ys[0]) / span
            if r <= lo or r >= hi:
                info['settling_time'] = ts[i] - ts[0]
                break
    info['static_error'] = abs(ys[-1] - y_base)
    return info


class Storage:
    """Named lists of recorded values (tools/general.py:315-329; plotting / xlsx export are offline tools and
    stay out of scope)."""

    def __init__(self, data=None):
        self.storage = {} if data is None else data

    def record(self, name, value):
        self.storage.setdefault(name, []).append(value)

    def clear(self, name):
        del self.storage[name]

    def clear_all(self):
        self.storage = {}
