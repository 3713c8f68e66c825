"""Drop-in for ``pyipn/rff.py``: same names, same argument meaning, evaluated on the B200.

``RFF`` (pyipn/rff.py:5-16) and ``RFF_multiscale`` (pyipn/rff.py:19-29) return a float64 ``(T,)`` array
like the reference.  They call ``ipn_rff_eval`` through ctypes; without the CUDA library / a B200 they raise.
"""
from __future__ import annotations

import numpy as np

from ._lib import default_context
from .params import fold_draws


def RFF(time, omega1, omega2, beta1, beta2, bw, scale):
    omega, a, b = fold_draws(omega1, omega2, beta1, beta2, bw, scale)
    time = np.ascontiguousarray(time, dtype=np.float64)
    return default_context().rff_eval(time, omega, a, b, np.zeros(1), np.ones(1))[0]


def RFF_multiscale(time, omega1, omega2, beta1, beta2, bw, scale1, scale2):
    scale = np.array([[scale1], [scale2]], dtype=np.float64)
    omega, a, b = fold_draws(omega1, omega2, beta1, beta2, bw, scale)
    time = np.ascontiguousarray(time, dtype=np.float64)
    return default_context().rff_eval(time, omega, a, b, np.zeros(1), np.ones(1))[0]
