"""Host-side (float64) folding of the reference's parameter layout into the canonical device form.

Reference layout (pyipn/fit.py:62-106): ``omega1, omega2, beta1, beta2`` are ``(k, S)`` with the draw as the
FAST axis (strided columns), ``scale`` is ``(S,)`` (single scale, pyipn/rff.py:11-14) or ``(2, S)``
(multiscale, pyipn/rff.py:25-27), ``bw`` is ``(S,)``.

Canonical form used by libipn_b200 (include/ipn.h):  ``log f(tau) = sum_j a_j cos(w_j tau) + b_j sin(w_j tau)``
with ``J = 2k`` and, per draw, ``w = [bw*omega1 ; bw*omega2]``, ``a = [s1*beta1 ; s2*beta1]``,
``b = [s1*beta2 ; s2*beta2]`` stored ``(S, J)`` C-order (one draw per row, contiguous).
"""
from __future__ import annotations

import numpy as np


def fold_draws(omega1, omega2, beta1, beta2, bw, scale):
    """(k, S) reference arrays -> (omega, a, b) each (S, 2k) float64 C-contiguous.

    ``scale`` may be ``(S,)`` / scalar (single scale) or ``(2, S)`` (multiscale)."""
    omega1 = np.asarray(omega1, dtype=np.float64)
    omega2 = np.asarray(omega2, dtype=np.float64)
    beta1 = np.asarray(beta1, dtype=np.float64)
    beta2 = np.asarray(beta2, dtype=np.float64)
    if omega1.ndim == 1:
        omega1, omega2, beta1, beta2 = (x[:, None] for x in (omega1, omega2, beta1, beta2))
    k, S = omega1.shape
    for name, x in (("omega2", omega2), ("beta1", beta1), ("beta2", beta2)):
        if x.shape != (k, S):
            raise ValueError(f"{name}: expected shape {(k, S)}, got {x.shape}")
    bw = np.broadcast_to(np.asarray(bw, dtype=np.float64), (S,))
    scale = np.asarray(scale, dtype=np.float64)
    if scale.ndim == 2:
        if scale.shape != (2, S):
            raise ValueError(f"scale: expected shape {(2, S)}, got {scale.shape}")
        s1, s2 = scale[0], scale[1]
    else:
        s1 = s2 = np.broadcast_to(scale, (S,))
    omega = np.empty((S, 2 * k))
    a = np.empty((S, 2 * k))
    b = np.empty((S, 2 * k))
    omega[:, :k] = (omega1 * bw).T
    omega[:, k:] = (omega2 * bw).T
    a[:, :k] = (beta1 * s1).T
    a[:, k:] = (beta1 * s2).T
    b[:, :k] = (beta2 * s1).T
    b[:, k:] = (beta2 * s2).T
    return omega, a, b
