"""Init-time constants of the spherical basis (host side, numpy/scipy only).

The reference derives them symbolically at construction (utils/sbf.py:13-61, ~15 s of sympy); only the
numbers are needed: zeros z_lm of the spherical Bessel functions j_l (kept in float32 there, utils/sbf.py:15)
and the normalisers N_lm = (0.5 j_{l+1}(z_lm)^2)^(-1/2) (utils/sbf.py:44-49).
"""
import functools

import numpy as np

from ._lib import SbfConsts

NUM_SPHERICAL, NUM_RADIAL = 7, 6


@functools.lru_cache(maxsize=None)
def sbf_tables(n=NUM_SPHERICAL, k=NUM_RADIAL):
    from scipy.optimize import brentq
    from scipy.special import spherical_jn

    z = np.zeros((n, k), dtype=np.float32)
    z[0] = np.arange(1, k + 1) * np.pi
    lo = np.arange(1, k + n) * np.pi            # j_0 zeros bracket the zeros of j_1, and so on upward
    for l in range(1, n):
        roots = np.array([brentq(lambda r: spherical_jn(l, r), float(lo[m]), float(lo[m + 1]))
                          for m in range(len(lo) - 1)], dtype=np.float32)
        z[l] = roots[:k]
        lo = roots
    z64 = z.astype(np.float64)
    norm = np.stack([1.0 / np.sqrt(0.5 * spherical_jn(l + 1, z64[l]) ** 2) for l in range(n)])
    return z64, norm


@functools.lru_cache(maxsize=None)
def sbf_consts_struct():
    z, nrm = sbf_tables()
    c = SbfConsts()
    for i, (a, b) in enumerate(zip(z.reshape(-1), nrm.reshape(-1))):
        c.zeros[i] = float(a)
        c.norm[i] = float(b)
    return c
