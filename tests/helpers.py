"""Shared test helpers (layout conversion, comparisons)."""
import numpy as np


def to_time_major(rvp, n, nobs):
    """reference flat rvp[i + j*NOBS] -> u[i][j] (the device layout)."""
    return np.ascontiguousarray(np.asarray(rvp).reshape(n, nobs).T)


def relerr(a, b, floor=0.0):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    scale = max(float(np.max(np.abs(b))), floor, 1e-300)
    return float(np.max(np.abs(a - b))) / scale


def first_mismatch_step(A_dev, A_ref):
    """First time index where the ancestor matrices differ (None if identical)."""
    neq = np.any(A_dev != A_ref, axis=1)
    idx = np.nonzero(neq)[0]
    return None if idx.size == 0 else int(idx[0])
