"""Population-vectorised calibration shared by the models' ``fit()``.

The reference hands ``scipy.optimize.differential_evolution`` a scalar ``_loss`` that runs one
member per call (``rrmpg/models/hbvedu.py:305-346``).  Here the whole trial population of a
generation is one ensemble: ``vectorized=True, updating='deferred'`` makes scipy pass a
``(k, S)`` matrix and the fused-objective kernel returns the S mean squared errors without ever
materialising qsim.  The returned object is the same ``OptimizeResult``.
"""
import numpy as np
from scipy import optimize


def as_population(X):
    """scipy's (k,) or (k, S) trial matrix -> C-contiguous [S, k] parameter matrix."""
    X = np.asarray(X, dtype=np.float64)
    if X.ndim == 1:
        X = X[:, None]
    return np.ascontiguousarray(X.T)


def minimise(loss, bounds, args, de_kwargs=None):
    kwargs = dict(vectorized=True, updating='deferred')
    if de_kwargs:
        kwargs.update(de_kwargs)
    return optimize.differential_evolution(loss, bounds=bounds, args=args, **kwargs)


def finish(mse, X):
    """Scalar for a single trial vector, [S] array for a population."""
    return float(mse[0]) if np.ndim(X) == 1 else mse
