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
    """differential_evolution with a whole generation per launch, then scipy's own polishing step (L-BFGS-B from the
    best member, ``differential_evolution(polish=True)``, the default the reference runs with) -- with the k + 1
    evaluations of every forward-difference gradient batched into ONE launch instead of k + 1 single-member launches."""
    kwargs = dict(vectorized=True, updating='deferred')
    if de_kwargs:
        kwargs.update(de_kwargs)
    polish = kwargs.pop('polish', True)
    res = optimize.differential_evolution(loss, bounds=bounds, args=args, polish=False, **kwargs)
    if polish:
        res = _polish(loss, res, np.asarray(bounds, dtype=np.float64), args)
    return res


_REL_STEP = float(np.sqrt(np.finfo(np.float64).eps))  # scipy's '2-point' relative step


def _polish(loss, res, limits, args):
    lo, hi = limits[:, 0], limits[:, 1]
    k = lo.size
    calls = [0]

    def fun_and_grad(x):
        h = _REL_STEP * np.maximum(1.0, np.abs(x))
        h = np.where(x + h > hi, -h, h)          # one-sided step that stays inside the bounds
        X = np.repeat(np.asarray(x, dtype=np.float64)[:, None], k + 1, axis=1)
        X[np.arange(k), np.arange(1, k + 1)] += h
        f = np.asarray(loss(X, *args), dtype=np.float64)
        calls[0] += k + 1
        return float(f[0]), (f[1:] - f[0]) / h

    r = optimize.minimize(fun_and_grad, np.copy(res.x), method='L-BFGS-B', jac=True, bounds=limits)
    res.nfev += calls[0]
    # acceptance rule of scipy's DifferentialEvolutionSolver.solve()
    if r.fun < res.fun and r.success and np.all(r.x <= hi) and np.all(lo <= r.x):
        res.fun, res.x, res.jac = r.fun, r.x, r.jac
    return res


def finish(mse, X):
    """Scalar for a single trial vector, [S] array for a population."""
    return float(mse[0]) if np.ndim(X) == 1 else mse
