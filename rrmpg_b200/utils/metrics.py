"""The skill scores the hot path's callers use (``rrmpg/utils/metrics.py``).

Only ``calc_mse`` (:110-136) is on the path (``_loss`` of every model and ``monte_carlo``); NSE and
RMSE (:28-107) are kept because the reference's tutorials score Monte-Carlo ensembles with them.
The remaining scores of the reference (KGE, alpha/beta-NSE, r) are out of scope.
"""
import numpy as np

from .array_checks import validate_array_input


def _pair(obs, sim):
    obs = validate_array_input(obs, np.float64, 'obs')
    sim = validate_array_input(sim, np.float64, 'sim')
    if len(obs) != len(sim):
        raise ValueError("Arrays must have the same size.")
    return obs, sim


def calc_mse(obs, sim):
    """Mean squared error, ``np.mean((obs - sim)**2)``."""
    obs, sim = _pair(obs, sim)
    return np.mean((obs - sim) ** 2)


def calc_rmse(obs, sim):
    """Root mean squared error."""
    obs, sim = _pair(obs, sim)
    return np.sqrt(np.mean((obs - sim) ** 2))


def calc_nse(obs, sim):
    """Nash-Sutcliffe efficiency; raises RuntimeError when obs is constant (denominator 0)."""
    obs, sim = _pair(obs, sim)
    denominator = np.sum((obs - np.mean(obs)) ** 2)
    if denominator == 0:
        msg = ["The Nash-Sutcliffe-Efficiency coefficient is not defined ",
               "for the case, that all values in the observations are equal.",
               " Maybe you should use the Mean-Squared-Error instead."]
        raise RuntimeError("".join(msg))
    numerator = np.sum((sim - obs) ** 2)
    return 1 - numerator / denominator
