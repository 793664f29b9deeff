"""The skill scores the hot path's callers use (``rrmpg/utils/metrics.py``).

``calc_mse`` (:110-136) is on the path (``_loss`` of every model and ``monte_carlo``); ``calc_kge``
(:139-188) is the alternative loss of the Hyst models' ``fit`` (``cemaneigehystgr4j.py:600-605``); NSE and
RMSE (:28-107) are kept because the reference's tutorials score Monte-Carlo ensembles with them.
alpha/beta-NSE and r are out of scope.
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


def calc_kge(obs, sim):
    """Kling-Gupta efficiency (Gupta et al. 2009): 1 - sqrt((r-1)^2 + (alpha-1)^2 + (beta-1)^2)."""
    obs, sim = _pair(obs, sim)
    mean_obs = np.mean(obs)
    if mean_obs == 0:
        raise RuntimeError("KGE not definied if the mean of the observations equals 0.")
    std_obs = np.std(obs)
    if std_obs == 0:
        raise RuntimeError("KGE not definied if the standard deviation of the "
                           "observations equals 0.")
    r = np.corrcoef(obs, sim)[0, 1]  # Pearson r (scipy.stats.pearsonr in the reference)
    alpha = np.std(sim) / std_obs
    beta = np.mean(sim) / mean_obs
    return 1 - np.sqrt((r - 1) ** 2 + (alpha - 1) ** 2 + (beta - 1) ** 2)


def kge_columns(obs, sim):
    """calc_kge of every column of ``sim [T, S]`` against ``obs [T]`` in one vectorised pass."""
    obs = np.asarray(obs, dtype=np.float64)
    mean_obs, std_obs = np.mean(obs), np.std(obs)
    if mean_obs == 0:
        raise RuntimeError("KGE not definied if the mean of the observations equals 0.")
    if std_obs == 0:
        raise RuntimeError("KGE not definied if the standard deviation of the "
                           "observations equals 0.")
    ms, ss = np.mean(sim, axis=0), np.std(sim, axis=0)
    with np.errstate(invalid="ignore", divide="ignore"):
        r = np.mean((obs[:, None] - mean_obs) * (sim - ms), axis=0) / (std_obs * ss)
    return 1 - np.sqrt((r - 1) ** 2 + (ss / std_obs - 1) ** 2 + (ms / mean_obs - 1) ** 2)
