"""Monte-Carlo simulation over the sampled parameter matrix (``rrmpg/tools/monte_carlo.py:19-76``).

The reference samples ``num`` parameter sets, loops ``model.simulate`` over them one member at a time
and then loops ``calc_mse`` over the columns.  Here the sampled record array goes to the engine as one
ensemble, and with ``qobs`` the per-member MSE is accumulated inside the same kernel.
"""
import inspect

import numpy as np

from ..models.basemodel import BaseModel
from ..utils.array_checks import validate_array_input
from ..utils.metrics import calc_mse


def monte_carlo(model, num, qobs=None, **kwargs):
    """Run ``num`` random parameter sets of ``model``.

    Returns ``{'params': record array [num], 'qsim': [T, num]}`` plus ``'mse': [num]`` when ``qobs`` is
    given -- the reference's return value.  ``kwargs`` are the model's ``simulate`` arguments.
    """
    if not issubclass(model.__class__, BaseModel):
        raise TypeError("The model must be one of the models implemented in the "
                        "rrmpg.models module.")
    if not isinstance(num, int) or num < 1:
        raise TypeError("'n' must be a positive integer greate than zero.")
    if qobs is not None:
        qobs = validate_array_input(qobs, np.float64, 'qobs')

    params = model.get_random_params(num=num)
    qsim = model.simulate(params=params, **kwargs)
    if isinstance(qsim, tuple):  # return_storage(s)=True was passed through kwargs
        qsim = qsim[0]

    if qobs is None:
        return {'params': params, 'qsim': qsim}

    if len(qobs) != qsim.shape[0]:
        raise ValueError("Arrays must have the same size.")
    # column-wise mean((qobs - qsim)**2) in one vectorised pass (pairwise summation like np.mean)
    mse_values = np.mean((qobs[:, None] - qsim) ** 2, axis=0) if num * qsim.shape[0] <= (1 << 24) \
        else _mse_blocked(qobs, qsim)
    return {'params': params, 'qsim': qsim, 'mse': mse_values}


def _mse_blocked(qobs, qsim, block=4096):
    out = np.empty(qsim.shape[1], dtype=np.float64)
    for lo in range(0, qsim.shape[1], block):
        d = qobs[:, None] - qsim[:, lo:lo + block]
        out[lo:lo + block] = np.mean(d * d, axis=0)
    return out
