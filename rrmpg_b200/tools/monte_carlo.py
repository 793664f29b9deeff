"""Monte-Carlo simulation over the sampled parameter matrix (``rrmpg/tools/monte_carlo.py:19-76``).

The reference samples ``num`` parameter sets, loops ``model.simulate`` over them one member at a time
(``:64``) and then loops ``calc_mse`` over the columns (``:70-71``).  Here the sampled record array goes to the
engine as ONE ensemble call, and with ``qobs`` that same call accumulates every member's mean squared error in the
kernel's registers (``rrb_opts.qobs / mse``, ``engine.fused``): no second pass over ``[T, num]`` on the host.
With the extension ``return_qsim=False`` the discharge array is not materialised at all -- nothing but ``mse[num]``
crosses the PCIe link (7.66 GB less at 65 536 members x 40 years).
"""
import numpy as np

from .. import engine
from ..models.basemodel import BaseModel
from ..utils.array_checks import validate_array_input


def monte_carlo(model, num, qobs=None, **kwargs):
    """Run ``num`` random parameter sets of ``model``.

    Returns ``{'params': record array [num], 'qsim': [T, num]}`` plus ``'mse': [num]`` when ``qobs`` is
    given -- the reference's return value.  ``kwargs`` are the model's ``simulate`` arguments, plus two extensions
    that are not passed on: ``return_qsim`` (default True; False drops ``'qsim'`` from the result and skips the
    transfer) and ``objective`` ('mse' (default, the reference's), 'nse' or 'kge': what ``'mse'`` holds).
    """
    if not issubclass(model.__class__, BaseModel):
        raise TypeError("The model must be one of the models implemented in the "
                        "rrmpg.models module.")
    if not isinstance(num, int) or num < 1:
        raise TypeError("'n' must be a positive integer greate than zero.")
    return_qsim = bool(kwargs.pop('return_qsim', True))
    objective = kwargs.pop('objective', 'mse')
    if qobs is not None:
        qobs = validate_array_input(qobs, np.float64, 'qobs')
    elif not return_qsim:
        raise ValueError("return_qsim=False needs qobs: there would be nothing to return")

    params = model.get_random_params(num=num)
    if qobs is None:
        qsim = model.simulate(params=params, **kwargs)
        if isinstance(qsim, tuple):  # return_storage(s)=True was passed through kwargs
            qsim = qsim[0]
        return {'params': params, 'qsim': qsim}

    # one launch: the ensemble and its per-member objective (a length mismatch raises the reference's
    # ValueError("Arrays must have the same size."), rrmpg/utils/metrics.py:127-128)
    with engine.fused(qobs, objective=objective, want_qsim=return_qsim) as f:
        qsim = model.simulate(params=params, **kwargs)
    if isinstance(qsim, tuple):
        qsim = qsim[0]
    out = {'params': params, 'mse': np.asarray(f.values)}
    if return_qsim:
        out['qsim'] = qsim
    return out
