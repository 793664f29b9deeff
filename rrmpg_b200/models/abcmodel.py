"""Drop-in ``ABCModel`` (interface of ``rrmpg/models/abcmodel.py``) running on the B200 engine."""
import numbers

import numpy as np

from .. import engine
from ..utils.array_checks import check_for_negatives, validate_array_input
from . import _fit
from .basemodel import BaseModel


def _check_initial_state(initial_state):
    if not isinstance(initial_state, numbers.Number) or initial_state < 0:
        raise TypeError("The variable 'initial_state' must be a numercial scaler "
                        "greate than 0.")
    return float(initial_state)


class ABCModel(BaseModel):
    """The linear ABC model (Fiering 1967); parameters a, b, c with a + b <= 1."""

    _param_list = ['a', 'b', 'c']
    _default_bounds = {'a': (0, 1), 'b': (0, 1), 'c': (0, 1)}
    _dtype = np.dtype([('a', np.float64), ('b', np.float64), ('c', np.float64)])

    def __init__(self, params=None):
        super().__init__(params=params)

    def get_random_params(self, num=1):
        """Random sets honouring b <= 1 - a; same draw order as ``abcmodel.py:70-103``."""
        params = np.zeros(num, dtype=self._dtype)
        for name in ('a', 'c'):
            low, high = self._default_bounds[name]
            params[name][:] = np.random.uniform(low=low, high=high, size=num)
        for i in range(num):
            params['b'][i] = np.random.uniform(low=self._default_bounds['b'][0],
                                               high=(1 - params['a'][i]), size=1).item()
        return params

    def simulate(self, prec, initial_state=0, return_storage=False, params=None):
        """Simulate streamflow for one or many parameter sets.

        Same arguments, checks, exceptions and return layout as ``abcmodel.py:105-185``; the member
        loop (:174-181) is one ``rrb_abc_simulate`` call.  Returns ``qsim [T, N]`` (and
        ``storage [T, N]``).
        """
        prec = validate_array_input(prec, np.float64, 'precipitation')
        if check_for_negatives(prec):
            raise ValueError("In the precipitation array are negative values.")
        initial_state = _check_initial_state(initial_state)
        if not isinstance(return_storage, bool):
            raise TypeError("The return_storage arg must be a boolean.")
        params = self._resolve_params(params)
        res = engine.abc(prec, initial_state, params, return_storage=return_storage)
        if return_storage:
            return res['qsim'], res['storage']
        return res['qsim']

    def fit(self, qobs, prec, initial_state=0):
        """Calibrate a, b, c against ``qobs`` with differential evolution (``abcmodel.py:188-232``)."""
        qobs = validate_array_input(qobs, np.float64, 'qobs')
        prec = validate_array_input(prec, np.float64, 'precipitation')
        if check_for_negatives(prec):
            raise ValueError("In the precipitation array are negative values.")
        initial_state = _check_initial_state(initial_state)
        args = (prec, initial_state, qobs, self._dtype)
        return _fit.minimise(_loss, self._bounds(), args)


def _loss(X, *args):
    """MSE of one trial vector (k,) or of a whole trial population (k, S)."""
    prec, initial_state, qobs = args[0], args[1], args[2]
    res = engine.abc(prec, initial_state, _fit.as_population(X), qobs=qobs, want_qsim=False)
    return _fit.finish(res['mse'], X)
