"""Drop-in ``GR4J`` (interface of ``rrmpg/models/gr4j.py``) running on the B200 engine."""
import numbers

import numpy as np

from .. import engine
from ..utils.array_checks import check_for_negatives, validate_array_input
from . import _fit
from .basemodel import BaseModel


def _validate_forcing(prec, etp):
    prec = validate_array_input(prec, np.float64, 'precipitation')
    etp = validate_array_input(etp, np.float64, 'pot. evapotranspiration')
    return prec, etp


def _check_forcing(prec, etp):
    if check_for_negatives(prec):
        raise ValueError("The precipitation array contains negative values.")
    if len(prec) != len(etp):
        raise RuntimeError("The arrays of precipitation and pot. evapotranspiration,"
                           " must be of the same size.")


def _check_fractions(s_init, r_init):
    if (s_init < 0) or (s_init > 1):
        raise ValueError("The initial value of the production storage must be in "
                         "the range [0,1].")
    if (r_init < 0) or (r_init > 1):
        raise ValueError("The initial value of the routing storage must be in the"
                         " range [0,1].")


class GR4J(BaseModel):
    """GR4J (Perrin et al. 2003): production + routing store and two unit hydrographs, 4 parameters."""

    _param_list = ['x1', 'x2', 'x3', 'x4']
    _default_bounds = {'x1': (100, 1200), 'x2': (-5, 3), 'x3': (20, 300), 'x4': (1.1, 2.9)}
    _dtype = np.dtype([('x1', np.float64), ('x2', np.float64), ('x3', np.float64), ('x4', np.float64)])

    def __init__(self, params=None):
        super().__init__(params=params)

    def simulate(self, prec, etp, s_init=0., r_init=0., return_storage=False, params=None):
        """Simulate discharge for one or many parameter sets.

        Same arguments, checks and exceptions as ``gr4j.py:76-182``; returns ``qsim [T, N]`` (and
        ``s_store, r_store``).  ``s_init`` / ``r_init`` are fractions of x1 / x3.

        Deviation from the reference, on purpose: with ``return_storage=False`` the reference returns
        from inside its member loop after the first parameter set (``gr4j.py:178``) and leaves the
        other columns 0.  Here every member is simulated, as its docstring (``gr4j.py:94-97``) says.
        """
        prec, etp = _validate_forcing(prec, etp)
        _check_forcing(prec, etp)
        if not isinstance(s_init, numbers.Number):
            raise TypeError("'s1_init' must be a Number.")
        if not isinstance(r_init, numbers.Number):
            raise TypeError("'r_init' must be a Number.")
        s_init, r_init = float(s_init), float(r_init)
        _check_fractions(s_init, r_init)
        params = self._resolve_params(params)
        res = engine.gr4j(prec, etp, s_init, r_init, params, return_storage=return_storage)
        if return_storage:
            return res['qsim'], res['s_store'], res['r_store']
        return res['qsim']

    def fit(self, qobs, prec, etp, s_init=0., r_init=0.):
        """Calibrate x1..x4 against ``qobs`` with differential evolution (``gr4j.py:185-249``)."""
        prec, etp = _validate_forcing(prec, etp)
        qobs = validate_array_input(qobs, np.float64, 'observed discharge')
        _check_forcing(prec, etp)
        s_init, r_init = float(s_init), float(r_init)
        _check_fractions(s_init, r_init)
        args = (qobs, prec, etp, s_init, r_init, self._dtype)
        return _fit.minimise(_loss, self._bounds(), args)


def _loss(X, *args):
    """MSE of one trial vector (k,) or of a whole trial population (k, S); args as in gr4j.py:252-277."""
    qobs, prec, etp, s_init, r_init = args[:5]
    res = engine.gr4j(prec, etp, s_init, r_init, _fit.as_population(X), qobs=qobs, want_qsim=False)
    return _fit.finish(res['mse'], X)
