"""Drop-in ``HBVEdu`` (interface of ``rrmpg/models/hbvedu.py``) running on the B200 engine."""
import numpy as np

from .. import engine
from ..utils.array_checks import check_for_negatives, validate_array_input
from . import _fit
from .basemodel import BaseModel


def _validate_forcing(temp, prec, month, PE_m, T_m):
    """Checks of ``hbvedu.py:133-164``; returns validated copies with a 0-based month."""
    temp = validate_array_input(temp, np.float64, 'temperature')
    prec = validate_array_input(prec, np.float64, 'precipitation')
    if check_for_negatives(prec):
        raise ValueError("In the precipitation array are negative values.")
    month = validate_array_input(month, np.int8, 'month')
    if any(len(arr) != len(temp) for arr in [prec, month]):
        raise RuntimeError("The arrays of the temperature, precipitation and month "
                           "data must be of equal size.")
    PE_m = validate_array_input(PE_m, np.float64, 'PE_m')
    T_m = validate_array_input(T_m, np.float64, 'T_m')
    if any(len(arr) != 12 for arr in [PE_m, T_m]):
        raise RuntimeError("The monthly potential evapotranspiration and temperature"
                           " array must be of length 12.")
    if (np.min(month) < 1) or (np.max(month) > 12):
        raise ValueError("The month array must be between an integer1 (Jan) and " "12 (Dec).")
    month -= 1  # the kernel indexes PE_m / T_m from 0 (hbvedu.py:164); month is our private copy
    return temp, prec, month, PE_m, T_m


class HBVEdu(BaseModel):
    """The educational HBV model (Aghakouchak & Habib 2010), daily timestep, 11 parameters."""

    _param_list = ['T_t', 'DD', 'FC', 'Beta', 'C', 'PWP', 'K_0', 'K_1', 'K_2', 'K_p', 'L']
    _default_bounds = {'T_t': (-1, 1), 'DD': (3, 7), 'FC': (100, 200), 'Beta': (1, 7),
                       'C': (0.01, 0.07), 'PWP': (90, 180), 'K_0': (0.05, 0.2), 'K_1': (0.01, 0.1),
                       'K_2': (0.01, 0.05), 'K_p': (0.01, 0.05), 'L': (2, 5)}
    _dtype = np.dtype([(name, np.float64) for name in _param_list])

    def __init__(self, params=None):
        super().__init__(params=params)

    def simulate(self, temp, prec, month, PE_m, T_m, snow_init=0, soil_init=0, s1_init=0, s2_init=0,
                 return_storage=False, params=None):
        """Simulate discharge for one or many parameter sets.

        Same arguments, checks, exceptions and return order (``qsim, snow, soil, s1, s2``, each
        ``[T, N]``) as ``hbvedu.py:82-214``; the member loop (:199-209) is one
        ``rrb_hbvedu_simulate`` call.  ``month`` holds 1..12.
        """
        temp, prec, month, PE_m, T_m = _validate_forcing(temp, prec, month, PE_m, T_m)
        inits = [float(snow_init), float(soil_init), float(s1_init), float(s2_init)]
        params = self._resolve_params(params)
        res = engine.hbvedu(temp, prec, month, PE_m, T_m, inits, params, return_storage=return_storage)
        if return_storage:
            return res['qsim'], res['snow'], res['soil'], res['s1'], res['s2']
        return res['qsim']

    def fit(self, qobs, temp, prec, month, PE_m, T_m, snow_init=0., soil_init=0., s1_init=0., s2_init=0.):
        """Calibrate the 11 parameters against ``qobs`` (``hbvedu.py:216-307``)."""
        qobs = validate_array_input(qobs, np.float64, 'qobs')
        temp, prec, month, PE_m, T_m = _validate_forcing(temp, prec, month, PE_m, T_m)
        inits = [float(snow_init), float(soil_init), float(s1_init), float(s2_init)]
        args = (qobs, temp, prec, month, PE_m, T_m, inits[0], inits[1], inits[2], inits[3], self._dtype)
        return _fit.minimise(_loss, self._bounds(), args)


def _loss(X, *args):
    """MSE of one trial vector (k,) or of a whole trial population (k, S); args as in hbvedu.py:310-346."""
    qobs, temp, prec, month, PE_m, T_m = args[:6]
    res = engine.hbvedu(temp, prec, month, PE_m, T_m, args[6:10], _fit.as_population(X), qobs=qobs,
                        want_qsim=False)
    return _fit.finish(res['mse'], X)
