"""Drop-in ``Cemaneige`` (interface of ``rrmpg/models/cemaneige.py``) running on the B200 engine."""
import numpy as np

from .. import engine
from ..utils.array_checks import validate_array_input
from . import _fit, _snow_inputs
from .basemodel import BaseModel


class Cemaneige(BaseModel):
    """Cemaneige snow accounting routine (Valery 2010): snow pack + thermal state per elevation layer."""

    _param_list = ['CTG', 'Kf']
    _default_bounds = {'CTG': (0, 1), 'Kf': (0, 10)}
    _dtype = np.dtype([('CTG', np.float64), ('Kf', np.float64)])

    def __init__(self, params=None):
        super().__init__(params=params)

    def _prepare(self, prec, mean_temp, min_temp, max_temp, met_station_height, snow_pack_init,
                 thermal_state_init, altitudes):
        prec, mean_temp, min_temp, max_temp, _ = _snow_inputs.validate_series(prec, mean_temp, min_temp, max_temp)
        altitudes = _snow_inputs.validate_altitudes(altitudes, met_station_height)
        snow_pack_init = float(_snow_inputs.validate_number(snow_pack_init, 'snow_pack_init'))
        thermal_state_init = float(_snow_inputs.validate_number(thermal_state_init, 'thermal_state_init'))
        prec, mean_temp, frac, _ = _snow_inputs.to_layers(prec, mean_temp, min_temp, max_temp,
                                                          met_station_height, altitudes)
        return prec, mean_temp, frac, snow_pack_init, thermal_state_init

    def simulate(self, prec, mean_temp, min_temp, max_temp, met_station_height, snow_pack_init=0,
                 thermal_state_init=0, altitudes=[], return_storages=False, params=None):
        """Simulate the liquid outflow of the snow routine for one or many parameter sets.

        Same arguments, checks and exceptions as ``cemaneige.py:81-245``.  Returns ``outflow [T, N]``
        and, with ``return_storages=True``, ``G`` and ``eTG`` as ``[T, L, N]``.
        """
        prec, mean_temp, frac, g0, e0 = self._prepare(prec, mean_temp, min_temp, max_temp, met_station_height,
                                                      snow_pack_init, thermal_state_init, altitudes)
        params = self._resolve_params(params)
        res = engine.cemaneige(prec, mean_temp, frac, g0, e0, params, return_storages=bool(return_storages))
        if return_storages:
            return res['outflow'], res['G'], res['eTG']
        return res['outflow']

    def fit(self, obs, prec, mean_temp, min_temp, max_temp, met_station_height, snow_pack_init=0,
            thermal_state_init=0, altitudes=[]):
        """Calibrate CTG, Kf against an observed outflow series (``cemaneige.py:247-359``)."""
        obs = validate_array_input(obs, np.float64, 'obs')
        prec, mean_temp, frac, g0, e0 = self._prepare(prec, mean_temp, min_temp, max_temp, met_station_height,
                                                      snow_pack_init, thermal_state_init, altitudes)
        args = (obs, prec, mean_temp, frac, g0, e0, self._dtype)
        return _fit.minimise(_loss, self._bounds(), args)


def _loss(X, *args):
    """MSE of one trial vector or a whole trial population; args as in cemaneige.py:362-387."""
    obs, prec, mean_temp, frac, g0, e0 = args[:6]
    res = engine.cemaneige(prec, mean_temp, frac, g0, e0, _fit.as_population(X), qobs=obs, want_outflow=False)
    return _fit.finish(res['mse'], X)
