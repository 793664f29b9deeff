"""Drop-in ``CemaneigeGR4J`` (interface of ``rrmpg/models/cemaneigegr4j.py``) on the B200 engine."""
import numpy as np

from .. import engine
from ..utils.array_checks import validate_array_input
from . import _fit, _snow_inputs
from .basemodel import BaseModel


class CemaneigeGR4J(BaseModel):
    """Cemaneige snow routine feeding GR4J; 6 parameters (CTG, Kf, x1..x4)."""

    _param_list = ['CTG', 'Kf', 'x1', 'x2', 'x3', 'x4']
    _default_bounds = {'CTG': (0, 1), 'Kf': (0, 10), 'x1': (100, 1200), 'x2': (-5, 3), 'x3': (20, 300),
                       'x4': (1.1, 2.9)}
    _dtype = np.dtype([('CTG', np.float64), ('Kf', np.float64), ('x1', np.float64), ('x2', np.float64),
                       ('x3', np.float64), ('x4', np.float64)])

    def __init__(self, params=None):
        super().__init__(params=params)

    def _prepare(self, prec, mean_temp, min_temp, max_temp, etp, met_station_height, snow_pack_init,
                 thermal_state_init, s_init, r_init, altitudes):
        prec, mean_temp, min_temp, max_temp, etp = _snow_inputs.validate_series(prec, mean_temp, min_temp,
                                                                                max_temp, etp)
        altitudes = _snow_inputs.validate_altitudes(altitudes, met_station_height)
        inits = [_snow_inputs.validate_number(snow_pack_init, 'snow_pack_init'),
                 _snow_inputs.validate_number(thermal_state_init, 'thermal_state_init'),
                 _snow_inputs.validate_number(s_init, 's1_init'),
                 _snow_inputs.validate_number(r_init, 'r_init')]
        inits = [float(v) for v in inits]
        prec, mean_temp, frac, _ = _snow_inputs.to_layers(prec, mean_temp, min_temp, max_temp,
                                                          met_station_height, altitudes)
        return prec, mean_temp, etp, frac, inits

    def simulate(self, prec, mean_temp, min_temp, max_temp, etp, met_station_height, snow_pack_init=0,
                 thermal_state_init=0, s_init=0, r_init=0, altitudes=[], return_storages=False, params=None):
        """Simulate discharge of the coupled model for one or many parameter sets.

        Same arguments, checks and exceptions as ``cemaneigegr4j.py:88-273``.  Returns ``qsim [T, N]``
        and, with ``return_storages=True``, ``G, eTG`` (``[T, L, N]``) and ``s_store, r_store``.
        """
        prec, mean_temp, etp, frac, inits = self._prepare(prec, mean_temp, min_temp, max_temp, etp,
                                                          met_station_height, snow_pack_init,
                                                          thermal_state_init, s_init, r_init, altitudes)
        params = self._resolve_params(params)
        res = engine.cemaneigegr4j(prec, mean_temp, etp, frac, inits, params,
                                   return_storages=bool(return_storages))
        if return_storages:
            return res['qsim'], res['G'], res['eTG'], res['s_store'], res['r_store']
        return res['qsim']

    def fit(self, obs, prec, mean_temp, min_temp, max_temp, etp, met_station_height, snow_pack_init=0,
            thermal_state_init=0, s_init=0, r_init=0, altitudes=[]):
        """Calibrate the 6 parameters against an observed discharge series (``cemaneigegr4j.py:275-400``)."""
        obs = validate_array_input(obs, np.float64, 'obs')
        prec, mean_temp, etp, frac, inits = self._prepare(prec, mean_temp, min_temp, max_temp, etp,
                                                          met_station_height, snow_pack_init,
                                                          thermal_state_init, s_init, r_init, altitudes)
        args = (obs, prec, mean_temp, frac, etp, inits[0], inits[1], inits[2], inits[3], self._dtype)
        return _fit.minimise(_loss, self._bounds(), args)


def _loss(X, *args):
    """MSE of one trial vector or a whole trial population; args as in cemaneigegr4j.py:403-436."""
    obs, prec, mean_temp, frac, etp = args[:5]
    res = engine.cemaneigegr4j(prec, mean_temp, etp, frac, args[5:9], _fit.as_population(X), qobs=obs,
                               want_qsim=False)
    return _fit.finish(res['mse'], X)
