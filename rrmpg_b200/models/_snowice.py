"""The snow(+hysteresis)(+ice)+GR4J drop-ins: one implementation of what
``rrmpg/models/cemaneigegr4jice.py``, ``cemaneigehystgr4j.py`` and ``cemaneigehystgr4jice.py`` each spell out
three to four times (simulate, fit, fit_Q_SCA).  Same argument names / order / defaults, same checks in the
same order, same exception types and messages, same return tuples.
"""
import numpy as np

from .. import engine
from ..utils.array_checks import validate_array_input
from ..utils.metrics import kge_columns
from . import _fit, _snow_inputs
from .basemodel import BaseModel


class _SnowIceModel(BaseModel):
    """Shared machinery; subclasses set ``_hyst`` / ``_ice`` and the public signatures."""

    _hyst = False
    _ice = False

    def _prepare(self, prec, mean_temp, min_temp, max_temp, etp, frac_ice, met_station_height, snow_pack_init,
                 thermal_state_init, sca_init, s_init, r_init, altitudes):
        prec, mean_temp, min_temp, max_temp, etp = _snow_inputs.validate_series(prec, mean_temp, min_temp, max_temp,
                                                                                etp)
        altitudes = _snow_inputs.validate_altitudes(altitudes, met_station_height)
        inits = [_snow_inputs.validate_number(snow_pack_init, 'snow_pack_init'),
                 _snow_inputs.validate_number(thermal_state_init, 'thermal_state_init')]
        if self._hyst:
            inits.append(_snow_inputs.validate_number(sca_init, 'sca_init'))
        inits.append(_snow_inputs.validate_number(s_init, 's_init'))
        inits.append(_snow_inputs.validate_number(r_init, 'r_init'))
        if self._ice:
            if isinstance(frac_ice, np.ndarray) and frac_ice.ndim != 1:
                raise ValueError("frac_ice must be a 1D array.")
            frac_ice = np.asarray(frac_ice)
        inits = [float(v) for v in inits]
        prec, mean_temp, frac, _ = _snow_inputs.to_layers(prec, mean_temp, min_temp, max_temp, met_station_height,
                                                          altitudes)
        return prec, mean_temp, etp, frac, frac_ice, inits

    def _simulate(self, prepared, return_storages, params):
        prec, mean_temp, etp, frac, frac_ice, inits = prepared
        params = self._resolve_params(params)
        res = engine.snowice_gr4j(self._hyst, self._ice, prec, mean_temp, etp, frac_ice, frac, inits, params,
                                  return_storages=bool(return_storages))
        if not return_storages:
            return res['qsim']
        out = [res['qsim'], res['G'], res['eTG'], res['s_store'], res['r_store']]
        if self._hyst:
            out.append(res['sca'])
        if self._ice:
            out.append(res['icemelt'])
        if self._hyst and self._ice:
            out.append(res['snowmelt'])
        if self._hyst:
            # rain = prec - prec * frac_solid per layer (cemaneigehyst_model.py:98-99), the same for every member
            rain = prec - prec * frac
            out.append(np.repeat(rain[:, :, None], params.size, axis=2))
        return tuple(out)

    def _fit(self, loss, obs, prepared, loss_metric, extra=()):
        prec, mean_temp, etp, frac, frac_ice, inits = prepared
        args = (obs, prec, mean_temp, frac, etp, frac_ice, tuple(inits), self._dtype, loss_metric,
                self._hyst, self._ice) + tuple(extra)
        return _fit.minimise(loss, self._bounds(), args)


def _ensemble(X, args, return_storages=False, qobs=None, objective="mse"):
    obs, prec, mean_temp, frac, etp, frac_ice, inits, _dtype, _metric, hyst, ice = args[:11]
    return engine.snowice_gr4j(hyst, ice, prec, mean_temp, etp, frac_ice, frac, inits, _fit.as_population(X),
                               return_storages=return_storages, qobs=qobs, want_qsim=qobs is None, objective=objective)


def _loss(X, *args):
    """Loss of one trial vector (k,) or of a whole trial population (k, S).

    Both metrics are accumulated inside the kernel (``rrb_opts.objective``): no [T, S] discharge array is
    materialised.  'kge' returns calc_kge itself, as the reference's ``_loss`` does
    (``cemaneigehystgr4j.py:600-605``; its fit_Q_SCA uses 1 - KGE instead).
    """
    obs, metric = args[0], args[8]
    if metric in ("mse", "kge"):
        return _fit.finish(_ensemble(X, args, qobs=obs, objective=metric)['mse'], X)
    raise ValueError("Invalid loss_metric. Choose 'mse' or 'kge'.")


def _loss_Q_SCA(X, *args):
    """75 % discharge + 5 % per elevation band on the snow-covered area (``cemaneigehystgr4j.py:608-691``)."""
    obs, metric, ndsi = args[0], args[8], args[11]
    if metric not in ("mse", "kge"):
        raise ValueError("Invalid loss_metric. Choose 'mse' or 'kge'.")
    # the discharge term comes out of the kernel's registers (no [T, S] qsim); the snow-covered-area terms need the
    # per-band series, which the same launch writes as storages
    res = engine.snowice_gr4j(args[9], args[10], args[1], args[2], args[4], args[5], args[3], args[6],
                              _fit.as_population(X), return_storages=True, qobs=obs, want_qsim=False, objective=metric)
    sca = res['sca']
    if sca.shape[1] < 5:
        raise IndexError("fit_Q_SCA needs five elevation bands")
    if metric == "mse":
        loss = 0.75 * res['mse']
        for b in range(5):
            loss = loss + 0.05 * np.mean((np.asarray(ndsi[b])[:, None] - sca[:, b, :] * 100) ** 2, axis=0)
    else:
        loss = 0.75 * (1 - res['mse'])
        for b in range(5):
            loss = loss + 0.05 * (1 - kge_columns(ndsi[b], sca[:, b, :] * 100))
    return _fit.finish(loss, X)


class CemaneigeGR4JIce(_SnowIceModel):
    """IceMelt + Cemaneige + GR4J (Nepal et al. 2017); 7 parameters (CTG, Kf, x1..x4, DDF)."""

    _ice = True
    _param_list = ['CTG', 'Kf', 'x1', 'x2', 'x3', 'x4', 'DDF']
    _default_bounds = {'CTG': (0, 1), 'Kf': (1, 15), 'x1': (100, 1200), 'x2': (-5, 3), 'x3': (20, 300),
                       'x4': (1.1, 2.9), 'DDF': (1, 30)}
    _dtype = np.dtype([(name, np.float64) for name in _param_list])

    def __init__(self, params=None):
        super().__init__(params=params)

    def simulate(self, prec, mean_temp, min_temp, max_temp, etp, frac_ice, met_station_height, snow_pack_init=0,
                 thermal_state_init=0, s_init=0, r_init=0, altitudes=[], return_storages=False, params=None):
        """``cemaneigegr4jice.py:95-288``.  Returns qsim or (qsim, G, eTG, s_store, r_store, ice_melt)."""
        prepared = self._prepare(prec, mean_temp, min_temp, max_temp, etp, frac_ice, met_station_height,
                                 snow_pack_init, thermal_state_init, 0, s_init, r_init, altitudes)
        return self._simulate(prepared, return_storages, params)

    def fit(self, obs, prec, mean_temp, min_temp, max_temp, etp, frac_ice, met_station_height, snow_pack_init=0,
            thermal_state_init=0, s_init=0, r_init=0, altitudes=[]):
        """``cemaneigegr4jice.py:290-417``: differential evolution on the MSE."""
        obs = validate_array_input(obs, np.float64, 'obs')
        prepared = self._prepare(prec, mean_temp, min_temp, max_temp, etp, frac_ice, met_station_height,
                                 snow_pack_init, thermal_state_init, 0, s_init, r_init, altitudes)
        return self._fit(_loss, obs, prepared, "mse")


class CemaneigeHystGR4J(_SnowIceModel):
    """Cemaneige with SWE-SCA hysteresis (Riboust et al. 2019) + GR4J; 8 parameters."""

    _hyst = True
    _param_list = ['CTG', 'Kf', 'Thacc', 'Rsp', 'x1', 'x2', 'x3', 'x4']
    _default_bounds = {'CTG': (0, 1), 'Kf': (0, 10), 'Thacc': (0, 1000), 'Rsp': (0, 1), 'x1': (10, 1200),
                       'x2': (-5, 3), 'x3': (20, 5000), 'x4': (1.1, 10)}
    _dtype = np.dtype([(name, np.float64) for name in _param_list])

    def __init__(self, params=None):
        super().__init__(params=params)

    def simulate(self, prec, mean_temp, min_temp, max_temp, etp, met_station_height, snow_pack_init=0,
                 thermal_state_init=0, sca_init=0, s_init=0, r_init=0, altitudes=[], return_storages=False,
                 params=None):
        """``cemaneigehystgr4j.py:95-290``.  Returns qsim or (qsim, G, eTG, s_store, r_store, sca, rain)."""
        prepared = self._prepare(prec, mean_temp, min_temp, max_temp, etp, None, met_station_height, snow_pack_init,
                                 thermal_state_init, sca_init, s_init, r_init, altitudes)
        return self._simulate(prepared, return_storages, params)

    def fit(self, obs, prec, mean_temp, min_temp, max_temp, etp, met_station_height, loss_metric="mse",
            snow_pack_init=0, thermal_state_init=0, sca_init=0, s_init=0, r_init=0, altitudes=[]):
        """``cemaneigehystgr4j.py:292-424``; ``loss_metric`` 'mse' or 'kge'."""
        obs = validate_array_input(obs, np.float64, 'obs')
        prepared = self._prepare(prec, mean_temp, min_temp, max_temp, etp, None, met_station_height, snow_pack_init,
                                 thermal_state_init, sca_init, s_init, r_init, altitudes)
        return self._fit(_loss, obs, prepared, loss_metric)

    def fit_Q_SCA(self, obs, prec, mean_temp, min_temp, max_temp, etp, NDSI1, NDSI2, NDSI3, NDSI4, NDSI5,
                  met_station_height, loss_metric="mse", snow_pack_init=0, thermal_state_init=0, sca_init=0, s_init=0,
                  r_init=0, altitudes=[]):
        """``cemaneigehystgr4j.py:427-570``: calibrate on discharge (75 %) and snow-covered area of five bands."""
        obs = validate_array_input(obs, np.float64, 'obs')
        prepared = self._prepare(prec, mean_temp, min_temp, max_temp, etp, None, met_station_height, snow_pack_init,
                                 thermal_state_init, sca_init, s_init, r_init, altitudes)
        ndsi = tuple(np.asarray(a, dtype=np.float64).flatten() for a in (NDSI1, NDSI2, NDSI3, NDSI4, NDSI5))
        return self._fit(_loss_Q_SCA, obs, prepared, loss_metric, extra=(ndsi,))


class CemaneigeHystGR4JIce(_SnowIceModel):
    """IceMelt + Cemaneige with hysteresis + GR4J; 9 parameters."""

    _hyst = True
    _ice = True
    _param_list = ['CTG', 'Kf', 'Thacc', 'Rsp', 'x1', 'x2', 'x3', 'x4', 'DDF']
    _default_bounds = {'CTG': (0, 1), 'Kf': (0, 10), 'Thacc': (0, 1000), 'Rsp': (0, 1), 'x1': (10, 1200),
                       'x2': (-5, 3), 'x3': (20, 5000), 'x4': (1.1, 10), 'DDF': (0, 30)}
    _dtype = np.dtype([(name, np.float64) for name in _param_list])

    def __init__(self, params=None):
        super().__init__(params=params)

    def simulate(self, prec, mean_temp, min_temp, max_temp, etp, frac_ice, met_station_height, snow_pack_init=0,
                 thermal_state_init=0, sca_init=0, s_init=0, r_init=0, altitudes=[], return_storages=False,
                 params=None):
        """``cemaneigehystgr4jice.py:102-306``.  Returns qsim or
        (qsim, G, eTG, s_store, r_store, sca, ice_melt, snowmelt, rain)."""
        prepared = self._prepare(prec, mean_temp, min_temp, max_temp, etp, frac_ice, met_station_height,
                                 snow_pack_init, thermal_state_init, sca_init, s_init, r_init, altitudes)
        return self._simulate(prepared, return_storages, params)

    def fit(self, obs, prec, mean_temp, min_temp, max_temp, etp, frac_ice, met_station_height, loss_metric="mse",
            snow_pack_init=0, thermal_state_init=0, sca_init=0, s_init=0, r_init=0, altitudes=[]):
        """``cemaneigehystgr4jice.py:308-445``; ``loss_metric`` 'mse' or 'kge'."""
        obs = validate_array_input(obs, np.float64, 'obs')
        prepared = self._prepare(prec, mean_temp, min_temp, max_temp, etp, frac_ice, met_station_height,
                                 snow_pack_init, thermal_state_init, sca_init, s_init, r_init, altitudes)
        return self._fit(_loss, obs, prepared, loss_metric)

    def fit_Q_SCA(self, obs, prec, mean_temp, min_temp, max_temp, etp, frac_ice, NDSI1, NDSI2, NDSI3, NDSI4, NDSI5,
                  met_station_height, loss_metric="mse", snow_pack_init=0, thermal_state_init=0, sca_init=0, s_init=0,
                  r_init=0, altitudes=[]):
        """``cemaneigehystgr4jice.py:447-593``: calibrate on discharge and snow-covered area of five bands."""
        obs = validate_array_input(obs, np.float64, 'obs')
        prepared = self._prepare(prec, mean_temp, min_temp, max_temp, etp, frac_ice, met_station_height,
                                 snow_pack_init, thermal_state_init, sca_init, s_init, r_init, altitudes)
        ndsi = tuple(np.asarray(a, dtype=np.float64).flatten() for a in (NDSI1, NDSI2, NDSI3, NDSI4, NDSI5))
        return self._fit(_loss_Q_SCA, obs, prepared, loss_metric, extra=(ndsi,))
