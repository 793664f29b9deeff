"""Member-independent forcing preprocessing of the Cemaneige family (host side, run once per call).

Vectorised numpy restatements of ``rrmpg/models/cemaneige_utils.py``: ``calculate_solid_fraction``
(:16-98), ``extrapolate_precipitation`` (:101-158) and ``extrapolate_temperature`` (:161-208).
They turn the station series ``[T]`` into per-elevation-layer arrays ``[T, L]``; results are
bit-identical to the numba originals (same IEEE operations per element; the scalar ``exp`` goes
through libm via ``math.exp`` like numba's).
"""
import math

import numpy as np


def calculate_solid_fraction(prec, altitudes, mean_temp, min_temp, max_temp):
    """Fraction of solid precipitation per timestep and layer (airGR's USACE / Hydrotel split)."""
    z_thresh = 1500
    altitudes = np.asarray(altitudes)
    num_timesteps, num_layers = prec.shape[0], len(altitudes)
    solid_fraction = np.zeros((num_timesteps, num_layers), dtype=np.float64)
    for l in range(num_layers):
        if altitudes[l] < z_thresh:
            mx, mn = max_temp[:, l], min_temp[:, l]
            mixed = ~(mx <= 0) & ~(mn >= 0)
            col = np.where(mx <= 0, 1.0, 0.0)
            col[mixed] = 1 - (mx[mixed] / (mx[mixed] - mn[mixed]))
        else:
            me = mean_temp[:, l]
            mixed = ~(me >= 3) & ~(me <= 0)
            col = np.where(me >= 3, 0.0, np.where(me <= 0, 1.0, 0.0))
            col[mixed] = 1 - (me[mixed] + 1) / 4
        solid_fraction[:, l] = col
    return solid_fraction


def extrapolate_precipitation(prec, altitudes, met_station_height):
    """Station precipitation -> layer precipitation (exponential altitude gradient, capped at 4000 m)."""
    beta_altitude = 0.0004
    z_thresh = 4000
    altitudes = np.asarray(altitudes)
    layer_prec = np.zeros((prec.shape[0], len(altitudes)), dtype=np.float64)
    for l in range(len(altitudes)):
        if altitudes[l] <= z_thresh:
            layer_prec[:, l] = prec * math.exp(float((altitudes[l] - met_station_height) * beta_altitude))
        elif met_station_height <= z_thresh:
            layer_prec[:, l] = prec * math.exp(float((z_thresh - met_station_height) * beta_altitude))
        else:
            layer_prec[:, l] = prec
    return layer_prec


def extrapolate_temperature(min_temp, mean_temp, max_temp, altitudes, met_station_height):
    """Station temperatures -> layer temperatures (-0.65 K / 100 m lapse rate)."""
    theta_temp = -0.0065
    altitudes = np.asarray(altitudes)
    shape = (min_temp.shape[0], len(altitudes))
    layer_min, layer_mean, layer_max = (np.zeros(shape, dtype=np.float64) for _ in range(3))
    for l in range(len(altitudes)):
        delta_temp = (altitudes[l] - met_station_height) * theta_temp
        layer_min[:, l] = min_temp + delta_temp
        layer_mean[:, l] = mean_temp + delta_temp
        layer_max[:, l] = max_temp + delta_temp
    return layer_min, layer_mean, layer_max
