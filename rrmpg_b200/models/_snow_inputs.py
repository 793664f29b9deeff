"""Argument checks and layer preprocessing shared by Cemaneige and CemaneigeGR4J.

One implementation of what ``rrmpg/models/cemaneige.py:134-219`` and
``rrmpg/models/cemaneigegr4j.py:146-219`` each spell out twice (simulate and fit): same order of
checks, same exception types and messages.
"""
import numbers

import numpy as np

from ..utils.array_checks import check_for_negatives, validate_array_input
from .cemaneige_utils import (calculate_solid_fraction, extrapolate_precipitation,
                              extrapolate_temperature)


def validate_series(prec, mean_temp, min_temp, max_temp, etp=None):
    prec = validate_array_input(prec, np.float64, 'prec')
    mean_temp = validate_array_input(mean_temp, np.float64, 'mean_temp')
    min_temp = validate_array_input(min_temp, np.float64, 'min_temp')
    max_temp = validate_array_input(max_temp, np.float64, 'max_temp')
    others = [mean_temp, min_temp, max_temp]
    if etp is not None:
        etp = validate_array_input(etp, np.float64, 'pot. evapotranspiration')
        others.append(etp)
    if check_for_negatives(prec):
        raise ValueError("The precipitation array contains negative values.")
    if any(len(ar) != len(prec) for ar in others):
        raise RuntimeError("All meteorological input arrays must have the same length.")
    return prec, mean_temp, min_temp, max_temp, etp


def validate_altitudes(altitudes, met_station_height):
    if not isinstance(altitudes, list):
        raise TypeError("'altitudes' must be a list.")
    if len(altitudes) > 0:
        for val in altitudes:
            if not isinstance(val, numbers.Number):
                raise TypeError("All elements in 'altitudes must be numbers.")
        if met_station_height is None:
            raise ValueError(["The height of the meteorological station is missing."])
        if not isinstance(met_station_height, numbers.Number):
            raise TypeError("'met_station_height' must be a number.")
        altitudes = np.array(altitudes)
    if not isinstance(met_station_height, numbers.Number):
        raise TypeError("'met_station_height' must be a Number.")
    return altitudes


def validate_number(value, label):
    if not isinstance(value, numbers.Number):
        raise TypeError("'{}' must be a Number.".format(label))
    return value


def to_layers(prec, mean_temp, min_temp, max_temp, met_station_height, altitudes):
    """[T] station series -> ([T,L] prec, [T,L] mean_temp, [T,L] frac_solid, L)."""
    if len(altitudes) > 0:
        prec = extrapolate_precipitation(prec, altitudes, met_station_height)
        min_temp, mean_temp, max_temp = extrapolate_temperature(min_temp, mean_temp, max_temp, altitudes,
                                                                met_station_height)
    else:
        prec = np.expand_dims(prec, axis=-1)
        mean_temp = np.expand_dims(mean_temp, axis=-1)
        min_temp = np.expand_dims(min_temp, axis=-1)
        max_temp = np.expand_dims(max_temp, axis=-1)
        altitudes = np.array([met_station_height])
    frac_solid_prec = calculate_solid_fraction(prec, altitudes, mean_temp, min_temp, max_temp)
    return prec, mean_temp, frac_solid_prec, len(altitudes)
