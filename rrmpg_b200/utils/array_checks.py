"""Input validation with the reference's exception types and messages.

Mirrors ``rrmpg/utils/array_checks.py`` (``check_for_negatives`` :15-32, ``validate_array_input``
:35-73): the drop-in models raise exactly what the reference raises, before any kernel runs.
"""
import numpy as np


def _is_series(arr):
    t = type(arr)
    return t.__name__ == "Series" and t.__module__.split(".")[0] == "pandas"


def check_for_negatives(arr):
    """True if any element is < 0 (NaN compares false, as in the reference's scan)."""
    return bool(np.any(np.asarray(arr) < 0))


def validate_array_input(arr, dtype, arr_name):
    """list / ndarray / pandas.Series -> fresh flat ndarray of ``dtype`` (always a copy)."""
    if not (isinstance(arr, (list, np.ndarray)) or _is_series(arr)):
        raise TypeError("The array {} must be either a list, ".format(arr_name)
                        + "numpy.ndarray or pandas.Series")
    try:
        return np.array(arr, dtype=dtype).flatten()
    except Exception:
        raise ValueError("The data in the parameter array '{}'".format(arr_name)
                         + " must be purely numerical.")
