"""Parent class of the drop-in models (API contract of ``rrmpg/models/basemodel.py:20-175``).

Sub-classes declare ``_param_list``, ``_default_bounds`` and the structured ``_dtype`` (the
parameter record ABI shared with the kernels: field order = column order of the [N, k] matrix
handed to ``librrmpg_b200``).  Sampling uses the global ``numpy.random`` state exactly like the
reference, so ``np.random.seed(s); model.get_random_params(n)`` yields identical ensembles.
"""
import numbers

import numpy as np


class BaseModel(object):
    """Parameter handling shared by all rainfall-runoff models."""

    _param_list = []
    _default_bounds = {}
    _dtype = np.dtype([])

    def __init__(self, params=None):
        if params:
            missing = [name for name in self._param_list if name not in params.keys()]
            if len(missing) > 0:
                raise AttributeError("Missing the following model parameters: "
                                     "{}".format(missing))
        else:
            params = self.get_random_params()
        self.set_params(params)

    def get_random_params(self, num=1):
        """``num`` parameter sets drawn uniformly inside ``_default_bounds`` (basemodel.py:68-91)."""
        params = np.zeros(num, dtype=self._dtype)
        for name in self._param_list:
            low, high = self._default_bounds[name]
            params[name] = np.random.uniform(low=low, high=high, size=num)
        return params

    def get_params(self):
        """Dict of the model's current parameter values."""
        return {name: getattr(self, name) for name in self._param_list}

    def set_params(self, params):
        """Set parameters from a dict, a record (``np.void``) or a record array of ``_dtype``."""
        if isinstance(params, dict):
            for name, value in params.items():
                if name not in self._param_list:
                    raise AttributeError("Unknow parameter '{}'.".format(name)
                                         + "Name must match one of the model parameters."
                                         + "Use {}".format(self.__class__.__name__)
                                         + ".get_parameter_names() to get a list of valid names.")
                if not isinstance(value, numbers.Number):
                    raise ValueError("The value of parameter '{}'".format(name) + "must be numerical")
                setattr(self, name, value)
        elif isinstance(params, (np.void, np.ndarray)):
            if params.dtype != self._dtype:
                raise TypeError("The parameter array has the wrong data type. "
                                "It must be the custom data type of the model.")
            for name in self._param_list:
                setattr(self, name, params[name] if isinstance(params, np.void) else params[name][0])
        else:
            raise TypeError("Wrong input data type. Must be either a dict or a numpy.ndarray")

    def get_parameter_names(self):
        return self._param_list

    def get_default_bounds(self):
        return self._default_bounds

    def get_dtype(self):
        return self._dtype

    # -- helpers shared by the sub-classes' simulate() / fit() ---------------------------------
    def _resolve_params(self, params):
        """``params`` argument of simulate() -> 1-D record array (hbvedu.py:172-188)."""
        if params is None:
            params = np.zeros(1, dtype=self._dtype)
            for name in self._param_list:
                params[name] = getattr(self, name)
            return params
        if params.dtype != self._dtype:
            raise TypeError("The model parameters must be a numpy array of the "
                            "models own custom data type.")
        if isinstance(params, np.void):
            params = np.expand_dims(params, params.ndim)
        return params

    def _bounds(self):
        return tuple(self._default_bounds[name] for name in self._param_list)
