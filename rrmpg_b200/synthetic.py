"""Seeded synthetic forcing for parity tests and bench.py.

The reference ships no benchmark inputs, so the workload is a deterministic synthetic
catchment: gamma wet-day precipitation, sinusoidal temperature + noise, sinusoidal PET and
calendar months, with the HBV monthly tables and Cemaneige elevation bands of the
reference's own fixtures (``test/test_models.py:227-268``).
"""
import numpy as np

SEED = 20260101
PARAM_SEED = 12345

PE_M = np.array([0.2, 0.3, 0.8, 1.6, 2.6, 3.4, 3.8, 3.3, 2.2, 1.2, 0.5, 0.2])
T_M = np.array([-3.0, -2.0, 2.0, 7.0, 12.0, 16.0, 18.0, 17.0, 13.0, 8.0, 3.0, -1.0])
MET_STATION_HEIGHT = 495
ALTITUDES = [550, 620, 700, 785, 920]

HBV_INITS = dict(snow_init=0, soil_init=100, s1_init=3, s2_init=10)
GR4J_INITS = dict(s_init=0.6, r_init=0.7)

T_DAILY_40Y = 14610
T_DAILY_10Y = 3652
T_HOURLY_10Y = 87660


def forcing(T=T_DAILY_40Y, seed=SEED, hourly=False):
    """Return a dict of float64 [T] series (+ 1-based int month) for one catchment."""
    rng = np.random.default_rng(seed)
    steps_per_day = 24 if hourly else 1
    day = np.arange(T) / steps_per_day
    doy = day % 365.25
    season = np.sin(2 * np.pi * (doy - 110) / 365.25)
    if hourly:
        wet = rng.random(T) < 0.10
        prec = np.where(wet, rng.gamma(0.8, 1.0, T), 0.0)
    else:
        wet = rng.random(T) < 0.45
        prec = np.where(wet, rng.gamma(0.8, 6.0, T), 0.0)
    temp = 8 + 12 * season + rng.normal(0, 3, T)
    if hourly:
        temp = temp + 4 * np.sin(2 * np.pi * (np.arange(T) % 24 - 9) / 24)
    etp = np.maximum(0.0, 2.0 + 1.8 * season)
    if hourly:
        etp = etp / 24
    dates = np.datetime64("1980-01-01") + np.floor(day).astype("timedelta64[D]")
    month = (dates.astype("datetime64[M]").astype(np.int64) % 12 + 1).astype(np.int8)
    return dict(prec=prec, temp=temp, etp=etp, month=month,
                min_temp=temp - 4, max_temp=temp + 4, PE_m=PE_M.copy(), T_m=T_M.copy())


def random_params(model, num, seed=PARAM_SEED):
    """``np.random.seed(seed); model.get_random_params(num)`` (rrmpg/models/basemodel.py:68-91)."""
    state = np.random.get_state()
    try:
        np.random.seed(seed)
        return model.get_random_params(num)
    finally:
        np.random.set_state(state)


def qobs_like(qsim_column, seed=SEED + 7):
    """A noisy 'observed' series derived from one simulated column (for MSE paths)."""
    rng = np.random.default_rng(seed)
    return np.maximum(0.0, qsim_column * (1 + 0.1 * rng.normal(size=qsim_column.shape)))
