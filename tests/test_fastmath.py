"""FAST-path transcendentals evaluated on the host (same header, same tables as the kernels) vs glibc."""
import ctypes as C

import numpy as np

from rrmpg_b200 import _lib


def _pow(x, y):
    x = np.ascontiguousarray(x, np.float64); y = np.ascontiguousarray(y, np.float64); out = np.empty_like(x)
    _lib.lib().rrb_host_fast_pow(x.ctypes.data, y.ctypes.data, x.size, out.ctypes.data)
    return out


def _exp2m1(z):
    z = np.ascontiguousarray(z, np.float64); out = np.empty_like(z)
    _lib.lib().rrb_host_fast_exp2m1(z.ctypes.data, z.size, out.ctypes.data)
    return out


def test_fast_pow_accuracy_on_the_hbv_operand_range():
    rng = np.random.default_rng(0)
    x = rng.uniform(0.2, 1.6, 400_000); y = rng.uniform(1, 7, 400_000)  # soil/FC, Beta (hbvedu.py:47-60)
    rel = np.abs(_pow(x, y) - np.power(x, y)) / np.power(x, y)
    assert rel.max() < 4e-15 and rel.mean() < 2.5e-16


def test_fast_pow_accuracy_wide_range():
    rng = np.random.default_rng(1)
    x = 10 ** rng.uniform(-6, 6, 400_000); y = rng.uniform(-9, 9, 400_000)
    ref = np.power(x, y)
    rel = np.abs(_pow(x, y) - ref) / ref
    assert rel.max() < 5e-14


def test_fast_pow_special_values_follow_libm():
    x = np.array([0.0, -1.0, -2.0, 1.0, np.inf, np.nan, 1e-310, 2.0, 1e300, 1e-300, 0.5, 1.0, 0.0, 3.0])
    y = np.array([2.0, 2.5, 2.0, 5.0, 2.0, 1.0, 2.0, 2000.0, 5.0, 5.0, -3000.0, np.nan, -1.0, 0.0])
    with np.errstate(all="ignore"):
        ref = np.power(x, y)
    got = _pow(x, y)
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    assert np.array_equal(got[~np.isnan(ref)], ref[~np.isnan(ref)])


def test_exp2m1_is_relatively_accurate_down_to_zero():
    rng = np.random.default_rng(2)
    z = np.concatenate([rng.uniform(0, 4, 200_000), 10 ** rng.uniform(-14, 0, 200_000), [0.0]])
    import math
    ref = np.array([math.expm1(v * math.log(2)) for v in z[:2000]])  # libm
    got = _exp2m1(z)
    assert got[-1] == 0.0
    rel = np.abs(got[:2000] - ref) / np.maximum(ref, 1e-300)
    assert rel.max() < 1e-15
    ref_all = np.expm1(z * np.log(2))
    ok = ref_all > 0
    assert (np.abs(got[ok] - ref_all[ok]) / ref_all[ok]).max() < 2e-15
    # saturation used by the tanh formula: clamp at z = 64
    assert _exp2m1(np.array([64.0, 500.0, np.inf]))[1] == _exp2m1(np.array([64.0]))[0]
    assert np.isnan(_exp2m1(np.array([np.nan]))[0])
