import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """`gpu` tests need a CUDA device: skip them (with the reason) on a box that has none, so that a plain `pytest tests`
    on a CPU box runs the CPU suite instead of stopping at the first GPU test.  On a GPU box nothing is skipped: a
    library that is missing or does not load makes the tests fail loudly."""
    if os.path.exists("/dev/nvidia0") or os.path.exists("/dev/nvidiactl"):
        return
    try:
        from rrmpg_b200 import _lib
        if _lib.device_count() > 0:
            return
    except Exception:
        pass
    skip = pytest.mark.skip(reason="no CUDA device on this box")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name + ".npz")) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def golden():
    return load_golden


def bits_equal(a, b):
    a = np.ascontiguousarray(a, np.float64)
    b = np.ascontiguousarray(b, np.float64)
    return a.shape == b.shape and np.array_equal(a.view(np.uint64), b.view(np.uint64))


def assert_bits_equal(a, b, what=""):
    a = np.ascontiguousarray(a, np.float64)
    b = np.ascontiguousarray(b, np.float64)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    same = a.view(np.uint64) == b.view(np.uint64)
    if not same.all():
        idx = np.argwhere(~same)[0]
        raise AssertionError(f"{what}: {np.count_nonzero(~same)} of {same.size} values differ bitwise; "
                             f"first at {tuple(idx)}: {a[tuple(idx)]!r} vs {b[tuple(idx)]!r}")


# Stated parity tolerance of the engine against the reference numba path:
# bit-exact for the +,-,*,/ models (ABC, Cemaneige); for the models with pow/tanh
# (HBVEdu, GR4J, CemaneigeGR4J) rtol 1e-10 / atol 1e-12 with an identical NaN mask.
RTOL, ATOL = 1e-10, 1e-12


def assert_close(got, ref, what="", rtol=RTOL, atol=ATOL):
    got = np.asarray(got); ref = np.asarray(ref)
    assert got.shape == ref.shape, f"{what}: shape {got.shape} vs {ref.shape}"
    assert np.array_equal(np.isnan(got), np.isnan(ref)), f"{what}: NaN masks differ"
    ok = np.isclose(got, ref, rtol=rtol, atol=atol, equal_nan=True)
    if not ok.all():
        err = np.nanmax(np.abs(got - ref) / (atol + rtol * np.abs(ref)))
        raise AssertionError(f"{what}: {np.count_nonzero(~ok)} values outside rtol={rtol} atol={atol} "
                             f"(worst {err:.3g}x the bound)")
