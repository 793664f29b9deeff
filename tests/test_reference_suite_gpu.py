"""The reference's OWN test-suite, verbatim, against the drop-in package (VERDICT r1, next #1c).

`oracle/build_ref.py` installs the unmodified reference under the git-ignored `oracle/_ref/` (it travels to the GPU box
with the snapshot) together with its `test/` directory (test_models.py, test_tools.py, test_utils.py and the MATLAB /
Excel fixtures under test/data).  Here `rrmpg` is aliased to `rrmpg_b200` in `sys.modules`, the reference's test modules
are imported from their files and every `unittest.TestCase` in them is run unchanged: same constructors, same
`simulate` / `monte_carlo` / metric calls, same assertions (`np.allclose` against the MATLAB and Excel series,
test/test_models.py:142-356; literal exception messages :94-140; monte_carlo shapes test/test_tools.py:26-29).
"""
import importlib
import importlib.util
import os
import sys
import unittest

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_TESTS = os.path.join(ROOT, "oracle", "_ref", "reference_tests")
ALIASES = ["", ".models", ".models.basemodel", ".tools", ".tools.monte_carlo", ".utils", ".utils.metrics",
           ".utils.array_checks"]


@pytest.fixture
def rrmpg_is_the_drop_in():
    saved = {k: v for k, v in sys.modules.items() if k == "rrmpg" or k.startswith("rrmpg.")}
    for k in saved:
        del sys.modules[k]
    for suffix in ALIASES:
        sys.modules["rrmpg" + suffix] = importlib.import_module("rrmpg_b200" + suffix)
    yield
    for k in [k for k in sys.modules if k == "rrmpg" or k.startswith("rrmpg.")]:
        del sys.modules[k]
    sys.modules.update(saved)


@pytest.mark.parametrize("module", ["test_models", "test_tools", "test_utils"])
def test_reference_test_module_passes_against_the_drop_in(module, rrmpg_is_the_drop_in):
    path = os.path.join(REF_TESTS, module + ".py")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/reference_tests is absent (run oracle/build_ref.py where /root/reference exists)")
    spec = importlib.util.spec_from_file_location("reference_suite_" + module, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)  # its `from rrmpg... import` lines resolve to rrmpg_b200
    for name in ("ABCModel", "monte_carlo", "calc_mse", "validate_array_input"):
        if hasattr(mod, name):
            assert getattr(mod, name).__module__.startswith("rrmpg_b200"), "the reference package leaked into the run"
    suite = unittest.defaultTestLoader.loadTestsFromModule(mod)
    assert suite.countTestCases() > 0
    result = unittest.TextTestRunner(verbosity=0).run(suite)
    problems = [f"{t.id()}:\n{tb}" for t, tb in result.failures + result.errors]
    assert not problems, f"{len(problems)} of {result.testsRun} reference tests fail against rrmpg_b200:\n" + "\n".join(problems)
    print(f"{module}: {result.testsRun} reference tests passed against rrmpg_b200")
