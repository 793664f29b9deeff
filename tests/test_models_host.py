"""Host-side logic of the drop-in models: BaseModel API, validation (exception types + messages of the
reference), preprocessing and param packing.  Everything here runs before any kernel -> no GPU needed."""
import numpy as np
import pandas as pd
import pytest

import oracle
from rrmpg_b200 import engine, synthetic
from rrmpg_b200.distributed import member_block
from rrmpg_b200.models import ABCModel, HBVEdu, GR4J, Cemaneige, CemaneigeGR4J
from rrmpg_b200.models import cemaneige_utils
from rrmpg_b200.models.basemodel import BaseModel
from rrmpg_b200.tools import monte_carlo
from rrmpg_b200.utils import calc_mse, calc_nse, calc_rmse, check_for_negatives, validate_array_input
from conftest import assert_bits_equal, load_golden

MODELS = [ABCModel, HBVEdu, GR4J, Cemaneige, CemaneigeGR4J]


# ---- BaseModel contract (test/test_models.py:20-77)
def test_abc_names_bounds_dtype():
    m = ABCModel()
    assert m.get_parameter_names() == ['a', 'b', 'c']
    assert m.get_default_bounds() == {'a': (0, 1), 'b': (0, 1), 'c': (0, 1)}
    assert m.get_dtype() == np.dtype([('a', np.float64), ('b', np.float64), ('c', np.float64)])


@pytest.mark.parametrize("cls", MODELS)
def test_random_params_within_bounds_and_count(cls):
    m = cls()
    assert issubclass(cls, BaseModel)
    p = m.get_random_params(num=24)
    assert p.size == 24 and p.dtype == m.get_dtype()
    for name in m.get_parameter_names():
        lo, hi = m.get_default_bounds()[name]
        assert np.all((p[name] >= lo) & (p[name] <= hi))
    if cls is ABCModel:
        assert np.all(p['b'] <= 1 - p['a'])
    assert set(m.get_params()) == set(m.get_parameter_names())


def test_record_layouts_match_the_c_abi_column_order():
    """Field order of each _dtype is the column order the kernels read (include/rrmpg_b200.h)."""
    assert HBVEdu._dtype.names == ('T_t', 'DD', 'FC', 'Beta', 'C', 'PWP', 'K_0', 'K_1', 'K_2', 'K_p', 'L')
    assert GR4J._dtype.names == ('x1', 'x2', 'x3', 'x4')
    assert Cemaneige._dtype.names == ('CTG', 'Kf')
    assert CemaneigeGR4J._dtype.names == ('CTG', 'Kf', 'x1', 'x2', 'x3', 'x4')
    assert CemaneigeGR4J._default_bounds['x4'] == (1.1, 2.9) and HBVEdu._default_bounds['Beta'] == (1, 7)
    for cls in MODELS:
        p = cls().get_random_params(5)
        mat = engine.pack_params(p)
        assert mat.shape == (5, len(cls._param_list)) and mat.flags.c_contiguous
        assert np.shares_memory(mat, p)  # zero-copy view of the record array
        for j, name in enumerate(p.dtype.names):
            assert np.array_equal(mat[:, j], p[name])


def test_seeded_sampling_reproduces_the_reference_draw_order():
    """np.random.seed(s); get_random_params(n) must give the reference's ensemble (basemodel.py:68-91)."""
    g = load_golden("ensemble_hbvedu")
    p = synthetic.random_params(HBVEdu(), 24)
    assert_bits_equal(engine.pack_params(p), g["params"][:24])
    g = load_golden("ensemble_abc")  # ABC has its own sampler (abcmodel.py:70-103)
    assert_bits_equal(engine.pack_params(synthetic.random_params(ABCModel(), 24)), g["params"][:24])


def test_set_params_round_trip_and_errors():
    m = ABCModel()
    rp = m.get_random_params()
    d = {k: rp[k][0] for k in m.get_parameter_names()}
    m.set_params(d)
    assert m.get_params() == d
    m.set_params(rp[0])          # np.void
    m.set_params(rp)             # record array
    with pytest.raises(AttributeError, match="Unknow parameter 'z'"):
        m.set_params({'z': 1.0})
    with pytest.raises(ValueError, match="must be numerical"):
        m.set_params({'a': 'x'})
    with pytest.raises(TypeError, match="wrong data type"):
        m.set_params(GR4J().get_random_params())
    with pytest.raises(TypeError, match="Wrong input data type"):
        m.set_params([1, 2, 3])
    with pytest.raises(AttributeError, match="Missing the following model parameters"):
        ABCModel(params={'a': 0.1})
    assert HBVEdu(params={k: 1.0 for k in HBVEdu._param_list}).get_params()['Beta'] == 1.0


# ---- validation happens before the kernel, with the reference's messages
def test_negative_precipitation_messages():
    with pytest.raises(ValueError, match="In the precipitation array are negative values."):
        ABCModel().simulate([-1, 1, 1])                                     # test_models.py:94-99
    with pytest.raises(ValueError, match="In the precipitation array are negative values."):
        HBVEdu().simulate(temp=np.zeros(100), prec=np.arange(-1, 99), month=np.ones(100, int),
                          PE_m=np.zeros(12), T_m=np.zeros(12))              # test_models.py:131-140
    with pytest.raises(ValueError, match="The precipitation array contains negative values."):
        GR4J().simulate([-1.0, 1.0], [0.0, 0.0])
    with pytest.raises(ValueError, match="The precipitation array contains negative values."):
        Cemaneige().simulate([-1.0], [0.0], [0.0], [0.0], met_station_height=100)
    with pytest.raises(ValueError, match="The precipitation array contains negative values."):
        CemaneigeGR4J().simulate([-1.0], [0.0], [0.0], [0.0], [0.0], met_station_height=100)


def test_argument_errors_of_each_model():
    z = np.zeros(10)
    with pytest.raises(TypeError, match="initial_state"):
        ABCModel().simulate(z, initial_state=-1)
    with pytest.raises(TypeError, match="return_storage arg must be a boolean"):
        ABCModel().simulate(z, return_storage=1)
    with pytest.raises(TypeError, match="models own custom data type"):
        ABCModel().simulate(z, params=GR4J().get_random_params(2))
    h = HBVEdu()
    with pytest.raises(RuntimeError, match="must be of equal size"):
        h.simulate(z, z[:5], np.ones(10, int), np.zeros(12), np.zeros(12))
    with pytest.raises(RuntimeError, match="must be of length 12"):
        h.simulate(z, z, np.ones(10, int), np.zeros(11), np.zeros(12))
    with pytest.raises(ValueError, match="month array must be between"):
        h.simulate(z, z, np.zeros(10, int), np.zeros(12), np.zeros(12))
    with pytest.raises(ValueError, match="month array must be between"):
        h.simulate(z, z, np.full(10, 13), np.zeros(12), np.zeros(12))
    g = GR4J()
    with pytest.raises(RuntimeError, match="must be of the same size"):
        g.simulate(z, z[:4])
    with pytest.raises(TypeError, match="'s1_init' must be a Number"):
        g.simulate(z, z, s_init="a")
    with pytest.raises(ValueError, match="production storage must be in the range"):
        g.simulate(z, z, s_init=1.5)
    with pytest.raises(ValueError, match="routing storage must be in the range"):
        g.simulate(z, z, r_init=-0.1)
    with pytest.raises(ValueError, match="routing storage must be in the range"):
        g.fit(z, z, z, r_init=2)
    c = Cemaneige()
    with pytest.raises(RuntimeError, match="same length"):
        c.simulate(z, z, z[:3], z, met_station_height=10)
    with pytest.raises(TypeError, match="'altitudes' must be a list"):
        c.simulate(z, z, z, z, met_station_height=10, altitudes=(1, 2))
    with pytest.raises(TypeError, match="All elements in 'altitudes must be numbers"):
        c.simulate(z, z, z, z, met_station_height=10, altitudes=[1, "a"])
    with pytest.raises(ValueError, match="height of the meteorological station is missing"):
        c.simulate(z, z, z, z, met_station_height=None, altitudes=[1, 2])
    with pytest.raises(TypeError, match="'met_station_height' must be a Number"):
        c.simulate(z, z, z, z, met_station_height=None)
    with pytest.raises(TypeError, match="'snow_pack_init' must be a Number"):
        c.simulate(z, z, z, z, met_station_height=10, snow_pack_init=[1])
    cg = CemaneigeGR4J()
    with pytest.raises(RuntimeError, match="same length"):
        cg.simulate(z, z, z, z, z[:2], met_station_height=10)
    with pytest.raises(TypeError, match="'r_init' must be a Number"):
        cg.simulate(z, z, z, z, z, met_station_height=10, r_init="x")


def test_caller_arrays_are_never_mutated():
    month = np.arange(1, 11)
    keep = month.copy()
    try:
        HBVEdu().simulate(np.zeros(10), np.zeros(10), month, np.zeros(12), np.zeros(12))
    except RuntimeError:
        pass  # no GPU here: the kernel call fails after validation, which is what we are testing
    assert np.array_equal(month, keep)


# ---- utils (test/test_utils.py)
def test_validate_array_input_and_negatives():
    s = validate_array_input(pd.Series([1, 2, 3]), np.float64, 'x')
    assert isinstance(s, np.ndarray) and s.dtype == np.float64
    assert validate_array_input([[1, 2], [3, 4]], np.float64, 'x').shape == (4,)
    a = np.arange(3.0)
    assert validate_array_input(a, np.float64, 'x') is not a  # always a copy (array_checks.py:62)
    with pytest.raises(ValueError, match="The data in the parameter array 'x' must be purely numerical."):
        validate_array_input([1, 'a'], np.float64, 'x')
    with pytest.raises(TypeError, match="The array x must be either a list, numpy.ndarray or pandas.Series"):
        validate_array_input((1, 2), np.float64, 'x')
    assert check_for_negatives(np.array([1, -1.0])) and not check_for_negatives(np.array([0.0, np.nan, 2]))


def test_metrics_known_values():
    assert calc_mse([1, 2, 3], [1, 2, 3]) == 0 and calc_mse([1, 2, 3], [2, 3, 4]) == 1
    assert calc_rmse([1, 2, 3], [3, 4, 5]) == 2
    assert calc_nse([1, 2, 3], [1, 2, 3]) == 1 and calc_nse([1, 2, 3], [2, 2, 2]) == 0
    with pytest.raises(RuntimeError):
        calc_nse([1, 1, 1], [1, 2, 3])
    with pytest.raises(ValueError, match="same size"):
        calc_mse([1, 2], [1])


# ---- Cemaneige forcing preprocessing (numpy) is bit-identical to the numba originals
def test_cemaneige_preprocessing_matches_reference_bitwise():
    for name in ("fixture_cemaneige", "ensemble_cemaneige"):
        g = load_golden(name)
        alts = [float(a) for a in g["altitudes"]]
        h = float(g["met_station_height"])
        p = cemaneige_utils.extrapolate_precipitation(g["prec"], np.array(alts), h)
        mn, me, mx = cemaneige_utils.extrapolate_temperature(g["min_temp"], g["mean_temp"], g["max_temp"],
                                                             np.array(alts), h)
        fr = cemaneige_utils.calculate_solid_fraction(p, np.array(alts), me, mn, mx)
        assert_bits_equal(p, g["layer_prec"], name + " prec")
        assert_bits_equal(me, g["layer_mean_temp"], name + " mean_temp")
        assert_bits_equal(fr, g["frac_solid"], name + " frac_solid")
    # integer altitudes (a python list, as users pass them) and the high-altitude branches
    g = load_golden("ensemble_cemaneige")
    alts = [int(a) for a in g["altitudes_high"]]
    p = cemaneige_utils.extrapolate_precipitation(g["prec"], np.array(alts), 1450)
    mn, me, mx = cemaneige_utils.extrapolate_temperature(g["min_temp"], g["mean_temp"], g["max_temp"], np.array(alts), 1450)
    fr = cemaneige_utils.calculate_solid_fraction(p, np.array(alts), me, mn, mx)
    assert_bits_equal(p, g["layer_prec_high"]); assert_bits_equal(me, g["layer_mean_temp_high"])
    assert_bits_equal(fr, g["frac_solid_high"])
    # and against the oracle on a case with exact zeros / equal min-max temperatures
    t = np.array([-3.0, 0.0, 0.0, 2.0, 5.0, 1.0]); mnn = np.array([-5.0, 0.0, -1.0, 2.0, 1.0, -1.0])
    mxx = np.array([-1.0, 0.0, 1.0, 2.0, 9.0, 3.0]); pr = np.ones(6)
    for alt in ([800.0], [1700.0]):
        a = cemaneige_utils.calculate_solid_fraction(pr[:, None], np.array(alt), t[:, None], mnn[:, None], mxx[:, None])
        b = oracle.calculate_solid_fraction(pr[:, None], alt, t[:, None], mnn[:, None], mxx[:, None])
        assert_bits_equal(a, b)


def test_monte_carlo_argument_checks():
    with pytest.raises(TypeError, match="must be one of the models"):
        monte_carlo(object(), 3)
    with pytest.raises(TypeError, match="positive integer"):
        monte_carlo(ABCModel(), 0, prec=np.zeros(3))
    with pytest.raises(TypeError, match="positive integer"):
        monte_carlo(ABCModel(), 2.0, prec=np.zeros(3))


def test_member_blocks_partition_the_ensemble():
    for n, w in [(65536, 8), (10, 3), (7, 8), (1, 1), (0, 4)]:
        blocks = [member_block(n, r, w) for r in range(w)]
        assert blocks[0][0] == 0 and blocks[-1][1] == n
        assert all(blocks[i][1] == blocks[i + 1][0] for i in range(w - 1))
        sizes = [b - a for a, b in blocks]
        assert max(sizes) - min(sizes) <= 1


def test_multi_catchment_argument_checks_fail_before_the_library_is_touched():
    z = np.zeros
    with pytest.raises(ValueError, match=r"\[C, N, 4\]"):
        engine.gr4j_multi(z((2, 10)), z((2, 10)), (0.5, 0.5), z((3, 5, 4)))          # one params block per catchment
    with pytest.raises(ValueError, match=r"\[C, N, 4\]"):
        engine.gr4j_multi(z((2, 10)), z((2, 10)), (0.5, 0.5), z((2, 5, 6)))          # GR4J records have 4 fields
    with pytest.raises(ValueError):
        engine.gr4j_multi(z((2, 10)), z((2, 11)), (0.5, 0.5), z((2, 5, 4)))          # etp length
    with pytest.raises(ValueError, match=r"\[C, T, L\]"):
        engine.cemaneigegr4j_multi(z((2, 10)), z((2, 10)), z((2, 10)), z((2, 10)), (0, 0, .5, .5), z((2, 5, 6)))
    with pytest.raises(ValueError):
        engine.cemaneigegr4j_multi(z((2, 10, 3)), z((2, 10, 3)), z((2, 9)), z((2, 10, 3)), (0, 0, .5, .5), z((2, 5, 6)))
    with pytest.raises(ValueError):
        engine.snow_layers(z((4, 2)), z((4, 2)), z((4, 2)), z((4, 2)), 300.0, [400.0])  # station series are 1-D


def test_vectorised_fit_driver_polishes_with_batched_gradients():
    """models/_fit.py: one loss call per DE generation ((k, S) trial matrix) and one per L-BFGS-B gradient
    ((k, k + 1) matrix) -- checked on a numpy objective, no GPU involved."""
    from rrmpg_b200.models import _fit
    target = np.array([0.3, -1.2, 2.5])
    shapes = []

    def loss(X, scale):
        X = np.asarray(X, dtype=np.float64)
        shapes.append(X.shape)
        P = _fit.as_population(X)                       # [S, k]
        return _fit.finish(scale * ((P - target) ** 2).sum(axis=1) + 1.0, X)

    bounds = [(-3.0, 3.0)] * 3
    res = _fit.minimise(loss, bounds, (2.0,), de_kwargs=dict(seed=3, maxiter=30, tol=1e-3))
    assert np.allclose(res.x, target, atol=1e-6) and abs(res.fun - 1.0) < 1e-10
    assert all(len(s) == 2 for s in shapes)             # never a scalar trial vector: every call is a batch
    assert (3, 4) in shapes                             # the polish step: k + 1 members per gradient
    rough = _fit.minimise(loss, bounds, (2.0,), de_kwargs=dict(seed=3, maxiter=30, tol=1e-3, polish=False))
    assert res.fun <= rough.fun and res.nfev > rough.nfev


# ------------------------------------------------------------------ round 2: host-side logic of the new options (no GPU)
def test_fused_objective_argument_errors_are_raised_before_the_library_call():
    from rrmpg_b200 import engine
    f = dict(temp=np.zeros(20), prec=np.ones(20), month0=np.zeros(20, np.int8), PE_m=np.ones(12), T_m=np.zeros(12))
    P = np.ones((3, 11))
    args = (f["temp"], f["prec"], f["month0"], f["PE_m"], f["T_m"], (0, 100, 3, 10), P)
    with pytest.raises(RuntimeError, match="Nash-Sutcliffe"):          # rrmpg/utils/metrics.py:66-70
        engine.hbvedu(*args, qobs=np.full(20, 2.0), objective="nse")
    with pytest.raises(RuntimeError, match="mean of the observations"):  # :166-168
        engine.hbvedu(*args, qobs=np.r_[np.ones(10), -np.ones(10)], objective="kge")
    with pytest.raises(RuntimeError, match="standard deviation"):        # :170-173
        engine.hbvedu(*args, qobs=np.full(20, 2.0), objective="kge")
    with pytest.raises(ValueError, match="objective must be one of"):
        engine.hbvedu(*args, qobs=np.arange(20.0), objective="rmse")
    with pytest.raises(ValueError, match="Arrays must have the same size"):   # :127-128
        engine.hbvedu(*args, qobs=np.arange(19.0))
    with pytest.raises(ValueError, match="month0 must be"):
        engine.hbvedu(f["temp"], f["prec"], np.full(20, 12, np.int8), f["PE_m"], f["T_m"], (0, 100, 3, 10), P)
    with pytest.raises(ValueError, match="devices must be"):
        engine.hbvedu(*args, devices="every")


def test_fused_context_and_monte_carlo_argument_handling():
    from rrmpg_b200 import engine
    from rrmpg_b200.tools import monte_carlo
    with engine.fused(np.ones(4)):
        with pytest.raises(RuntimeError, match="do not nest"):
            with engine.fused(np.ones(4)):
                pass
    assert engine._FUSED is None
    with pytest.raises(ValueError, match="needs qobs"):
        monte_carlo(ABCModel(), num=3, return_qsim=False, prec=np.ones(5))
    with pytest.raises(TypeError):
        monte_carlo(object(), num=3)
    with pytest.raises(TypeError):
        monte_carlo(ABCModel(), num=0)


def test_state_row_layout_of_the_resumable_models():
    from rrmpg_b200 import _lib
    L = _lib.lib()
    assert L.rrb_state_rows(_lib.MODEL_ABC, 0.0) == 1
    assert L.rrb_state_rows(_lib.MODEL_HBVEDU, 0.0) == 4
    assert [L.rrb_state_rows(_lib.MODEL_GR4J, x) for x in (2.9, 3.0, 3.5, 10.0, 64.0)] == [12, 12, 15, 33, 2 + 64 + 129]
    assert L.rrb_state_rows(_lib.MODEL_GR4J, 65.0) == -1 and L.rrb_state_rows(99, 1.0) == -1
