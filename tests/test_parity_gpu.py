"""GPU parity: the CUDA path (through the C ABI) against the golden vectors made by the live numba
reference and against the CPU oracle on fresh seeded inputs.  Run on the B200 box: pytest -m gpu."""
import numpy as np
import pytest

import oracle
from rrmpg_b200 import engine, synthetic
from rrmpg_b200.models import (ABCModel, HBVEdu, GR4J, Cemaneige, CemaneigeGR4J, CemaneigeGR4JIce,
                               CemaneigeHystGR4J, CemaneigeHystGR4JIce)
from rrmpg_b200.tools import monte_carlo
from conftest import assert_bits_equal, assert_close, load_golden

pytestmark = pytest.mark.gpu

MATHS = ["fast", "precise"]


# ------------------------------------------------------------------ reference fixtures (KATs)
def test_fixture_hbvedu_matlab():
    g = load_golden("fixture_hbvedu")
    names = ['T_t', 'DD', 'FC', 'Beta', 'C', 'PWP', 'K_0', 'K_1', 'K_2', 'K_p', 'L']
    model = HBVEdu(params=dict(zip(names, [float(v) for v in g["params"][0]])))
    out = model.simulate(temp=g["temp"], prec=g["prec"], month=g["month"], PE_m=g["PE_m"], T_m=g["T_m"],
                         snow_init=0, soil_init=100, s1_init=3, s2_init=10, return_storage=True)
    # the reference's own criterion (test/test_models.py:172-174)
    assert np.allclose((out[0] * 410 * 1000 / (24 * 60 * 60)).flatten(), g["expected"])
    for nm, a in zip(["qsim", "snow", "soil", "s1", "s2"], out):
        assert_close(a, g[nm], "fixture_hbvedu." + nm)


def test_fixture_gr4j_excel():
    g = load_golden("fixture_gr4j")
    model = GR4J(params=dict(zip(['x1', 'x2', 'x3', 'x4'], [float(v) for v in g["params"][0]])))
    out = model.simulate(g["prec"], g["etp"], s_init=0.6, r_init=0.7, return_storage=True)
    assert np.allclose(out[0].flatten(), g["expected"])  # test/test_models.py:201-210
    for nm, a in zip(["qsim", "s_store", "r_store"], out):
        assert_close(a, g[nm], "fixture_gr4j." + nm)
    # return_storage=False gives the same discharge
    assert_bits_equal(model.simulate(g["prec"], g["etp"], s_init=0.6, r_init=0.7), out[0], "gr4j qsim-only")


def test_fixture_cemaneige_excel():
    g = load_golden("fixture_cemaneige")
    model = Cemaneige(params={'CTG': 0.25, 'Kf': 3.74})
    out = model.simulate(g["prec"], g["mean_temp"], g["min_temp"], g["max_temp"], met_station_height=495,
                         altitudes=[550, 620, 700, 785, 920], return_storages=True)
    assert np.allclose(out[0].flatten(), g["expected"])  # test/test_models.py:227-236
    for nm, a in zip(["outflow", "G", "eTG"], out):
        assert_bits_equal(a, g[nm], "fixture_cemaneige." + nm)


def test_fixture_cemaneigegr4j_excel():
    g = load_golden("fixture_cemaneigegr4j")
    names = ['CTG', 'Kf', 'x1', 'x2', 'x3', 'x4']
    model = CemaneigeGR4J(params=dict(zip(names, [float(v) for v in g["params"][0]])))
    out = model.simulate(g["prec"], g["mean_temp"], g["min_temp"], g["max_temp"], g["etp"],
                         met_station_height=495, altitudes=[550, 620, 700, 785, 920], s_init=0.6, r_init=0.7,
                         return_storages=True)
    assert np.allclose(out[0].flatten(), g["expected"])  # test/test_models.py:258-268
    assert_bits_equal(out[1], g["G"], "fixture_cemaneigegr4j.G")
    assert_bits_equal(out[2], g["eTG"], "fixture_cemaneigegr4j.eTG")
    for nm, a in zip(["qsim", "s_store", "r_store"], (out[0], out[3], out[4])):
        assert_close(a, g[nm], "fixture_cemaneigegr4j." + nm)


# ------------------------------------------------------------------ numba ensembles (golden)
@pytest.mark.parametrize("math", MATHS)
def test_ensemble_abc(math):
    g = load_golden("ensemble_abc")
    r = engine.abc(g["prec"], float(g["initial_state"]), g["params"], return_storage=True, math=math)
    assert_bits_equal(r["qsim"], g["qsim"], "abc.qsim")
    assert_bits_equal(r["storage"], g["storage"], "abc.storage")


@pytest.mark.parametrize("math", MATHS)
def test_ensemble_hbvedu(math):
    g = load_golden("ensemble_hbvedu")
    m0 = (g["month"] - 1).astype(np.int8)
    r = engine.hbvedu(g["temp"], g["prec"], m0, g["PE_m"], g["T_m"], g["inits"], g["params"],
                      return_storage=True, math=math)
    for nm in ["qsim", "snow", "soil", "s1", "s2"]:
        assert_close(r[nm], g[nm], f"hbvedu[{math}]." + nm)
    assert_bits_equal(r["snow"], g["snow"], "hbvedu.snow")  # the snow store has no pow: exact
    r = engine.hbvedu(g["temp"] - 6, g["prec"], m0, g["PE_m"], g["T_m"], g["inits_cold"], g["params"],
                      return_storage=True, math=math)
    assert_close(r["qsim"], g["qsim_cold"], f"hbvedu[{math}].cold.qsim")
    assert_close(r["soil"], g["soil_cold"], f"hbvedu[{math}].cold.soil")


@pytest.mark.parametrize("math", MATHS)
def test_ensemble_gr4j(math):
    g = load_golden("ensemble_gr4j")
    r = engine.gr4j(g["prec"], g["etp"], 0.6, 0.7, g["params"], return_storage=True, math=math)
    for nm in ["qsim", "s_store", "r_store"]:
        assert_close(r[nm], g[nm], f"gr4j[{math}]." + nm)
    # long unit hydrographs: x4 = 3.1 / 4.5 / 7.25 / 10 / 15.5 in ONE batch and one by one
    r = engine.gr4j(g["prec"], g["etp"], 0.3, 0.5, g["params_longuh"], return_storage=True, math=math)
    for nm in ["qsim", "s_store", "r_store"]:
        assert_close(r[nm], g[nm + "_longuh"], f"gr4j[{math}].longuh." + nm)
    for i in range(g["params_longuh"].shape[0]):
        r = engine.gr4j(g["prec"], g["etp"], 0.3, 0.5, g["params_longuh"][i:i + 1], math=math)
        assert_close(r["qsim"][:, 0], g["qsim_longuh"][:, i], f"gr4j[{math}].longuh[{i}]")


@pytest.mark.parametrize("math", MATHS)
def test_ensemble_cemaneige(math):
    g = load_golden("ensemble_cemaneige")
    r = engine.cemaneige(g["layer_prec"], g["layer_mean_temp"], g["frac_solid"], 0.0, 0.0, g["params"],
                         return_storages=True, math=math)
    for nm in ["outflow", "G", "eTG"]:
        assert_bits_equal(r[nm], g[nm], f"cemaneige[{math}]." + nm)
    r = engine.cemaneige(g["prec"][:, None], g["mean_temp"][:, None], g["frac_solid_L1"], 12.0, -1.5,
                         g["params"], return_storages=True, math=math)
    for nm in ["outflow", "G", "eTG"]:
        assert_bits_equal(r[nm], g[nm + "_L1"], f"cemaneige[{math}].L1." + nm)
    r = engine.cemaneige(g["layer_prec_high"], g["layer_mean_temp_high"], g["frac_solid_high"], 0.0, 0.0,
                         g["params"], math=math)
    assert_bits_equal(r["outflow"], g["outflow_high"], f"cemaneige[{math}].high.outflow")


@pytest.mark.parametrize("math", MATHS)
def test_ensemble_cemaneigegr4j(math):
    g = load_golden("ensemble_cemaneigegr4j")
    c = load_golden("ensemble_cemaneige")
    r = engine.cemaneigegr4j(c["layer_prec"], c["layer_mean_temp"], g["etp"], c["frac_solid"], g["inits"],
                             g["params"], return_storages=True, math=math)
    assert_bits_equal(r["G"], g["G"], "cemaneigegr4j.G")
    assert_bits_equal(r["eTG"], g["eTG"], "cemaneigegr4j.eTG")
    for nm in ["qsim", "s_store", "r_store"]:
        assert_close(r[nm], g[nm], f"cemaneigegr4j[{math}]." + nm)
    r = engine.cemaneigegr4j(g["prec"][:, None], g["mean_temp"][:, None], g["etp"], c["frac_solid_L1"],
                             g["inits_L1"], g["params"], return_storages=True, math=math)
    assert_close(r["qsim"], g["qsim_L1"], f"cemaneigegr4j[{math}].L1.qsim")
    assert_close(r["s_store"], g["s_store_L1"], f"cemaneigegr4j[{math}].L1.s_store")


# ------------------------------------------------------------------ drop-in models vs oracle, fresh inputs
def _hbv_case(T, N, seed=1):
    f = synthetic.forcing(T, seed=synthetic.SEED + seed)
    P = synthetic.random_params(HBVEdu(), N, seed=synthetic.PARAM_SEED + seed)
    return f, P


@pytest.mark.parametrize("N", [1, 31, 64, 1000, 4099])
def test_hbvedu_model_vs_oracle_ragged_sizes(N):
    f, P = _hbv_case(1500, N)
    q = HBVEdu().simulate(f["temp"], f["prec"], f["month"], f["PE_m"], f["T_m"], params=P, **synthetic.HBV_INITS)
    ref = oracle.hbvedu(f["temp"], f["prec"], f["month"] - 1, f["PE_m"], f["T_m"], (0, 100, 3, 10), P)
    assert q.shape == (1500, N)
    assert_close(q, ref, f"hbvedu N={N}")


@pytest.mark.parametrize("T", [1, 2, 127, 128, 129, 385])
def test_hbvedu_short_series(T):
    f, P = _hbv_case(T, 70, seed=3)
    r = engine.hbvedu(f["temp"], f["prec"], f["month"] - 1, f["PE_m"], f["T_m"], (1, 100, 3, 10), P,
                      return_storage=True)
    ref = oracle.hbvedu(f["temp"], f["prec"], f["month"] - 1, f["PE_m"], f["T_m"], (1, 100, 3, 10), P,
                        return_storage=True)
    for nm, b in zip(["qsim", "snow", "soil", "s1", "s2"], ref):
        assert_close(r[nm], b, f"hbvedu T={T} {nm}")


def test_single_member_from_attributes_and_void():
    f = synthetic.forcing(400)
    m = GR4J()
    q1 = m.simulate(f["prec"], f["etp"], s_init=0.5, r_init=0.5)
    P = np.zeros(1, m.get_dtype())
    for k, v in m.get_params().items():
        P[k] = v
    q2 = m.simulate(f["prec"], f["etp"], s_init=0.5, r_init=0.5, params=P[0])  # np.void
    assert q1.shape == (400, 1)
    assert_bits_equal(q1, q2, "params=None vs np.void")
    assert_close(q1, oracle.gr4j(f["prec"], f["etp"], 0.5, 0.5, P), "gr4j single")


def test_gr4j_multi_member_without_storage_computes_all_members():
    """The reference returns after member 0 here (rrmpg/models/gr4j.py:178); the engine must not."""
    f = synthetic.forcing(600)
    P = synthetic.random_params(GR4J(), 17)
    q = GR4J().simulate(f["prec"], f["etp"], s_init=0.6, r_init=0.7, params=P)
    assert np.all(np.abs(q).sum(axis=0) > 0)
    assert_close(q, oracle.gr4j(f["prec"], f["etp"], 0.6, 0.7, P), "gr4j all members")


def test_zero_rain_gives_zero_discharge():
    # test/test_models.py:90-92, :123-129, :195-199
    assert np.sum(ABCModel().simulate(np.zeros(100))) == 0
    rng = np.random.default_rng(5)
    q = HBVEdu().simulate(temp=rng.uniform(-15, 25, 100), prec=np.zeros(100), month=rng.integers(1, 12, 100),
                          PE_m=rng.uniform(0, 4, 12), T_m=rng.uniform(-5, 15, 12))
    assert np.sum(q) == 0
    q = GR4J().simulate(prec=np.zeros(100), etp=rng.uniform(0, 3, 100), s_init=0, r_init=0)
    assert np.sum(q) == 0


def test_cemaneige_models_vs_oracle_without_altitudes_and_layer_counts():
    f = synthetic.forcing(900, seed=77)
    P = synthetic.random_params(Cemaneige(), 45)
    out = Cemaneige().simulate(f["prec"], f["temp"], f["min_temp"], f["max_temp"], met_station_height=900,
                               snow_pack_init=3, thermal_state_init=-0.25, params=P, return_storages=True)
    fr = oracle.calculate_solid_fraction(f["prec"][:, None], [900.0], f["temp"][:, None], f["min_temp"][:, None],
                                         f["max_temp"][:, None])
    ref = oracle.cemaneige(f["prec"][:, None], f["temp"][:, None], fr, 3.0, -0.25, P, return_storages=True)
    for a, b, nm in zip(out, ref, ["outflow", "G", "eTG"]):
        assert_bits_equal(a, b, "cemaneige no-altitudes " + nm)
    for alts in ([600, 800], [500, 700, 900, 1100, 1300, 1600, 2100], list(range(400, 2000, 100))):
        out = Cemaneige().simulate(f["prec"], f["temp"], f["min_temp"], f["max_temp"], met_station_height=450,
                                   altitudes=alts, params=P, return_storages=True)
        p = oracle.extrapolate_precipitation(f["prec"], alts, 450)
        mn, me, mx = oracle.extrapolate_temperature(f["min_temp"], f["temp"], f["max_temp"], alts, 450)
        fr = oracle.calculate_solid_fraction(p, alts, me, mn, mx)
        ref = oracle.cemaneige(p, me, fr, 0.0, 0.0, P, return_storages=True)
        for a, b, nm in zip(out, ref, ["outflow", "G", "eTG"]):
            assert_bits_equal(a, b, f"cemaneige L={len(alts)} {nm}")
    PC = synthetic.random_params(CemaneigeGR4J(), 33)
    alts = [500, 700, 900]
    out = CemaneigeGR4J().simulate(f["prec"], f["temp"], f["min_temp"], f["max_temp"], f["etp"],
                                   met_station_height=450, altitudes=alts, s_init=0.4, r_init=0.3, params=PC,
                                   return_storages=True)
    p = oracle.extrapolate_precipitation(f["prec"], alts, 450)
    mn, me, mx = oracle.extrapolate_temperature(f["min_temp"], f["temp"], f["max_temp"], alts, 450)
    fr = oracle.calculate_solid_fraction(p, alts, me, mn, mx)
    ref = oracle.cemaneigegr4j(p, me, f["etp"], fr, (0, 0, 0.4, 0.3), PC, return_storages=True)
    for a, b, nm in zip(out, ref, ["qsim", "G", "eTG", "s_store", "r_store"]):
        assert_close(a, b, "cemaneigegr4j L=3 " + nm)


def test_too_many_layers_and_too_long_uh_fail_loudly():
    f = synthetic.forcing(50)
    with pytest.raises(RuntimeError, match="elevation layers"):
        Cemaneige().simulate(f["prec"], f["temp"], f["min_temp"], f["max_temp"], met_station_height=450,
                             altitudes=list(range(500, 500 + 17 * 50, 50)))
    P = synthetic.random_params(GR4J(), 3)
    P["x4"][1] = 70.0
    with pytest.raises(RuntimeError, match="x4"):
        GR4J().simulate(f["prec"], f["etp"], params=P)


def test_nan_propagation_matches_oracle():
    f = synthetic.forcing(300)
    P = synthetic.random_params(HBVEdu(), 40)
    P["FC"][3] = -150.0   # negative base of the pow -> NaN from the first wet step on
    P["Beta"][7] = np.nan
    P["T_t"][11] = np.copysign(np.nan, -1.0)   # a NaN threshold (either sign) is never "cold" in the reference
    P["T_t"][12] = np.nan
    P["L"][13] = np.nan                        # max(0, s1 - NaN) = 0 in numba
    P["DD"][14] = np.nan
    prec = f["prec"].copy()
    q = engine.hbvedu(f["temp"], prec, f["month"] - 1, f["PE_m"], f["T_m"], (0, 100, 3, 10), P)["qsim"]
    ref = oracle.hbvedu(f["temp"], prec, f["month"] - 1, f["PE_m"], f["T_m"], (0, 100, 3, 10), P)
    assert np.isnan(ref[:, 3]).any()
    assert_close(q, ref, "hbvedu NaN members")


def test_hbvedu_non_finite_precipitation_takes_the_reference_order_kernel():
    """The FAST HBV snow routine forms prec - prec for "no liquid water", exact for finite rain only; the packer
    flags inf / NaN precipitation and the PRECISE kernel launched behind the FAST one takes the launch."""
    f = synthetic.forcing(300)
    P = synthetic.random_params(HBVEdu(), 70, seed=3)
    cold_day = int(np.argmin(f["temp"][:200]))   # the reference stores that day's precipitation as snow: no liquid water
    for bad in (np.nan, np.inf):
        prec = f["prec"].copy()
        prec[cold_day] = bad
        prec[250] = bad
        ref = oracle.hbvedu(f["temp"], prec, f["month"] - 1, f["PE_m"], f["T_m"], (0, 100, 3, 10), P, return_storage=True)
        got = engine.hbvedu(f["temp"], prec, f["month"] - 1, f["PE_m"], f["T_m"], (0, 100, 3, 10), P, return_storage=True,
                            math="fast")
        for nm, r in zip(["qsim", "snow", "soil", "s1", "s2"], ref):
            if np.isnan(bad):
                assert_close(got[nm], r, f"hbvedu prec={bad} {nm}")
            else:
                # an infinite snow pack melts DD (temp - T_t) mm every warm day, the soil overshoots its capacity and
                # (soil/FC)^Beta amplifies last-bit differences of pow by orders of magnitude per step: compare the
                # special-value pattern over the whole series and the values up to the first infinite day only
                assert np.array_equal(np.isnan(got[nm]), np.isnan(r)) and np.array_equal(np.isinf(got[nm]), np.isinf(r)), nm
                assert_close(got[nm][:cold_day + 1], r[:cold_day + 1], f"hbvedu prec={bad} {nm}")


def test_gr4j_fast_path_contract_falls_back_to_reference_arithmetic():
    """The FAST GR4J step has no special-value handling; CTAs with a member outside its contract (or a
    non-finite / huge forcing value anywhere in the series) run the reference-order step instead.  Both
    routes must agree with the oracle, NaN mask included."""
    f = synthetic.forcing(400)
    P = synthetic.random_params(GR4J(), 200, seed=5)
    P["x1"][3] = 1e-5      # S/x1 astronomically large: (1 + u^4) overflows in the reference too
    P["x3"][70] = np.nan
    P["x2"][131] = 1e9
    P["x1"][199] = -20.0   # negative capacity: tanh of a negative argument, negative stores
    ref = oracle.gr4j(f["prec"], f["etp"], 0.6, 0.7, P, return_storage=True)
    got = engine.gr4j(f["prec"], f["etp"], 0.6, 0.7, P, return_storage=True, math="fast")
    for nm, r in zip(["qsim", "s_store", "r_store"], ref):
        assert_close(got[nm], r, "gr4j insane members " + nm)
    # non-finite and huge forcing: every CTA takes the reference-order step
    P = synthetic.random_params(GR4J(), 100, seed=6)
    for bad in (np.nan, np.inf, 1e9):
        etp = f["etp"].copy(); etp[250] = bad
        prec = f["prec"].copy(); prec[300] = abs(bad) if np.isfinite(bad) else f["prec"][300]
        ref = oracle.gr4j(prec, etp, 0.6, 0.7, P)
        got = engine.gr4j(prec, etp, 0.6, 0.7, P, math="fast")["qsim"]
        assert_close(got, ref, f"gr4j forcing with {bad}")
    # zero stores, zero rain, x2 strongly negative (routing store clamps at 0 -> w = 0 in w^3.5)
    P = synthetic.random_params(GR4J(), 64, seed=7)
    P["x2"][:] = -50.0
    ref = oracle.gr4j(f["prec"] * 0, f["etp"], 0.0, 0.0, P, return_storage=True)
    got = engine.gr4j(f["prec"] * 0, f["etp"], 0.0, 0.0, P, return_storage=True, math="fast")
    for nm, r in zip(["qsim", "s_store", "r_store"], ref):
        assert_close(got[nm], r, "gr4j dry catchment " + nm)
    # initial fractions outside [0, 1] (the wrapper rejects them, the C ABI does not): a negative routing store makes
    # (R/x3)^3.5 a NaN that numba's max(0, NaN) = 0 absorbs -- such calls take the reference-order step
    P = synthetic.random_params(GR4J(), 64, seed=9)
    for si, ri in ((0.5, -0.3), (-0.4, 0.5), (1.7, 2.5)):
        ref = oracle.gr4j(f["prec"], f["etp"], si, ri, P, return_storage=True)
        got = engine.gr4j(f["prec"], f["etp"], si, ri, P, return_storage=True, math="fast")
        for nm, r in zip(["qsim", "s_store", "r_store"], ref):
            assert_close(got[nm], r, f"gr4j inits ({si}, {ri}) " + nm)
    # an extreme storm relative to a tiny production store: tanh saturates (argument clamp of the FAST path)
    P = synthetic.random_params(GR4J(), 64, seed=8)
    P["x1"][:] = np.linspace(0.01, 5.0, 64)
    prec = f["prec"].copy(); prec[100:110] = 5000.0
    ref = oracle.gr4j(prec, f["etp"], 0.6, 0.7, P)
    got = engine.gr4j(prec, f["etp"], 0.6, 0.7, P, math="fast")["qsim"]
    assert_close(got, ref, "gr4j saturated tanh")


def test_cemaneigegr4j_fast_path_contract_falls_back_to_reference_arithmetic():
    c = load_golden("ensemble_cemaneige")
    g = load_golden("ensemble_cemaneigegr4j")
    Pm = np.array(g["params"], copy=True)   # [N, 6] = (CTG, Kf, x1, x2, x3, x4)
    Pm[1, 1] = np.nan      # Kf: NaN melt -> NaN water into GR4J
    Pm[5, 2] = 1e-6        # x1 outside the FAST contract
    for etp_bad in (None, np.nan):
        etp = g["etp"].copy()
        if etp_bad is not None:
            etp[77] = etp_bad
        ref = oracle.cemaneigegr4j(c["layer_prec"], c["layer_mean_temp"], etp, c["frac_solid"], g["inits"], Pm,
                                   return_storages=True)
        got = engine.cemaneigegr4j(c["layer_prec"], c["layer_mean_temp"], etp, c["frac_solid"], g["inits"], Pm,
                                   return_storages=True, math="fast")
        for nm, r in zip(["qsim", "G", "eTG", "s_store", "r_store"], ref):
            assert_close(got[nm], r, f"cemaneigegr4j insane ({etp_bad}) " + nm)


# ------------------------------------------------------------------ time-slab pipeline, objective, device mode
@pytest.mark.parametrize("slab", [1, 7, 128, 300])
def test_time_slab_pipeline_is_bit_identical_to_one_launch(slab):
    f, P = _hbv_case(700, 257, seed=9)
    args = (f["temp"], f["prec"], f["month"] - 1, f["PE_m"], f["T_m"], (0, 100, 3, 10), P)
    one = engine.hbvedu(*args, return_storage=True, slab_steps=700)
    many = engine.hbvedu(*args, return_storage=True, slab_steps=slab)
    for nm in one:
        assert_bits_equal(many[nm], one[nm], f"hbvedu slab={slab} {nm}")
    PG = synthetic.random_params(CemaneigeGR4J(), 130)
    c = load_golden("ensemble_cemaneige")
    etp = load_golden("ensemble_cemaneigegr4j")["etp"]
    args = (c["layer_prec"], c["layer_mean_temp"], etp, c["frac_solid"], (0, 0, 0.6, 0.7), PG)
    one = engine.cemaneigegr4j(*args, return_storages=True, slab_steps=1096)
    many = engine.cemaneigegr4j(*args, return_storages=True, slab_steps=slab)
    for nm in one:
        assert_bits_equal(many[nm], one[nm], f"cemaneigegr4j slab={slab} {nm}")


def test_fused_mse_matches_calc_mse_of_the_reference():
    g = load_golden("ensemble_hbvedu")
    m = load_golden("ensemble_hbvedu_mse")
    r = engine.hbvedu(g["temp"], g["prec"], (g["month"] - 1).astype(np.int8), g["PE_m"], g["T_m"], g["inits"],
                      g["params"], qobs=m["qobs"], want_qsim=False)
    assert "qsim" not in r
    np.testing.assert_allclose(r["mse"], m["mse"], rtol=1e-9)
    # slab-carried accumulator
    r2 = engine.hbvedu(g["temp"], g["prec"], (g["month"] - 1).astype(np.int8), g["PE_m"], g["T_m"], g["inits"],
                       g["params"], qobs=m["qobs"], slab_steps=100)
    np.testing.assert_allclose(r2["mse"], m["mse"], rtol=1e-9)
    assert_close(r2["qsim"], g["qsim"], "qsim with objective")


def test_monte_carlo_drop_in():
    f = synthetic.forcing(500)
    np.random.seed(42)
    res = monte_carlo(ABCModel(), num=100, prec=f["prec"][:100])
    assert res["qsim"].shape[1] == 100  # test/test_tools.py:26-29
    np.random.seed(7)
    qobs = f["prec"] * 0.3
    res = monte_carlo(HBVEdu(), num=64, qobs=qobs, temp=f["temp"], prec=f["prec"], month=f["month"],
                      PE_m=f["PE_m"], T_m=f["T_m"], soil_init=100)
    assert set(res) == {"params", "qsim", "mse"}
    ref = oracle.hbvedu(f["temp"], f["prec"], f["month"] - 1, f["PE_m"], f["T_m"], (0, 100, 0, 0), res["params"])
    assert_close(res["qsim"], ref, "monte_carlo qsim")
    np.testing.assert_allclose(res["mse"], oracle.mse_columns(qobs, ref), rtol=1e-9)


def test_device_mode_torch_tensors():
    import torch
    f, P = _hbv_case(900, 513, seed=4)
    dev = torch.device("cuda:0")
    t = lambda a, dt=torch.float64: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=dev)
    r = engine.hbvedu(t(f["temp"]), t(f["prec"]), t(f["month"] - 1, torch.int8), t(f["PE_m"]), t(f["T_m"]),
                      (0, 100, 3, 10), t(engine.pack_params(P)), return_storage=True)
    assert r["qsim"].is_cuda
    host = engine.hbvedu(f["temp"], f["prec"], f["month"] - 1, f["PE_m"], f["T_m"], (0, 100, 3, 10), P,
                         return_storage=True)
    torch.cuda.synchronize()
    for nm in host:
        assert_bits_equal(r[nm].cpu().numpy(), host[nm], "device vs host mode " + nm)
    PG = synthetic.random_params(GR4J(), 200)
    r = engine.gr4j(t(f["prec"]), t(f["etp"]), 0.6, 0.7, t(engine.pack_params(PG)))  # x4_max reduced on device
    assert_close(r["qsim"].cpu().numpy(), oracle.gr4j(f["prec"], f["etp"], 0.6, 0.7, PG), "gr4j device mode")


# ------------------------------------------------------------------ full-size, size-independent properties
def test_full_size_hbvedu_properties():
    """BASELINE config 2 shape (65 536 members x 14 610 steps): sampled columns against the oracle, the
    member axis is a pure batch axis (column i does not depend on its neighbours or on N)."""
    f = synthetic.forcing(synthetic.T_DAILY_40Y)
    N = 65536
    P = synthetic.random_params(HBVEdu(), N)
    m0 = f["month"] - 1
    q = engine.hbvedu(f["temp"], f["prec"], m0, f["PE_m"], f["T_m"], (0, 100, 3, 10), P)["qsim"]
    assert q.shape == (synthetic.T_DAILY_40Y, N) and np.isfinite(q).all() and (q[0] == 0).all()
    idx = np.r_[0:64, N - 64:N, np.random.default_rng(0).integers(0, N, 128)]
    ref = oracle.hbvedu(f["temp"], f["prec"], m0, f["PE_m"], f["T_m"], (0, 100, 3, 10), P[idx])
    assert_close(q[:, idx], ref, "hbvedu 65k sampled columns")
    sub = engine.hbvedu(f["temp"], f["prec"], m0, f["PE_m"], f["T_m"], (0, 100, 3, 10), P[idx])["qsim"]
    assert_bits_equal(sub, q[:, idx], "batch independence")


# ------------------------------------------------------------------ multi-catchment batch, fit(), empty inputs
def test_multi_catchment_equals_a_loop_over_catchments():
    Cn, T, N = 5, 400, 70
    fs = [synthetic.forcing(T, seed=100 + c) for c in range(Cn)]
    P = np.stack([engine.pack_params(synthetic.random_params(HBVEdu(), N, seed=200 + c)) for c in range(Cn)])
    temp = np.stack([f["temp"] for f in fs]); prec = np.stack([f["prec"] for f in fs])
    m0 = np.stack([f["month"] - 1 for f in fs]).astype(np.int8)
    PE = np.stack([f["PE_m"] * (1 + 0.1 * c) for c, f in enumerate(fs)]); TM = np.stack([f["T_m"] - c for c, f in enumerate(fs)])
    inits = np.array([[c, 100 + 5 * c, 3, 10] for c in range(Cn)], float)
    qobs = np.abs(np.random.default_rng(3).normal(1.0, 0.5, (Cn, T)))
    multi = engine.hbvedu_multi(temp, prec, m0, PE, TM, inits, P, return_storage=True, qobs=qobs)
    assert multi["qsim"].shape == (Cn, T, N) and multi["mse"].shape == (Cn, N)
    for c in range(Cn):
        one = engine.hbvedu(temp[c], prec[c], m0[c], PE[c], TM[c], inits[c], P[c], return_storage=True, qobs=qobs[c])
        for nm in one:
            assert_bits_equal(multi[nm][c], one[nm], f"catchment {c} {nm}")
        assert_close(multi["qsim"][c], oracle.hbvedu(temp[c], prec[c], m0[c], PE[c], TM[c], inits[c], P[c]), f"catchment {c}")
    import torch  # device mode, objective only (the only feasible mode for BASELINE config 5)
    dev = torch.device("cuda:0")
    t = lambda a, dt=torch.float64: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=dev)
    r = engine.hbvedu_multi(t(temp), t(prec), t(m0, torch.int8), t(PE), t(TM), inits, t(P), qobs=t(qobs), want_qsim=False)
    torch.cuda.synchronize()
    assert set(r) == {"mse"}
    assert_bits_equal(r["mse"].cpu().numpy(), multi["mse"], "device-mode objective")


def test_fit_is_population_vectorised_and_finds_the_generating_parameters():
    f = synthetic.forcing(600)
    truth = {'a': 0.3, 'b': 0.25, 'c': 0.12}
    qobs = ABCModel(params=truth).simulate(f["prec"], initial_state=2.0).flatten()
    res = ABCModel().fit(qobs, f["prec"], initial_state=2.0)
    assert res.success and res.fun < 1e-10
    np.testing.assert_allclose(res.x, [truth['a'], truth['b'], truth['c']], atol=1e-4)
    truth = dict(zip(GR4J._param_list, [350.0, 0.8, 90.0, 1.7]))
    qobs = GR4J(params=truth).simulate(f["prec"], f["etp"], s_init=0.5, r_init=0.5).flatten()
    from rrmpg_b200.models import gr4j as gr4j_mod
    X = np.array([[350.0, 300.0], [0.8, 1.0], [90.0, 80.0], [1.7, 2.0]])  # (k, S) trial matrix as scipy passes it
    loss = gr4j_mod._loss(X, qobs, f["prec"], f["etp"], 0.5, 0.5, GR4J._dtype)
    assert loss.shape == (2,) and loss[0] < 1e-20 and loss[1] > 1e-3
    assert isinstance(gr4j_mod._loss(X[:, 0], qobs, f["prec"], f["etp"], 0.5, 0.5, GR4J._dtype), float)


def test_empty_series_and_empty_ensembles():
    P = synthetic.random_params(GR4J(), 4)
    q = GR4J().simulate([], [], params=P)
    assert q.shape == (0, 4)
    r = engine.hbvedu(np.zeros(5), np.zeros(5), np.zeros(5, np.int8), np.zeros(12), np.zeros(12), (0, 0, 0, 0),
                      np.zeros((0, 11)))
    assert r["qsim"].shape == (5, 0)


# ------------------------------------------------------------------ snow-ice family
SNOWICE = [("fixture_cemaneigehystgr4j", CemaneigeHystGR4J), ("fixture_cemaneigehystgr4jice", CemaneigeHystGR4JIce),
           ("ensemble_cemaneigegr4jice", CemaneigeGR4JIce), ("ensemble_cemaneigehystgr4j", CemaneigeHystGR4J),
           ("ensemble_cemaneigehystgr4jice", CemaneigeHystGR4JIce),
           ("ensemble_cemaneigehystgr4jice_L1", CemaneigeHystGR4JIce),
           ("ensemble_cemaneigehystgr4jice_T1", CemaneigeHystGR4JIce)]
EXACT_KEYS = {"G", "eTG", "sca", "icemelt", "snowmelt", "rain"}  # no pow / tanh on these: bit-exact


@pytest.mark.parametrize("name,cls", SNOWICE)
@pytest.mark.parametrize("math", MATHS)
def test_snow_ice_family_drop_ins(name, cls, math, monkeypatch):
    monkeypatch.setattr(engine, "DEFAULT_MATH", math)
    g = load_golden(name)
    model = cls()
    P = np.zeros(g["params"].shape[0], model.get_dtype())
    for j, k in enumerate(model.get_dtype().names):
        P[k] = g["params"][:, j]
    ini = g["inits"]  # (snow_pack_init, thermal_state_init, sca_init, s_init, r_init)
    kw = dict(met_station_height=float(g["met_station_height"]), altitudes=[float(a) for a in g["altitudes"]],
              snow_pack_init=ini[0], thermal_state_init=ini[1], s_init=ini[3], r_init=ini[4], params=P,
              return_storages=True)
    args = [g["prec"], g["mean_temp"], g["min_temp"], g["max_temp"], g["etp"]]
    if model._ice:
        args.append(g["frac_ice"])
    if model._hyst:
        kw["sca_init"] = ini[2]
    out = model.simulate(*args, **kw)
    keys = ["qsim", "G", "eTG", "s_store", "r_store"] + (["sca"] if model._hyst else []) + \
           (["icemelt"] if model._ice else []) + (["snowmelt"] if model._hyst and model._ice else []) + \
           (["rain"] if model._hyst else [])
    assert len(out) == len(keys)
    for k, a in zip(keys, out):
        ref = g[k]
        if k == "rain":
            ref = np.repeat(ref[:, :, None], P.size, axis=2)
        if k in EXACT_KEYS:
            assert_bits_equal(a, ref, f"{name}[{math}].{k}")
        else:
            assert_close(a, ref, f"{name}[{math}].{k}")
    if "expected" in g:  # the reference's own criterion, test/test_models.py:270-356
        assert np.allclose(out[0].flatten(), g["expected"])
    kw["return_storages"] = False
    assert_bits_equal(model.simulate(*args, **kw), out[0], "qsim-only call")


def test_snow_ice_time_slabs_and_objective():
    g = load_golden("ensemble_cemaneigehystgr4jice")
    from test_oracle import snowice_layers
    p, me, fr = snowice_layers(g)
    args = (True, True, p, me, g["etp"], g["frac_ice"], fr, g["inits"], g["params"])
    one = engine.snowice_gr4j(*args, return_storages=True, slab_steps=p.shape[0])
    many = engine.snowice_gr4j(*args, return_storages=True, slab_steps=97)
    for k in one:
        assert_bits_equal(many[k], one[k], f"snow-ice slabs {k}")
    qobs = synthetic.qobs_like(g["qsim"][:, 2])
    r = engine.snowice_gr4j(*args, qobs=qobs, want_qsim=False, slab_steps=200)
    np.testing.assert_allclose(r["mse"], oracle.mse_columns(qobs, g["qsim"]), rtol=1e-8)


# ------------------------------------------------------------------ BASELINE configs 3 and 4 at (per-GPU) full size
def test_full_size_gr4j_device_resident_properties(capsys):
    """BASELINE config 3 at its REAL size when the device has the room: GR4J, 1 048 576 members x 14 610 steps, the
    122.6 GB discharge array device resident (cudaMemGetInfo >= 130 GB free; otherwise 2^18 members, and the test
    prints which ran).  Size-independent properties over the whole array computed slab-wise on the device (finite,
    non-negative, a checksum that is independent of the launch geometry), sampled columns against the oracle, and
    batch independence: the first / last 2^16 members re-run as their own ensembles give bit-identical columns."""
    import torch
    T = synthetic.T_DAILY_40Y
    dev = torch.device("cuda:0")
    free, total = torch.cuda.mem_get_info(dev)
    N = 1 << 20 if free >= 130e9 else 1 << 18
    with capsys.disabled():  # which size ran belongs in the log of a -q run
        print(f"\n[config 3] GR4J {N} members x {T} steps ({N * T * 8 / 1e9:.1f} GB discharge array), "
              f"{free / 1e9:.0f} of {total / 1e9:.0f} GB free on the device")
    f = synthetic.forcing(T)
    P = engine.pack_params(synthetic.random_params(GR4J(), N))
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=dev)
    prec, etp, Pd = t(f["prec"]), t(f["etp"]), t(P)
    x4 = float(P[:, 3].max())
    q = engine.gr4j(prec, etp, 0.6, 0.7, Pd, x4_max=x4)["qsim"]
    torch.cuda.synchronize()
    assert q.shape == (T, N)
    finite, nonneg, checksum = True, True, torch.zeros((), dtype=torch.float64, device=dev)
    for r0 in range(0, T, 512):      # slab-wise: no second 122 GB temporary
        blk = q[r0:r0 + 512]
        finite &= bool(torch.isfinite(blk).all())
        nonneg &= bool((blk >= 0).all())
        checksum += blk.sum(dtype=torch.float64)
    assert finite and nonneg
    idx = np.r_[0:32, N - 32:N, np.random.default_rng(1).integers(0, N, 64)]
    ref = oracle.gr4j(f["prec"], f["etp"], 0.6, 0.7, P[idx])
    assert_close(q[:, torch.as_tensor(idx, device=dev)].cpu().numpy(), ref, f"gr4j {N} members, sampled columns")
    # batch independence / launch-geometry independence
    w = 1 << 16
    for lo in (0, N - w):
        part = engine.gr4j(prec, etp, 0.6, 0.7, Pd[lo:lo + w].contiguous(), x4_max=x4, block=64)["qsim"]
        assert bool(torch.equal(part, q[:, lo:lo + w])), f"members [{lo}, {lo + w}) differ when run as their own ensemble"
        del part
    mass_in = float(np.sum(f["prec"]))
    mean_q = float(checksum) / N   # per-member discharge total: bounded by rain + the groundwater exchange
    assert 0.0 < mean_q < 3.0 * mass_in
    del q
    torch.cuda.empty_cache()


def test_full_size_cemaneigegr4j_per_gpu_share():
    """BASELINE config 4: 262 144 members over 8 GPUs = 32 768 per GPU, 40 years, 5 layers; host mode."""
    T, N = synthetic.T_DAILY_40Y, 32768
    f = synthetic.forcing(T)
    P = synthetic.random_params(CemaneigeGR4J(), N)
    q = CemaneigeGR4J().simulate(f["prec"], f["temp"], f["min_temp"], f["max_temp"], f["etp"],
                                 met_station_height=synthetic.MET_STATION_HEIGHT, altitudes=synthetic.ALTITUDES,
                                 s_init=0.6, r_init=0.7, params=P)
    assert q.shape == (T, N) and np.isfinite(q).all()
    idx = np.r_[0:16, N - 16:N, np.random.default_rng(2).integers(0, N, 32)]
    alts = synthetic.ALTITUDES
    p = oracle.extrapolate_precipitation(f["prec"], alts, synthetic.MET_STATION_HEIGHT)
    mn, me, mx = oracle.extrapolate_temperature(f["min_temp"], f["temp"], f["max_temp"], alts, synthetic.MET_STATION_HEIGHT)
    fr = oracle.calculate_solid_fraction(p, alts, me, mn, mx)
    ref = oracle.cemaneigegr4j(p, me, f["etp"], fr, (0, 0, 0.6, 0.7), P[idx])
    assert_close(q[:, idx], ref, "cemaneigegr4j 32k sampled columns")


def test_full_size_hbvedu_multi_catchment_per_gpu_share():
    """BASELINE config 5: 1024 catchments x 4096 members over 8 GPUs = 128 catchments per GPU, 10 years hourly
    (T = 87 660).  The discharge of one GPU's share would be 368 GB, so the run uses the fused objective
    (device mode, no qsim): one launch, grid.y = catchment.  Sampled members are checked against the oracle."""
    import torch
    Cn, N, T = 128, 4096, synthetic.T_HOURLY_10Y
    dev = torch.device("cuda:0")
    fs = [synthetic.forcing(T, seed=synthetic.SEED + c, hourly=True) for c in range(Cn)]
    temp = np.stack([f["temp"] for f in fs]); prec = np.stack([f["prec"] for f in fs])
    m0 = np.stack([f["month"] - 1 for f in fs]).astype(np.int8)
    PE = np.stack([f["PE_m"] / 24.0 for f in fs]); TM = np.stack([f["T_m"] for f in fs])
    P = engine.pack_params(synthetic.random_params(HBVEdu(), Cn * N, seed=77)).reshape(Cn, N, 11)
    qobs = np.abs(np.random.default_rng(11).normal(0.05, 0.02, (Cn, T)))
    t = lambda a, dt=torch.float64: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=dev)
    args = (t(temp), t(prec), t(m0, torch.int8), t(PE), t(TM), (0, 100, 3, 10), t(P))
    r = engine.hbvedu_multi(*args, qobs=t(qobs), want_qsim=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    r = engine.hbvedu_multi(*args, qobs=t(qobs), want_qsim=False)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"\nconfig 5 per-GPU share: {Cn} x {N} members x {T} steps in {ms:.1f} ms = {Cn * N * T / ms / 1e6:.1f} G member-steps/s")
    mse = r["mse"].cpu().numpy()
    assert mse.shape == (Cn, N) and np.isfinite(mse).all()
    rng = np.random.default_rng(4)
    for c in (0, 63, 127):
        idx = np.r_[0, N - 1, rng.integers(0, N, 6)]
        ref = oracle.hbvedu(temp[c], prec[c], m0[c], PE[c], TM[c], (0, 100, 3, 10), P[c, idx])
        ref_mse = ((qobs[c][:, None] - ref) ** 2).mean(axis=0)
        np.testing.assert_allclose(mse[c, idx], ref_mse, rtol=1e-9)


def test_snow_layers_on_the_device_are_bit_identical_to_the_host_preprocessing():
    """rrb_snow_layers = extrapolate_precipitation + extrapolate_temperature + calculate_solid_fraction on the GPU."""
    import torch
    from rrmpg_b200.models import _snow_inputs
    f = synthetic.forcing(2000)
    cases = [([550, 620, 700, 785, 920], 495), ([], 318), ([1400, 1500, 2200, 3999, 4000, 4500], 1350),
             ([4200, 5000], 4100), ([495.0], 495)]
    for alts, h in cases:
        ref = _snow_inputs.to_layers(f["prec"], f["temp"], f["min_temp"], f["max_temp"], h, np.array(alts, dtype=float))
        got = engine.snow_layers(f["prec"], f["temp"], f["min_temp"], f["max_temp"], h, alts)
        for a, b, nm in zip(got, ref[:3], ["layer_prec", "layer_mean_temp", "frac_solid"]):
            assert_bits_equal(a, b, f"snow_layers {alts} {nm}")
    dev = torch.device("cuda:0")
    t = lambda a: torch.as_tensor(a, dtype=torch.float64, device=dev)
    alts, h = cases[0]
    got = engine.snow_layers(t(f["prec"]), t(f["temp"]), t(f["min_temp"]), t(f["max_temp"]), h, alts)
    ref = _snow_inputs.to_layers(f["prec"], f["temp"], f["min_temp"], f["max_temp"], h, np.array(alts, dtype=float))
    for a, b in zip(got, ref[:3]):
        assert a.is_cuda
        assert_bits_equal(a.cpu().numpy(), b, "snow_layers device mode")
    # feeds the ensemble kernel without leaving the device
    P = synthetic.random_params(CemaneigeGR4J(), 64)
    q = engine.cemaneigegr4j(got[0], got[1], t(f["etp"]), got[2], (0, 0, 0.6, 0.7), t(engine.pack_params(P)))["qsim"]
    qr = oracle.cemaneigegr4j(ref[0], ref[1], f["etp"], ref[2], (0, 0, 0.6, 0.7), P)
    assert_close(q.cpu().numpy(), qr, "device preprocessing -> coupled kernel")


def test_gr4j_family_multi_catchment_equals_a_loop_over_catchments():
    """rrb_gr4j_simulate_multi / rrb_cemaneigegr4j_simulate_multi: grid.y = catchment, bit-identical to one call per
    catchment (host mode incl. the chunked D2H ring, device mode with the fused objective only)."""
    import torch
    from rrmpg_b200.models import _snow_inputs
    Cn, T, N = 4, 500, 90
    fs = [synthetic.forcing(T, seed=300 + c) for c in range(Cn)]
    prec = np.stack([f["prec"] for f in fs]); etp = np.stack([f["etp"] * (1 + 0.05 * c) for c, f in enumerate(fs)])
    qobs = np.abs(np.random.default_rng(5).normal(1.0, 0.5, (Cn, T)))
    # ---- GR4J
    P = np.stack([engine.pack_params(synthetic.random_params(GR4J(), N, seed=400 + c)) for c in range(Cn)])
    P[2, 5, 0] = 1e-6   # one catchment has a CTA outside the FAST contract: per-catchment fallback
    inits = np.array([[0.6, 0.7], [0.1, 0.9], [0.0, 0.0], [1.0, 0.3]])
    multi = engine.gr4j_multi(prec, etp, inits, P, return_storage=True, qobs=qobs)
    assert multi["qsim"].shape == (Cn, T, N) and multi["mse"].shape == (Cn, N)
    for c in range(Cn):
        one = engine.gr4j(prec[c], etp[c], inits[c, 0], inits[c, 1], P[c], return_storage=True, qobs=qobs[c])
        for nm in one:
            assert_bits_equal(multi[nm][c], one[nm], f"gr4j catchment {c} {nm}")
        assert_close(multi["qsim"][c], oracle.gr4j(prec[c], etp[c], inits[c, 0], inits[c, 1], P[c]), f"gr4j catchment {c}")
    dev = torch.device("cuda:0")
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=dev)
    r = engine.gr4j_multi(t(prec), t(etp), inits, t(P), qobs=t(qobs), want_qsim=False)
    torch.cuda.synchronize()
    assert set(r) == {"mse"}
    assert_bits_equal(r["mse"].cpu().numpy(), multi["mse"], "gr4j device-mode objective")
    # ---- CemaneigeGR4J, catchments at different station heights, one without rain -> G_tresh = 0
    alts = [550, 620, 700, 785, 920]
    lay = [_snow_inputs.to_layers(f["prec"] * (c != 1), f["temp"] - 2 * c, f["min_temp"] - 2 * c, f["max_temp"] - 2 * c,
                                  400 + 50 * c, np.array(alts, dtype=float)) for c, f in enumerate(fs)]
    lp = np.stack([l[0] for l in lay]); lt = np.stack([l[1] for l in lay]); fr = np.stack([l[2] for l in lay])
    PC = np.stack([engine.pack_params(synthetic.random_params(CemaneigeGR4J(), N, seed=500 + c)) for c in range(Cn)])
    ini4 = np.array([[0, 0, 0.6, 0.7], [5.0, -1.0, 0.5, 0.5], [0, 0, 0.2, 0.9], [20.0, 0.0, 0.6, 0.7]])
    multi = engine.cemaneigegr4j_multi(lp, lt, etp, fr, ini4, PC, return_storages=True, qobs=qobs)
    assert multi["G"].shape == (Cn, T, 5, N)
    for c in range(Cn):
        one = engine.cemaneigegr4j(lp[c], lt[c], etp[c], fr[c], ini4[c], PC[c], return_storages=True, qobs=qobs[c])
        for nm in one:
            assert_bits_equal(multi[nm][c], one[nm], f"cemaneigegr4j catchment {c} {nm}")
        ref = oracle.cemaneigegr4j(lp[c], lt[c], etp[c], fr[c], ini4[c], PC[c], return_storages=True)
        assert_bits_equal(multi["G"][c], ref[1], f"cemaneigegr4j catchment {c} G vs oracle")
        assert_close(multi["qsim"][c], ref[0], f"cemaneigegr4j catchment {c} qsim vs oracle")
    r = engine.cemaneigegr4j_multi(t(lp), t(lt), t(etp), t(fr), ini4, t(PC), qobs=t(qobs), want_qsim=False)
    torch.cuda.synchronize()
    assert_bits_equal(r["mse"].cpu().numpy(), multi["mse"], "cemaneigegr4j device-mode objective")


def test_randomized_parity_sweep_all_models():
    """Seeded random problem shapes (series length, ensemble size, launch geometry, initial states, parameter draws up
    to the edges of the default bounds) for every model, FAST math, against the oracle."""
    from rrmpg_b200.models import _snow_inputs
    rng = np.random.default_rng(20261017)
    worst = {}

    def note(name, got, ref):
        got, ref = np.asarray(got), np.asarray(ref)
        assert_close(got, ref, name)
        m = np.isfinite(ref) & (np.abs(ref) > 1e-6)
        if m.any():
            worst[name] = max(worst.get(name, 0.0), float(np.max(np.abs(got[m] - ref[m]) / np.abs(ref[m]))))

    def edge(P, frac=0.15):
        """push a fraction of the members onto the lower / upper bound of one random field each"""
        P = P.copy()
        names = P.dtype.names
        for i in rng.choice(P.shape[0], max(1, int(frac * P.shape[0])), replace=False):
            nm = names[rng.integers(len(names))]
            lo, hi = P[nm].min(), P[nm].max()
            P[nm][i] = lo if rng.random() < 0.5 else hi
        return P

    for trial in range(6):
        T = int(rng.integers(3, 900))
        N = int(rng.integers(1, 700))
        block = int(rng.choice([0, 32, 64, 128, 256]))
        f = synthetic.forcing(T, seed=int(rng.integers(1 << 30)))
        m0 = (f["month"] - 1).astype(np.int8)
        # HBVEdu
        P = edge(synthetic.random_params(HBVEdu(), N, seed=int(rng.integers(1 << 30))))
        ini = (float(rng.uniform(0, 50)), float(rng.uniform(20, 200)), float(rng.uniform(0, 20)), float(rng.uniform(0, 30)))
        got = engine.hbvedu(f["temp"], f["prec"], m0, f["PE_m"], f["T_m"], ini, P, return_storage=True, block=block)
        ref = oracle.hbvedu(f["temp"], f["prec"], m0, f["PE_m"], f["T_m"], ini, P, return_storage=True)
        for nm, r in zip(["qsim", "snow", "soil", "s1", "s2"], ref):
            note("hbvedu." + nm, got[nm], r)
        # ABC (bit-exact)
        Pa = synthetic.random_params(ABCModel(), N, seed=int(rng.integers(1 << 30)))
        s0 = float(rng.uniform(0, 10))
        assert_bits_equal(engine.abc(f["prec"], s0, Pa, block=block)["qsim"], oracle.abc(f["prec"], s0, Pa), "abc")
        # GR4J
        Pg = edge(synthetic.random_params(GR4J(), N, seed=int(rng.integers(1 << 30))))
        si, ri = float(rng.uniform(0, 1)), float(rng.uniform(0, 1))
        got = engine.gr4j(f["prec"], f["etp"], si, ri, Pg, return_storage=True, block=block)
        ref = oracle.gr4j(f["prec"], f["etp"], si, ri, Pg, return_storage=True)
        for nm, r in zip(["qsim", "s_store", "r_store"], ref):
            note("gr4j." + nm, got[nm], r)
        # Cemaneige (bit-exact) and CemaneigeGR4J, random number of layers
        L = int(rng.integers(1, 8))
        alts = sorted(float(a) for a in rng.uniform(300, 2500, L))
        lp, lt, fr, _ = _snow_inputs.to_layers(f["prec"], f["temp"], f["min_temp"], f["max_temp"], 480.0, np.array(alts))
        Pc = edge(synthetic.random_params(Cemaneige(), N, seed=int(rng.integers(1 << 30))))
        g0, e0 = float(rng.uniform(0, 40)), float(-rng.uniform(0, 3))
        got = engine.cemaneige(lp, lt, fr, g0, e0, Pc, return_storages=True, block=block)
        ref = oracle.cemaneige(lp, lt, fr, g0, e0, Pc, return_storages=True)
        for nm, r in zip(["outflow", "G", "eTG"], ref):
            assert_bits_equal(got[nm], r, f"cemaneige L={L} {nm}")
        Pcg = edge(synthetic.random_params(CemaneigeGR4J(), N, seed=int(rng.integers(1 << 30))))
        ini4 = (g0, e0, si, ri)
        got = engine.cemaneigegr4j(lp, lt, f["etp"], fr, ini4, Pcg, return_storages=True, block=block)
        ref = oracle.cemaneigegr4j(lp, lt, f["etp"], fr, ini4, Pcg, return_storages=True)
        assert_bits_equal(got["G"], ref[1], f"cemaneigegr4j L={L} G")
        assert_bits_equal(got["eTG"], ref[2], f"cemaneigegr4j L={L} eTG")
        for nm, r in zip(["qsim", "s_store", "r_store"], [ref[0], ref[3], ref[4]]):
            note("cemaneigegr4j." + nm, got[nm], r)
    print("\nworst relative deviation from the oracle (|ref| > 1e-6), FAST math: "
          + ", ".join(f"{k} {v:.1e}" for k, v in sorted(worst.items())))
    assert max(worst.values()) < 1e-10


@pytest.mark.parametrize("N", [2, 64, 1000, 1001])
def test_abc_pair_kernel_and_scalar_kernel_are_bit_identical(N):
    """Even ensembles take the two-members-per-thread kernel with 16-byte stores, odd ones (and any call with
    storages / an objective) the scalar kernel; both bit-identical to the oracle, also across time slabs."""
    f = synthetic.forcing(700, seed=21)
    P = synthetic.random_params(ABCModel(), N, seed=22)
    ref = oracle.abc(f["prec"], 1.5, P, return_storage=True)
    one = engine.abc(f["prec"], 1.5, P)["qsim"]
    assert_bits_equal(one, ref[0], f"abc N={N}")
    for slab in (1, 7, 128):
        assert_bits_equal(engine.abc(f["prec"], 1.5, P, slab_steps=slab)["qsim"], ref[0], f"abc N={N} slab={slab}")
    both = engine.abc(f["prec"], 1.5, P, return_storage=True)
    assert_bits_equal(both["qsim"], ref[0], "abc qsim with storage")
    assert_bits_equal(both["storage"], ref[1], "abc storage")


def test_multi_catchment_host_ring_with_many_small_chunks(monkeypatch):
    """Host-mode catchment batches go through a two-deep device ring in chunks of catchments (D2H of chunk k overlaps
    the kernel of chunk k+1).  The test knob RRMPG_B200_MULTI_CHUNK_BYTES forces one- and two-catchment chunks; results
    must not depend on the chunking."""
    Cn, T, N = 7, 300, 130
    fs = [synthetic.forcing(T, seed=700 + c) for c in range(Cn)]
    prec = np.stack([f["prec"] for f in fs]); etp = np.stack([f["etp"] for f in fs]); temp = np.stack([f["temp"] for f in fs])
    m0 = np.stack([f["month"] - 1 for f in fs]).astype(np.int8)
    PE = np.stack([f["PE_m"] for f in fs]); TM = np.stack([f["T_m"] for f in fs])
    qobs = np.abs(np.random.default_rng(8).normal(1.0, 0.5, (Cn, T)))
    Pg = np.stack([engine.pack_params(synthetic.random_params(GR4J(), N, seed=800 + c)) for c in range(Cn)])
    Ph = np.stack([engine.pack_params(synthetic.random_params(HBVEdu(), N, seed=900 + c)) for c in range(Cn)])
    run_g = lambda: engine.gr4j_multi(prec, etp, (0.6, 0.7), Pg, return_storage=True, qobs=qobs)
    run_h = lambda: engine.hbvedu_multi(temp, prec, m0, PE, TM, (0, 100, 3, 10), Ph, return_storage=True, qobs=qobs)
    monkeypatch.delenv("RRMPG_B200_MULTI_CHUNK_BYTES", raising=False)
    whole_g, whole_h = run_g(), run_h()
    # a batch of ONE catchment takes its initial states from the per-catchment array too
    one = engine.gr4j_multi(prec[:1], etp[:1], (0.6, 0.7), Pg[:1])["qsim"][0]
    assert_bits_equal(one, engine.gr4j(prec[0], etp[0], 0.6, 0.7, Pg[0])["qsim"], "gr4j_multi C=1")
    one = engine.hbvedu_multi(temp[:1], prec[:1], m0[:1], PE[:1], TM[:1], (0, 100, 3, 10), Ph[:1])["qsim"][0]
    assert_bits_equal(one, engine.hbvedu(temp[0], prec[0], m0[0], PE[0], TM[0], (0, 100, 3, 10), Ph[0])["qsim"], "hbvedu_multi C=1")
    per_catchment = T * N * 8
    for chunk in (per_catchment * 3, per_catchment * 5 * 2 + 1):   # GR4J: 3 outputs -> 1 catchment; HBV: 5 outputs -> 2
        monkeypatch.setenv("RRMPG_B200_MULTI_CHUNK_BYTES", str(chunk))
        for whole, run, tag in ((whole_g, run_g, "gr4j"), (whole_h, run_h, "hbvedu")):
            got = run()
            for nm in whole:
                assert_bits_equal(got[nm], whole[nm], f"{tag} chunk={chunk} {nm}")


# ------------------------------------------------------------------ round 2: HBV-Edu FAST kernel variants
@pytest.fixture
def hbv_variant():
    """Sets rrb_opts.variant for the calls of one test (1 / 2: hbv_fast2_kernel with one / two members per thread,
    3: hbv_rot_kernel, 5: two members per thread as one CTA per SM; include/rrmpg_b200.h) and restores the library
    default afterwards."""
    def set_variant(v):
        engine.VARIANT = v
    yield set_variant
    engine.VARIANT = 0


@pytest.mark.parametrize("variant", [1, 2, 3, 5])
@pytest.mark.parametrize("N", [70, 71, 1000])
def test_hbvedu_fast_kernel_variants_vs_oracle(variant, N, hbv_variant):
    hbv_variant(variant)
    T = 700
    f, P = _hbv_case(T, N, seed=21)
    qobs = np.abs(np.random.default_rng(5).normal(2.0, 1.0, T))
    args = (f["temp"], f["prec"], f["month"] - 1, f["PE_m"], f["T_m"], (2.0, 100, 3, 10), P)
    ref = oracle.hbvedu(*args, return_storage=True)
    got = engine.hbvedu(*args, return_storage=True, qobs=qobs)
    for nm, r in zip(["qsim", "snow", "soil", "s1", "s2"], ref):
        assert_close(got[nm], r, f"variant {variant} N={N} {nm}")
    assert_close(got["mse"], np.mean((qobs[:, None] - ref[0]) ** 2, axis=0), f"variant {variant} mse", rtol=1e-9)
    q_only = engine.hbvedu(*args)["qsim"]
    assert_bits_equal(q_only, got["qsim"], f"variant {variant}: discharge-only launch vs all outputs")
    sl = engine.hbvedu(*args, return_storage=True, qobs=qobs, slab_steps=97)
    for nm in got:
        assert_bits_equal(sl[nm], got[nm], f"variant {variant}: time slabs {nm}")


@pytest.mark.parametrize("variant", [1, 2])
def test_hbvedu_members_outside_the_fast_contract_fall_back_per_cta(variant, hbv_variant):
    """hbv_fast2_kernel has no range check and no slow path in its time loop: members outside its contract (decided
    before the loop) and members whose soil moisture leaves the range of the table-driven pow (sticky maximum, judged
    after the loop) flag their CTA, whose members the PRECISE kernel queued behind then simulates -- slab by slab,
    with the carry state handed back and forth between the two kernels."""
    hbv_variant(variant)
    T, N = 600, 1000
    f, P = _hbv_case(T, N, seed=33)
    P["FC"][5] = 1e-200          # outside [2^-500, 2^500]
    P["Beta"][70] = 40.0         # |Beta| >= 32
    P["PWP"][200] = 0.0
    P["K_0"][333] = np.inf
    P["FC"][450] = 3.5e6         # soil/FC below 2^-15 from the first step on
    P["FC"][777] = 1.0e6         # ... and here only after a dry spell (range exit in the middle of the series)
    P["FC"][999] = -150.0        # negative base of the pow: NaN from the first wet step on
    for prec in (f["prec"], np.zeros(T)):
        args = (f["temp"], prec, f["month"] - 1, f["PE_m"], f["T_m"], (0.0, 100, 3, 10), P)
        with np.errstate(all="ignore"):
            ref = oracle.hbvedu(*args, return_storage=True)
        for slab in (0, 128):
            got = engine.hbvedu(*args, return_storage=True, slab_steps=slab, block=64)
            for nm, r in zip(["qsim", "snow", "soil", "s1", "s2"], ref):
                assert_close(got[nm], r, f"variant {variant} slab={slab} {nm}")


def test_hbvedu_negative_zero_temperature_with_a_zero_threshold():
    """the FAST kernels read temp < T_t off the sign of temp - T_t; -0.0 - (+0.0) = -0.0 made a day with
    temp = -0.0 (np.round(-0.04, 1)) 'cold' for T_t = 0, where the reference's -0.0 < 0.0 is False."""
    T, N = 400, 96
    f, P = _hbv_case(T, N, seed=4)
    temp = f["temp"].copy()
    temp[::3] = -0.0
    assert np.signbit(temp[0]) and temp[0] == 0.0
    P["T_t"][::2] = 0.0
    P["T_t"][1::4] = -0.0
    args = (temp, f["prec"], f["month"] - 1, f["PE_m"], f["T_m"], (5.0, 100, 3, 10), P)
    ref = oracle.hbvedu(*args, return_storage=True)
    for v in (1, 2, 3, 5):
        engine.VARIANT = v
        try:
            got = engine.hbvedu(*args, return_storage=True)
        finally:
            engine.VARIANT = 0
        for nm, r in zip(["qsim", "snow", "soil", "s1", "s2"], ref):
            assert_close(got[nm], r, f"variant {v} {nm}")


# ------------------------------------------------------------------ round 2: launch layouts of the HBV-Edu FAST kernel
@pytest.mark.parametrize("sms,pairs", [(4, 7), (4, 6), (4, 5), (3, 3), (2, 8), (2, 1), (4, 11), (2, 14), (2, 16), (3, 13)])
def test_hbvedu_rotating_schedule_and_one_cta_layout_are_bit_identical(sms, pairs, hbv_variant, monkeypatch):
    """hbv_rot_kernel (variant 3: one persistent CTA per SM, the member pairs rotate over its warps, fast warps stop on a
    shared-memory counter, per-warp TMA rings), the one-CTA-per-SM launches (variant 5: two members per thread, capped and
    uncapped build; variant 0: what the library picks for the size, one member per thread up to 16 warps per SM) and one
    member per thread (variant 1) run the arithmetic of variant 2 member for member: discharge, fused objectives and
    carried time-slab states must be bit-identical.  RRMPG_B200_HBV_ROT_SMS pretends a GPU of `sms` SMs so that a small ensemble reaches every shape of the
    schedule (`pairs` warps per CTA: slow / fast warps, 8- and 16-warp instantiations, a ragged last pair)."""
    monkeypatch.setenv("RRMPG_B200_HBV_ROT_SMS", str(sms))
    N, T = 64 * pairs * sms - 6, 2301
    f, P = _hbv_case(T, N, seed=sms * 100 + pairs)
    qobs = np.abs(np.random.default_rng(5).normal(2.0, 1.0, T))
    args = (f["temp"], f["prec"], f["month"] - 1, f["PE_m"], f["T_m"], (2.0, 100, 3, 10), P)
    ref = oracle.hbvedu(*args)
    for kw in (dict(), dict(qobs=qobs, objective="kge"), dict(qobs=qobs, want_qsim=False),
               dict(qobs=qobs, objective="nse", slab_steps=777)):
        hbv_variant(2)
        base = engine.hbvedu(*args, **kw)
        if base.get("qsim") is not None:
            assert_close(base["qsim"], ref, f"variant 2 {sorted(kw)} vs oracle")
        for v in (0, 1, 3, 5):   # 0 = the library's own choice of members per thread and launch shape for this size
            hbv_variant(v)
            got = engine.hbvedu(*args, **kw)
            for nm in base:
                if base[nm] is not None:
                    assert_bits_equal(got[nm], base[nm], f"variant {v} sms={sms} pairs={pairs} {sorted(kw)} {nm}")


@pytest.mark.parametrize("variant", [0, 3, 5])
def test_hbvedu_layout_variants_leave_flagged_members_to_the_precise_kernel(variant, hbv_variant, monkeypatch):
    """Members outside the FAST contract and soil moistures that leave the table range: hbv_rot_kernel flags the pair (64
    members) and takes it out of its schedule, the one-CTA launch flags its CTA; the PRECISE kernel behind (64-thread
    CTAs, a whole fraction of a flag word's members) redoes them -- also slab by slab with carried states."""
    monkeypatch.setenv("RRMPG_B200_HBV_ROT_SMS", "3")
    hbv_variant(variant)
    T, N = 2200, 64 * 7 * 3
    f, P = _hbv_case(T, N, seed=35)
    P["FC"][5] = 1e-200          # outside [2^-500, 2^500]
    P["Beta"][70] = 40.0         # |Beta| >= 32
    P["K_0"][333] = np.inf
    P["FC"][450] = 3.5e6         # soil/FC below 2^-15 from the first step on
    P["FC"][777] = 1.0e6         # ... and here only after a dry spell (range exit in the middle of the series)
    P["FC"][N - 1] = -150.0      # negative base of the pow: NaN from the first wet step on
    qobs = np.abs(np.random.default_rng(6).normal(2.0, 1.0, T))
    args = (f["temp"], f["prec"], f["month"] - 1, f["PE_m"], f["T_m"], (0.0, 100, 3, 10), P)
    with np.errstate(all="ignore"):
        ref = oracle.hbvedu(*args)
        mse = np.mean((qobs[:, None] - ref) ** 2, axis=0)
    for slab in (0, 700):
        got = engine.hbvedu(*args, qobs=qobs, slab_steps=slab)
        assert_close(got["qsim"], ref, f"variant {variant} slab={slab} qsim")
        assert_close(got["mse"], mse, f"variant {variant} slab={slab} mse", rtol=1e-9)
