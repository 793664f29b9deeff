"""The CPU oracle against the committed golden vectors (made by the live numba reference) and the
reference's own known-answer fixtures.  This is what pins the oracle on machines without /root/reference."""
import numpy as np

import oracle
from conftest import assert_bits_equal, load_golden


def test_fixture_hbvedu_matlab():
    g = load_golden("fixture_hbvedu")
    out = oracle.hbvedu(g["temp"], g["prec"], g["month"] - 1, g["PE_m"], g["T_m"], g["inits"], g["params"],
                        return_storage=True)
    # the reference's own acceptance criterion (test/test_models.py:172-174)
    assert np.allclose((out[0] * 410 * 1000 / (24 * 60 * 60)).flatten(), g["expected"])
    for nm, a in zip(["qsim", "snow", "soil", "s1", "s2"], out):
        assert_bits_equal(a, g[nm], "fixture_hbvedu." + nm)


def test_fixture_gr4j_excel():
    g = load_golden("fixture_gr4j")
    out = oracle.gr4j(g["prec"], g["etp"], 0.6, 0.7, g["params"], return_storage=True)
    assert np.allclose(out[0].flatten(), g["expected"])  # test/test_models.py:201-210
    for nm, a in zip(["qsim", "s_store", "r_store"], out):
        assert_bits_equal(a, g[nm], "fixture_gr4j." + nm)


def test_fixture_cemaneige_excel_including_preprocessing():
    g = load_golden("fixture_cemaneige")
    alt, h = g["altitudes"], float(g["met_station_height"])
    p = oracle.extrapolate_precipitation(g["prec"], alt, h)
    mn, me, mx = oracle.extrapolate_temperature(g["min_temp"], g["mean_temp"], g["max_temp"], alt, h)
    fr = oracle.calculate_solid_fraction(p, alt, me, mn, mx)
    assert_bits_equal(p, g["layer_prec"], "layer_prec")
    assert_bits_equal(me, g["layer_mean_temp"], "layer_mean_temp")
    assert_bits_equal(mn, g["layer_min_temp"], "layer_min_temp")
    assert_bits_equal(mx, g["layer_max_temp"], "layer_max_temp")
    assert_bits_equal(fr, g["frac_solid"], "frac_solid")
    out = oracle.cemaneige(p, me, fr, 0.0, 0.0, g["params"], return_storages=True)
    assert np.allclose(out[0].flatten(), g["expected"])  # test/test_models.py:227-236
    for nm, a in zip(["outflow", "G", "eTG"], out):
        assert_bits_equal(a, g[nm], "fixture_cemaneige." + nm)


def test_fixture_cemaneigegr4j_excel():
    g = load_golden("fixture_cemaneigegr4j")
    alt, h = g["altitudes"], float(g["met_station_height"])
    p = oracle.extrapolate_precipitation(g["prec"], alt, h)
    mn, me, mx = oracle.extrapolate_temperature(g["min_temp"], g["mean_temp"], g["max_temp"], alt, h)
    fr = oracle.calculate_solid_fraction(p, alt, me, mn, mx)
    out = oracle.cemaneigegr4j(p, me, g["etp"], fr, g["inits"], g["params"], return_storages=True)
    assert np.allclose(out[0].flatten(), g["expected"])  # test/test_models.py:258-268
    for nm, a in zip(["qsim", "G", "eTG", "s_store", "r_store"], out):
        assert_bits_equal(a, g[nm], "fixture_cemaneigegr4j." + nm)


def test_ensembles_bit_exact_against_numba():
    g = load_golden("ensemble_abc")
    q, s = oracle.abc(g["prec"], float(g["initial_state"]), g["params"], return_storage=True)
    assert_bits_equal(q, g["qsim"]); assert_bits_equal(s, g["storage"])

    g = load_golden("ensemble_hbvedu")
    out = oracle.hbvedu(g["temp"], g["prec"], g["month"] - 1, g["PE_m"], g["T_m"], g["inits"], g["params"],
                        return_storage=True)
    for nm, a in zip(["qsim", "snow", "soil", "s1", "s2"], out):
        assert_bits_equal(a, g[nm], "ensemble_hbvedu." + nm)
    out = oracle.hbvedu(g["temp"] - 6, g["prec"], g["month"] - 1, g["PE_m"], g["T_m"], g["inits_cold"], g["params"])
    assert_bits_equal(out, g["qsim_cold"], "ensemble_hbvedu.cold")

    g = load_golden("ensemble_gr4j")
    out = oracle.gr4j(g["prec"], g["etp"], 0.6, 0.7, g["params"], return_storage=True)
    for nm, a in zip(["qsim", "s_store", "r_store"], out):
        assert_bits_equal(a, g[nm], "ensemble_gr4j." + nm)
    out = oracle.gr4j(g["prec"], g["etp"], 0.3, 0.5, g["params_longuh"], return_storage=True)
    for nm, a in zip(["qsim", "s_store", "r_store"], out):
        assert_bits_equal(a, g[nm + "_longuh"], "ensemble_gr4j.longuh." + nm)

    c = load_golden("ensemble_cemaneige")
    out = oracle.cemaneige(c["layer_prec"], c["layer_mean_temp"], c["frac_solid"], 0.0, 0.0, c["params"],
                           return_storages=True)
    for nm, a in zip(["outflow", "G", "eTG"], out):
        assert_bits_equal(a, c[nm], "ensemble_cemaneige." + nm)
    out = oracle.cemaneige(c["prec"][:, None], c["mean_temp"][:, None], c["frac_solid_L1"], 12.0, -1.5,
                           c["params"], return_storages=True)
    for nm, a in zip(["outflow", "G", "eTG"], out):
        assert_bits_equal(a, c[nm + "_L1"], "ensemble_cemaneige.L1." + nm)

    g = load_golden("ensemble_cemaneigegr4j")
    out = oracle.cemaneigegr4j(c["layer_prec"], c["layer_mean_temp"], g["etp"], c["frac_solid"], g["inits"],
                               g["params"], return_storages=True)
    for nm, a in zip(["qsim", "G", "eTG", "s_store", "r_store"], out):
        assert_bits_equal(a, g[nm], "ensemble_cemaneigegr4j." + nm)


def test_thread_count_does_not_change_results_and_ragged_blocks():
    g = load_golden("ensemble_hbvedu")
    P = np.repeat(g["params"], 3, axis=0)[:77]  # 77 members: not a multiple of the 8-member blocks
    a = oracle.hbvedu(g["temp"], g["prec"], g["month"] - 1, g["PE_m"], g["T_m"], g["inits"], P, nthreads=1)
    b = oracle.hbvedu(g["temp"], g["prec"], g["month"] - 1, g["PE_m"], g["T_m"], g["inits"], P, nthreads=5)
    assert_bits_equal(a, b)
    assert_bits_equal(a[:, :28], g["qsim"][:, np.arange(77) // 3][:, :28])


def test_mse_columns_matches_calc_mse():
    g = load_golden("ensemble_hbvedu"); m = load_golden("ensemble_hbvedu_mse")
    np.testing.assert_allclose(oracle.mse_columns(m["qobs"], g["qsim"]), m["mse"], rtol=1e-12)


def test_empty_and_single_step_series():
    P = load_golden("ensemble_hbvedu")["params"][:3]
    q = oracle.hbvedu(np.zeros(1), np.zeros(1), np.zeros(1, np.int8), np.zeros(12), np.zeros(12), (0, 1, 2, 3), P)
    assert q.shape == (1, 3) and (q == 0).all()
    q = oracle.gr4j(np.zeros(0), np.zeros(0), 0.5, 0.5, np.array([[300.0, 1.0, 50.0, 1.5]]))
    assert q.shape == (0, 1)


SNOWICE = [("fixture_cemaneigehystgr4j", 1, 0), ("fixture_cemaneigehystgr4jice", 1, 1),
           ("ensemble_cemaneigegr4jice", 0, 1), ("ensemble_cemaneigehystgr4j", 1, 0),
           ("ensemble_cemaneigehystgr4jice", 1, 1), ("ensemble_cemaneigehystgr4jice_L1", 1, 1),
           ("ensemble_cemaneigehystgr4jice_T1", 1, 1)]
SI_NAMES = ["qsim", "G", "eTG", "s_store", "r_store", "sca", "icemelt", "snowmelt", "rain"]


def snowice_layers(g):
    alts = list(g["altitudes"]); h = float(g["met_station_height"])
    if alts:
        p = oracle.extrapolate_precipitation(g["prec"], alts, h)
        mn, me, mx = oracle.extrapolate_temperature(g["min_temp"], g["mean_temp"], g["max_temp"], alts, h)
    else:
        alts = [h]
        p, me, mn, mx = (g[k][:, None] for k in ("prec", "mean_temp", "min_temp", "max_temp"))
    return p, me, oracle.calculate_solid_fraction(p, alts, me, mn, mx)


def test_snow_ice_family_bit_exact_against_numba_and_fixtures():
    import pytest
    for name, hyst, ice in SNOWICE:
        g = load_golden(name)
        p, me, fr = snowice_layers(g)
        out = dict(zip(SI_NAMES, oracle.snowice_gr4j(hyst, ice, p, me, g["etp"], g["frac_ice"] if ice else None, fr,
                                                     g["inits"], g["params"], return_storages=True)))
        for k in SI_NAMES:
            if k in g:
                assert_bits_equal(out[k], g[k], f"{name}.{k}")
        if "expected" in g:  # test/test_models.py:270-356
            assert np.allclose(out["qsim"].flatten(), g["expected"])
