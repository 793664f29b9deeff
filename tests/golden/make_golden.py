#!/usr/bin/env python
"""Generate the committed golden vectors under tests/golden/ from the LIVE reference.

Run in the build container only (needs /root/reference + numba):

    python tests/golden/make_golden.py

What it does
  1. imports the unmodified reference (``rrmpg`` from /root/reference, numba JIT),
  2. replays the reference's four known-answer tests (``test/test_models.py:142-268``) and
     stores inputs, parameters, the fixture's expected column and the numba output,
  3. runs seeded synthetic ensembles of all five hot-path models through the numba kernels
     (``run_<model>`` per member, which side-steps the early ``return`` at
     ``rrmpg/models/gr4j.py:178``) and stores inputs + every output array,
  4. PINS the C oracle: asserts ``oracle/`` reproduces every numba array bit-for-bit
     (same glibc pow/tanh/exp, no contraction) and reports it.

The .npz files travel to the GPU box; /root/reference does not.
"""
import os
import sys

import numpy as np
import pandas as pd

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from rrmpg.models import (ABCModel, HBVEdu, GR4J, Cemaneige, CemaneigeGR4J, CemaneigeGR4JIce,  # noqa: E402
                          CemaneigeHystGR4J, CemaneigeHystGR4JIce)
from rrmpg.models.abcmodel_model import run_abcmodel  # noqa: E402
from rrmpg.models.hbvedu_model import run_hbvedu  # noqa: E402
from rrmpg.models.gr4j_model import run_gr4j  # noqa: E402
from rrmpg.models.cemaneige_model import run_cemaneige  # noqa: E402
from rrmpg.models.cemaneigegr4j_model import run_cemaneigegr4j  # noqa: E402
from rrmpg.models.cemaneige_utils import (calculate_solid_fraction,  # noqa: E402
                                          extrapolate_precipitation, extrapolate_temperature)
from rrmpg.utils.metrics import calc_mse  # noqa: E402

import oracle  # noqa: E402
from rrmpg_b200 import synthetic  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
DATA = "/root/reference/test/data"
REPORT = []


def bits_equal(a, b):
    a = np.ascontiguousarray(a, np.float64); b = np.ascontiguousarray(b, np.float64)
    return a.shape == b.shape and np.array_equal(a.view(np.uint64), b.view(np.uint64))


def pin(name, ref, got):
    ok = bits_equal(ref, got)
    err = float(np.nanmax(np.abs(np.asarray(ref) - np.asarray(got)))) if not ok else 0.0
    REPORT.append((name, ok, err))
    print(f"  oracle==numba  {name:45s} {'BIT-EXACT' if ok else 'MISMATCH max_abs=%g' % err}")
    assert ok, name


def pack(params):
    return oracle.pack_params(params)


def save(name, **arrays):
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"wrote {path} ({os.path.getsize(path) / 1e3:.0f} kB)")


# --------------------------------------------------------------------------------------
# numba ensemble runners (one run_<model> call per member, like the wrappers' loops)
# --------------------------------------------------------------------------------------
def nb_abc(prec, s0, params):
    T, N = prec.size, params.size
    q = np.zeros((T, N)); s = np.zeros((T, N))
    for i in range(N):
        q[:, i], s[:, i] = run_abcmodel(prec, float(s0), params[i])
    return q, s


def nb_hbv(temp, prec, month0, PE_m, T_m, inits, params):
    T, N = prec.size, params.size
    o = [np.zeros((T, N)) for _ in range(5)]
    for i in range(N):
        r = run_hbvedu(temp, prec, month0, PE_m, T_m, *[float(v) for v in inits], params[i])
        for k in range(5):
            o[k][:, i] = r[k]
    return o


def nb_gr4j(prec, etp, s_init, r_init, params):
    T, N = prec.size, params.size
    o = [np.zeros((T, N)) for _ in range(3)]
    for i in range(N):
        r = run_gr4j(prec, etp, float(s_init), float(r_init), params[i])
        for k in range(3):
            o[k][:, i] = r[k]
    return o


def nb_cema(prec, mean_temp, frac, g0, e0, params):
    (T, L), N = prec.shape, params.size
    out = np.zeros((T, N)); G = np.zeros((T, L, N)); E = np.zeros((T, L, N))
    for i in range(N):
        out[:, i], G[:, :, i], E[:, :, i] = run_cemaneige(prec, mean_temp, frac, float(g0),
                                                          float(e0), params[i])
    return out, G, E


def nb_cg(prec, mean_temp, etp, frac, inits, params):
    (T, L), N = prec.shape, params.size
    q = np.zeros((T, N)); G = np.zeros((T, L, N)); E = np.zeros((T, L, N))
    s = np.zeros((T, N)); r = np.zeros((T, N))
    for i in range(N):
        (q[:, i], G[:, :, i], E[:, :, i], s[:, i], r[:, i]) = run_cemaneigegr4j(
            prec, mean_temp, etp, frac, *[float(v) for v in inits], params[i])
    return q, G, E, s, r


def prep_layers(prec, mean_temp, min_temp, max_temp, station, altitudes):
    """The [T]->[T,L] preprocessing of rrmpg/models/cemaneige.py:198-219 (reference code)."""
    if len(altitudes) > 0:
        alt = np.array(altitudes)
        p = extrapolate_precipitation(prec, alt, station)
        mn, me, mx = extrapolate_temperature(min_temp, mean_temp, max_temp, alt, station)
    else:
        alt = np.array([station])
        p, me, mn, mx = (np.expand_dims(a, -1) for a in (prec, mean_temp, min_temp, max_temp))
    frac = calculate_solid_fraction(p, alt, me, mn, mx)
    return p, me, mn, mx, frac, alt


def pin_prep(tag, prec, mean_temp, min_temp, max_temp, station, altitudes, p, me, mn, mx, frac):
    if len(altitudes) > 0:
        pin(tag + ".extrapolate_precipitation", p,
            oracle.extrapolate_precipitation(prec, altitudes, station))
        o = oracle.extrapolate_temperature(min_temp, mean_temp, max_temp, altitudes, station)
        pin(tag + ".extrapolate_temperature.min", mn, o[0])
        pin(tag + ".extrapolate_temperature.mean", me, o[1])
        pin(tag + ".extrapolate_temperature.max", mx, o[2])
        alt = np.array(altitudes, float)
    else:
        alt = np.array([station], float)
    pin(tag + ".calculate_solid_fraction", frac, oracle.calculate_solid_fraction(p, alt, me, mn, mx))


# --------------------------------------------------------------------------------------
# 1. the reference's own known-answer fixtures
# --------------------------------------------------------------------------------------
def fixture_hbvedu():
    daily = pd.read_csv(f"{DATA}/hbv_daily_inputs.txt", sep="\t", names=["date", "month", "temp", "prec"])
    monthly = pd.read_csv(f"{DATA}/hbv_monthly_inputs.txt", sep=" ", names=["temp", "not_needed", "evap"])
    expected = pd.read_csv(f"{DATA}/hbv_qsim.csv", header=None, names=["qsim"]).qsim.to_numpy()
    pdict = {'T_t': 0, 'DD': 4.25, 'FC': 177.1, 'Beta': 2.35, 'C': 0.02, 'PWP': 105.89,
             'K_0': 0.05, 'K_1': 0.03, 'K_2': 0.02, 'K_p': 0.05, 'L': 4.87}   # test_models.py:107-117
    model = HBVEdu(params=pdict)
    inits = (0, 100, 3, 10)                                                # test_models.py:155-158
    out = model.simulate(temp=daily.temp, prec=daily.prec, month=daily.month, PE_m=monthly.evap,
                         T_m=monthly.temp, snow_init=0, soil_init=100, s1_init=3, s2_init=10,
                         return_storage=True)
    scale = 410 * 1000 / (24 * 60 * 60)                                    # test_models.py:172
    assert np.allclose((out[0] * 410 * 1000 / (24 * 60 * 60)).flatten(), expected)
    P = np.array([[pdict[k] for k in model._param_list]], float)
    got = oracle.hbvedu(daily.temp, daily.prec, daily.month.to_numpy() - 1, monthly.evap,
                        monthly.temp, inits, P, return_storage=True)
    for nm, a, b in zip(["qsim", "snow", "soil", "s1", "s2"], out, got):
        pin("fixture_hbvedu." + nm, a, b)
    save("fixture_hbvedu", temp=daily.temp.to_numpy(float), prec=daily.prec.to_numpy(float),
         month=daily.month.to_numpy(np.int8), PE_m=monthly.evap.to_numpy(float),
         T_m=monthly.temp.to_numpy(float), params=P, inits=np.array(inits, float),
         expected=expected, expected_scale=np.array(scale), qsim=out[0], snow=out[1],
         soil=out[2], s1=out[3], s2=out[4])


def fixture_gr4j():
    data = pd.read_csv(f"{DATA}/gr4j_example_data.csv", sep=",")
    pdict = {'x1': np.exp(5.76865628090826), 'x2': np.sinh(1.61742503661094),
             'x3': np.exp(4.24316129943456), 'x4': np.exp(-0.117506799276908) + 0.5}  # :186-189
    model = GR4J(params=pdict)
    out = model.simulate(data.prec, data.etp, s_init=0.6, r_init=0.7, return_storage=True)
    assert np.allclose(out[0].flatten(), data.qsim_excel)
    P = np.array([[pdict[k] for k in model._param_list]], float)
    got = oracle.gr4j(data.prec, data.etp, 0.6, 0.7, P, return_storage=True)
    for nm, a, b in zip(["qsim", "s_store", "r_store"], out, got):
        pin("fixture_gr4j." + nm, a, b)
    save("fixture_gr4j", prec=data.prec.to_numpy(float), etp=data.etp.to_numpy(float), params=P,
         inits=np.array([0.6, 0.7]), expected=data.qsim_excel.to_numpy(float), qsim=out[0],
         s_store=out[1], r_store=out[2])


def fixture_cemaneige():
    df = pd.read_csv(f"{DATA}/cemaneige_validation_data.csv", sep=";")
    pdict = {'CTG': 0.25, 'Kf': 3.74}                                      # test_models.py:221
    model = Cemaneige(params=pdict)
    alts = [550, 620, 700, 785, 920]
    out = model.simulate(df.precipitation, df.mean_temp, df.min_temp, df.max_temp,
                         met_station_height=495, altitudes=alts, return_storages=True)
    assert np.allclose(out[0].flatten(), df.liquid_outflow.to_numpy())
    raw = [df[c].to_numpy(float) for c in ("precipitation", "mean_temp", "min_temp", "max_temp")]
    p, me, mn, mx, frac, _ = prep_layers(*raw, 495, alts)
    pin_prep("fixture_cemaneige", *raw, 495, alts, p, me, mn, mx, frac)
    P = np.array([[pdict[k] for k in model._param_list]], float)
    got = oracle.cemaneige(p, me, frac, 0.0, 0.0, P, return_storages=True)
    for nm, a, b in zip(["outflow", "G", "eTG"], out, got):
        pin("fixture_cemaneige." + nm, a, b)
    save("fixture_cemaneige", prec=raw[0], mean_temp=raw[1], min_temp=raw[2], max_temp=raw[3],
         met_station_height=np.array(495.0), altitudes=np.array(alts, float), params=P,
         inits=np.zeros(2), layer_prec=p, layer_mean_temp=me, layer_min_temp=mn, layer_max_temp=mx,
         frac_solid=frac, expected=df.liquid_outflow.to_numpy(float), outflow=out[0], G=out[1],
         eTG=out[2])


def fixture_cemaneigegr4j():
    df = pd.read_csv(f"{DATA}/cemaneigegr4j_validation_data.csv", sep=";", index_col=0)
    pdict = {'CTG': 0.25, 'Kf': 3.74, 'x1': np.exp(5.25483021675164),
             'x2': np.sinh(1.58209470624126), 'x3': np.exp(4.3853181982412),
             'x4': np.exp(0.954786342674327) + 0.5}                         # test_models.py:247-252
    model = CemaneigeGR4J(params=pdict)
    alts = [550, 620, 700, 785, 920]
    out = model.simulate(df.precipitation, df.mean_temp, df.min_temp, df.max_temp, df.pe,
                         met_station_height=495, altitudes=alts, s_init=0.6, r_init=0.7,
                         return_storages=True)
    assert np.allclose(out[0].flatten(), df.qsim.to_numpy())
    raw = [df[c].to_numpy(float) for c in ("precipitation", "mean_temp", "min_temp", "max_temp")]
    etp = df.pe.to_numpy(float)
    p, me, mn, mx, frac, _ = prep_layers(*raw, 495, alts)
    P = np.array([[pdict[k] for k in model._param_list]], float)
    inits = (0.0, 0.0, 0.6, 0.7)
    got = oracle.cemaneigegr4j(p, me, etp, frac, inits, P, return_storages=True)
    for nm, a, b in zip(["qsim", "G", "eTG", "s_store", "r_store"], out, got):
        pin("fixture_cemaneigegr4j." + nm, a, b)
    save("fixture_cemaneigegr4j", prec=raw[0], mean_temp=raw[1], min_temp=raw[2], max_temp=raw[3],
         etp=etp, met_station_height=np.array(495.0), altitudes=np.array(alts, float), params=P,
         inits=np.array(inits), expected=df.qsim.to_numpy(float), qsim=out[0], G=out[1],
         eTG=out[2], s_store=out[3], r_store=out[4])


# --------------------------------------------------------------------------------------
# 2. seeded synthetic ensembles through the numba kernels
# --------------------------------------------------------------------------------------
def with_edges(params, edges):
    """Append hand-picked edge-case members (dict field->value overrides of member 0)."""
    extra = np.repeat(params[:1], len(edges))
    for i, e in enumerate(edges):
        for k, v in e.items():
            extra[k][i] = v
    return np.concatenate([params, extra])


def ensembles():
    T = 1096
    f = synthetic.forcing(T)
    month0 = (f["month"] - 1).astype(np.int8)

    # ---- ABC
    m = ABCModel()
    P = synthetic.random_params(m, 24)
    P = with_edges(P, [dict(a=0.0, b=0.0, c=0.0), dict(a=1.0, b=0.0, c=1.0), dict(a=0.5, b=0.5, c=0.25)])
    q, s = nb_abc(f["prec"], 1.5, P)
    g = oracle.abc(f["prec"], 1.5, pack(P), return_storage=True)
    pin("ensemble_abc.qsim", q, g[0]); pin("ensemble_abc.storage", s, g[1])
    save("ensemble_abc", prec=f["prec"], initial_state=np.array(1.5), params=pack(P), qsim=q, storage=s)

    # ---- HBVEdu
    m = HBVEdu()
    P = synthetic.random_params(m, 24)
    P = with_edges(P, [dict(T_t=-1.0, Beta=1.0), dict(T_t=1.0, Beta=7.0, L=2.0), dict(PWP=180.0, FC=100.0),
                       dict(K_0=0.2, K_1=0.1, K_2=0.05, K_p=0.05, L=5.0)])
    inits = (0.0, 100.0, 3.0, 10.0)
    o = nb_hbv(f["temp"], f["prec"], month0, f["PE_m"], f["T_m"], inits, P)
    g = oracle.hbvedu(f["temp"], f["prec"], month0, f["PE_m"], f["T_m"], inits, pack(P), return_storage=True)
    names = ["qsim", "snow", "soil", "s1", "s2"]
    for nm, a, b in zip(names, o, g):
        pin("ensemble_hbvedu." + nm, a, b)
    # snow-heavy start: non-zero snow_init, cold bias
    inits2 = (35.0, 150.0, 0.0, 0.0)
    o2 = nb_hbv(f["temp"] - 6, f["prec"], month0, f["PE_m"], f["T_m"], inits2, P)
    g2 = oracle.hbvedu(f["temp"] - 6, f["prec"], month0, f["PE_m"], f["T_m"], inits2, pack(P), return_storage=True)
    for nm, a, b in zip(names, o2, g2):
        pin("ensemble_hbvedu.cold." + nm, a, b)
    save("ensemble_hbvedu", temp=f["temp"], prec=f["prec"], month=f["month"], PE_m=f["PE_m"],
         T_m=f["T_m"], inits=np.array(inits), params=pack(P), inits_cold=np.array(inits2),
         qsim_cold=o2[0], soil_cold=o2[2], **dict(zip(names, o)))

    # ---- GR4J  (edge members: x4 exactly integral, x4 at the bounds, negative/positive x2)
    m = GR4J()
    P = synthetic.random_params(m, 24)
    P = with_edges(P, [dict(x4=2.0), dict(x4=1.1, x2=-5.0), dict(x4=2.9, x2=3.0), dict(x4=1.0),
                       dict(x4=0.6), dict(x1=100.0, x3=20.0), dict(x1=1200.0, x3=300.0)])
    o = nb_gr4j(f["prec"], f["etp"], 0.6, 0.7, P)
    g = oracle.gr4j(f["prec"], f["etp"], 0.6, 0.7, pack(P), return_storage=True)
    names = ["qsim", "s_store", "r_store"]
    for nm, a, b in zip(names, o, g):
        pin("ensemble_gr4j." + nm, a, b)
    # large-x4 members (outside the default bounds, as in the CemaneigeGR4J fixture and the
    # Hyst family's x4 <= 10): exercises the long unit-hydrograph path
    Pl = with_edges(P[:1], [dict(x4=3.0981), dict(x4=4.5), dict(x4=7.25), dict(x4=10.0), dict(x4=15.5)])[1:]
    ol = nb_gr4j(f["prec"], f["etp"], 0.3, 0.5, Pl)
    gl = oracle.gr4j(f["prec"], f["etp"], 0.3, 0.5, pack(Pl), return_storage=True)
    for nm, a, b in zip(names, ol, gl):
        pin("ensemble_gr4j.longuh." + nm, a, b)
    save("ensemble_gr4j", prec=f["prec"], etp=f["etp"], inits=np.array([0.6, 0.7]), params=pack(P),
         params_longuh=pack(Pl), inits_longuh=np.array([0.3, 0.5]), qsim_longuh=ol[0],
         s_store_longuh=ol[1], r_store_longuh=ol[2], **dict(zip(names, o)))

    # ---- Cemaneige, 5 layers and single layer
    m = Cemaneige()
    P = synthetic.random_params(m, 16)
    P = with_edges(P, [dict(CTG=0.0, Kf=0.0), dict(CTG=1.0, Kf=10.0), dict(CTG=0.0, Kf=10.0)])
    raw = (f["prec"], f["temp"], f["min_temp"], f["max_temp"])
    p, me, mn, mx, frac, _ = prep_layers(*raw, synthetic.MET_STATION_HEIGHT, synthetic.ALTITUDES)
    pin_prep("ensemble_cemaneige", *raw, synthetic.MET_STATION_HEIGHT, synthetic.ALTITUDES, p, me, mn, mx, frac)
    o = nb_cema(p, me, frac, 0.0, 0.0, P)
    g = oracle.cemaneige(p, me, frac, 0.0, 0.0, pack(P), return_storages=True)
    for nm, a, b in zip(["outflow", "G", "eTG"], o, g):
        pin("ensemble_cemaneige." + nm, a, b)
    p1, me1, mn1, mx1, frac1, _ = prep_layers(*raw, synthetic.MET_STATION_HEIGHT, [])
    pin_prep("ensemble_cemaneige.L1", *raw, synthetic.MET_STATION_HEIGHT, [], p1, me1, mn1, mx1, frac1)
    o1 = nb_cema(p1, me1, frac1, 12.0, -1.5, P)
    g1 = oracle.cemaneige(p1, me1, frac1, 12.0, -1.5, pack(P), return_storages=True)
    for nm, a, b in zip(["outflow", "G", "eTG"], o1, g1):
        pin("ensemble_cemaneige.L1." + nm, a, b)
    # high-altitude bands (>1500 m solid-fraction branch, >4000 m precipitation branch)
    hi_alts = [1400, 1600, 2500, 4200]
    ph, meh, mnh, mxh, frach, _ = prep_layers(*raw, 1450, hi_alts)
    pin_prep("ensemble_cemaneige.high", *raw, 1450, hi_alts, ph, meh, mnh, mxh, frach)
    oh = nb_cema(ph, meh, frach, 0.0, 0.0, P)
    gh = oracle.cemaneige(ph, meh, frach, 0.0, 0.0, pack(P), return_storages=True)
    for nm, a, b in zip(["outflow", "G", "eTG"], oh, gh):
        pin("ensemble_cemaneige.high." + nm, a, b)
    save("ensemble_cemaneige", prec=raw[0], mean_temp=raw[1], min_temp=raw[2], max_temp=raw[3],
         met_station_height=np.array(float(synthetic.MET_STATION_HEIGHT)),
         altitudes=np.array(synthetic.ALTITUDES, float), params=pack(P), inits=np.zeros(2),
         layer_prec=p, layer_mean_temp=me, frac_solid=frac, outflow=o[0], G=o[1], eTG=o[2],
         inits_L1=np.array([12.0, -1.5]), frac_solid_L1=frac1, outflow_L1=o1[0], G_L1=o1[1], eTG_L1=o1[2],
         altitudes_high=np.array(hi_alts, float), station_high=np.array(1450.0),
         layer_prec_high=ph, layer_mean_temp_high=meh, frac_solid_high=frach, outflow_high=oh[0])

    # ---- CemaneigeGR4J (5 layers; includes an x4 outside the default bounds like the fixture)
    m = CemaneigeGR4J()
    P = synthetic.random_params(m, 16)
    P = with_edges(P, [dict(x4=3.0981), dict(x4=2.0, CTG=0.0), dict(CTG=1.0, Kf=0.0, x2=-5.0)])
    inits = (0.0, 0.0, 0.6, 0.7)
    o = nb_cg(p, me, f["etp"], frac, inits, P)
    g = oracle.cemaneigegr4j(p, me, f["etp"], frac, inits, pack(P), return_storages=True)
    names = ["qsim", "G", "eTG", "s_store", "r_store"]
    for nm, a, b in zip(names, o, g):
        pin("ensemble_cemaneigegr4j." + nm, a, b)
    o1 = nb_cg(p1, me1, f["etp"], frac1, (5.0, -0.5, 0.2, 0.9), P)
    g1 = oracle.cemaneigegr4j(p1, me1, f["etp"], frac1, (5.0, -0.5, 0.2, 0.9), pack(P), return_storages=True)
    for nm, a, b in zip(names, o1, g1):
        pin("ensemble_cemaneigegr4j.L1." + nm, a, b)
    save("ensemble_cemaneigegr4j", prec=raw[0], mean_temp=raw[1], min_temp=raw[2], max_temp=raw[3],
         etp=f["etp"], met_station_height=np.array(float(synthetic.MET_STATION_HEIGHT)),
         altitudes=np.array(synthetic.ALTITUDES, float), params=pack(P), inits=np.array(inits),
         inits_L1=np.array([5.0, -0.5, 0.2, 0.9]), qsim_L1=o1[0], s_store_L1=o1[3],
         **dict(zip(names, o)))

    # ---- monte_carlo 'mse' semantics (rrmpg/tools/monte_carlo.py:66-73) on the HBV ensemble
    d = np.load(os.path.join(OUT, "ensemble_hbvedu.npz"))
    qobs = synthetic.qobs_like(d["qsim"][:, 3])
    mse = np.array([calc_mse(qobs, d["qsim"][:, n]) for n in range(d["qsim"].shape[1])])
    assert np.allclose(mse, oracle.mse_columns(qobs, d["qsim"]), rtol=1e-12, atol=0)
    save("ensemble_hbvedu_mse", qobs=qobs, mse=mse)


# --------------------------------------------------------------------------------------
# 3. snow-ice family: CemaneigeGR4JIce, CemaneigeHystGR4J, CemaneigeHystGR4JIce
# --------------------------------------------------------------------------------------
SI_NAMES = ["qsim", "G", "eTG", "s_store", "r_store", "sca", "icemelt", "snowmelt", "rain"]


def si_reference(model, hyst, ice, kw, frac_ice):
    """Run the reference wrapper with return_storages=True and map its tuple onto SI_NAMES."""
    if ice:
        out = model.simulate(kw["prec"], kw["mean_temp"], kw["min_temp"], kw["max_temp"], kw["etp"], frac_ice,
                             **kw["rest"], return_storages=True)
    else:
        out = model.simulate(kw["prec"], kw["mean_temp"], kw["min_temp"], kw["max_temp"], kw["etp"],
                             **kw["rest"], return_storages=True)
    d = {}
    if hyst and ice:      # cemaneigehystgr4jice.py:304
        keys = ["qsim", "G", "eTG", "s_store", "r_store", "sca", "icemelt", "snowmelt", "rain"]
    elif hyst:            # cemaneigehystgr4j.py:288
        keys = ["qsim", "G", "eTG", "s_store", "r_store", "sca", "rain"]
    else:                 # cemaneigegr4jice.py:286
        keys = ["qsim", "G", "eTG", "s_store", "r_store", "icemelt"]
    for k, a in zip(keys, out):
        d[k] = a
    return d


def si_case(tag, model, hyst, ice, raw, etp, station, alts, frac_ice, inits5, params, extra=None):
    """inits5 = (snow_pack_init, thermal_state_init, sca_init, s_init, r_init)."""
    rest = dict(met_station_height=station, altitudes=list(alts), snow_pack_init=inits5[0],
                thermal_state_init=inits5[1], s_init=inits5[3], r_init=inits5[4], params=params)
    if hyst:
        rest["sca_init"] = inits5[2]
    ref = si_reference(model, hyst, ice, dict(prec=raw[0], mean_temp=raw[1], min_temp=raw[2], max_temp=raw[3],
                                              etp=etp, rest=rest), frac_ice)
    p, me, mn, mx, frac, _ = prep_layers(*raw, station, alts)
    got = dict(zip(SI_NAMES, oracle.snowice_gr4j(hyst, ice, p, me, etp, frac_ice, frac, inits5, pack(params),
                                                 return_storages=True)))
    for k, a in ref.items():
        b = got[k]
        if k == "rain":
            b = np.repeat(b[:, :, None], a.shape[2], axis=2)
        pin(f"{tag}.{k}", a, b)
    arrays = dict(prec=raw[0], mean_temp=raw[1], min_temp=raw[2], max_temp=raw[3], etp=etp,
                  met_station_height=np.array(float(station)), altitudes=np.array(alts, float),
                  frac_ice=np.array(frac_ice if frac_ice is not None else [], float), inits=np.array(inits5, float),
                  params=pack(params))
    for k, a in ref.items():
        if k == "rain":
            a = a[:, :, 0]
        arrays[k] = a
    if extra:
        arrays.update(extra)
    save(tag, **arrays)


def snow_ice_family():
    alts = [550, 620, 700, 785, 920]
    # ---- reference fixtures (test/test_models.py:270-356)
    df = pd.read_csv(f"{DATA}/cemaneigehystgr4j_validation_data.csv", index_col=0)
    pd_ = {"Thacc": 18.6, "Rsp": 0.22, "CTG": 0.78, "Kf": 4.02, "x1": 546, "x2": 0.53, "x3": 276, "x4": 1.32}
    m = CemaneigeHystGR4J(params=pd_)
    q = m.simulate(df.precipitation, df.mean_temp, df.min_temp, df.max_temp, df.pe, met_station_height=700,
                   altitudes=alts, s_init=0.5, r_init=0.4)
    assert np.allclose(q.flatten(), df.qsim.to_numpy())
    raw = [df[c].to_numpy(float) for c in ("precipitation", "mean_temp", "min_temp", "max_temp")]
    P = np.zeros(1, m.get_dtype())
    for k in m.get_parameter_names():
        P[k] = pd_[k]
    si_case("fixture_cemaneigehystgr4j", m, 1, 0, raw, df.pe.to_numpy(float), 700, alts, None,
            (0.0, 0.0, 0.0, 0.5, 0.4), P, extra=dict(expected=df.qsim.to_numpy(float)))

    df = pd.read_csv(f"{DATA}/cemaneigehystgr4jice_validation_data.csv", index_col=0)
    pd_ = dict(pd_, DDF=5)
    m = CemaneigeHystGR4JIce(params=pd_)
    frac_ice = np.array([0.02, 0.04, 0.25, 0.51, 0.71])
    q = m.simulate(df.precipitation, df.mean_temp, df.min_temp, df.max_temp, df.pe, frac_ice, met_station_height=700,
                   altitudes=alts, s_init=0.5, r_init=0.4, sca_init=0.2)
    assert np.allclose(q.flatten(), df.qsim.to_numpy())
    raw = [df[c].to_numpy(float) for c in ("precipitation", "mean_temp", "min_temp", "max_temp")]
    P = np.zeros(1, m.get_dtype())
    for k in m.get_parameter_names():
        P[k] = pd_[k]
    si_case("fixture_cemaneigehystgr4jice", m, 1, 1, raw, df.pe.to_numpy(float), 700, alts, frac_ice,
            (0.0, 0.0, 0.2, 0.5, 0.4), P, extra=dict(expected=df.qsim.to_numpy(float)))

    # ---- seeded ensembles on the synthetic catchment (shifted colder so that ice and hysteresis are active)
    T = 1096
    f = synthetic.forcing(T)
    raw = (f["prec"], f["temp"] - 3, f["min_temp"] - 3, f["max_temp"] - 3)
    fice = np.array([0.0, 0.05, 0.2, 0.5, 0.8])
    m = CemaneigeGR4JIce()
    P = with_edges(synthetic.random_params(m, 12), [dict(DDF=1.0, x4=2.0), dict(DDF=30.0, Kf=1.0)])
    si_case("ensemble_cemaneigegr4jice", m, 0, 1, raw, f["etp"], synthetic.MET_STATION_HEIGHT, alts, fice,
            (8.0, -0.5, 0.0, 0.6, 0.7), P)
    m = CemaneigeHystGR4J()
    P = with_edges(synthetic.random_params(m, 12), [dict(Thacc=1.0, Rsp=0.0), dict(Thacc=1000.0, Rsp=1.0, x4=10.0),
                                                    dict(x4=4.2, x1=10.0, x3=5000.0)])
    si_case("ensemble_cemaneigehystgr4j", m, 1, 0, raw, f["etp"], synthetic.MET_STATION_HEIGHT, alts, None,
            (5.0, 0.0, 0.3, 0.6, 0.7), P)
    m = CemaneigeHystGR4JIce()
    P = with_edges(synthetic.random_params(m, 12), [dict(DDF=0.0), dict(Thacc=5.0, Rsp=0.5, DDF=30.0, x4=6.5)])
    si_case("ensemble_cemaneigehystgr4jice", m, 1, 1, raw, f["etp"], synthetic.MET_STATION_HEIGHT, alts, fice,
            (0.0, 0.0, 0.0, 0.5, 0.4), P)
    # single layer (no altitudes) + one-step series (the sca[t-1] wrap-around quirk at T == 1)
    m = CemaneigeHystGR4JIce()
    P = synthetic.random_params(m, 6)
    si_case("ensemble_cemaneigehystgr4jice_L1", m, 1, 1, raw, f["etp"], 1800, [], np.array([0.4]),
            (3.0, -1.0, 0.7, 0.5, 0.4), P)
    raw1 = tuple(a[:1].copy() for a in raw)
    si_case("ensemble_cemaneigehystgr4jice_T1", m, 1, 1, raw1, f["etp"][:1].copy(), 1800, [], np.array([0.4]),
            (3.0, -1.0, 0.7, 0.5, 0.4), P)


if __name__ == "__main__":
    oracle.build(force=True)
    fixture_hbvedu(); fixture_gr4j(); fixture_cemaneige(); fixture_cemaneigegr4j()
    ensembles()
    snow_ice_family()
    n_ok = sum(ok for _, ok, _ in REPORT)
    print(f"\noracle pinned bit-for-bit against numba on {n_ok}/{len(REPORT)} arrays")
    with open(os.path.join(OUT, "PINNING.txt"), "w") as fh:
        import numba, scipy
        fh.write("oracle/rr_oracle.c vs live numba reference (/root/reference @ 7de78c2)\n")
        fh.write(f"numba {numba.__version__}, numpy {np.__version__}, python {sys.version.split()[0]}\n")
        for name, ok, err in REPORT:
            fh.write(f"{'BIT-EXACT' if ok else 'MISMATCH %g' % err:12s} {name}\n")
