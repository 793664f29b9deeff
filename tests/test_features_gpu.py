"""Round-2 features of the C ABI, on the GPU: fused NSE / KGE objectives, monte_carlo on the fused path, state_in /
state_out resume (incl. GR4J's UH1 / UH2 buffers), in-library multi-device member sharding."""
import numpy as np
import pytest
from scipy.stats import pearsonr

import oracle
from rrmpg_b200 import _lib, engine, synthetic
from rrmpg_b200.models import ABCModel, HBVEdu, GR4J, CemaneigeGR4J, CemaneigeHystGR4J
from rrmpg_b200.tools import monte_carlo
from conftest import assert_bits_equal, assert_close, load_golden

pytestmark = pytest.mark.gpu


# ---- the reference's metrics, restated (rrmpg/utils/metrics.py:29-78, 110-136, 139-188)
def ref_mse(obs, sim):
    return np.mean((obs - sim) ** 2)


def ref_nse(obs, sim):
    return 1 - np.sum((sim - obs) ** 2) / np.sum((obs - np.mean(obs)) ** 2)


def ref_kge(obs, sim):
    r = pearsonr(obs, sim)[0]
    alpha = np.std(sim) / np.std(obs)
    beta = np.mean(sim) / np.mean(obs)
    return 1 - np.sqrt((r - 1) ** 2 + (alpha - 1) ** 2 + (beta - 1) ** 2)


REF_METRIC = {"mse": ref_mse, "nse": ref_nse, "kge": ref_kge}
OBJ_RTOL = 1e-9  # stated tolerance of the fused objectives (sequential in-register sums vs numpy's pairwise ones)


def _columns(fn, obs, sim):
    return np.array([fn(obs, sim[:, i]) for i in range(sim.shape[1])])


def _hbv(T=900, N=130, seed=2):
    f = synthetic.forcing(T)
    P = synthetic.random_params(HBVEdu(), N, seed=synthetic.PARAM_SEED + seed)
    qobs = oracle.hbvedu(f["temp"], f["prec"], f["month"] - 1, f["PE_m"], f["T_m"], (0, 100, 3, 10), P[:1])[:, 0]
    qobs = qobs * (1 + 0.2 * np.sin(np.arange(T) / 17.0)) + 0.05
    return f, P, qobs


@pytest.mark.parametrize("objective", ["mse", "nse", "kge"])
@pytest.mark.parametrize("variant", [1, 2])
def test_fused_objectives_hbvedu(objective, variant):
    f, P, qobs = _hbv()
    args = (f["temp"], f["prec"], f["month"] - 1, f["PE_m"], f["T_m"], (0, 100, 3, 10), P)
    ref = _columns(REF_METRIC[objective], qobs, oracle.hbvedu(*args))
    engine.VARIANT = variant
    try:
        got = engine.hbvedu(*args, qobs=qobs, want_qsim=False, objective=objective)
        assert "qsim" not in got
        np.testing.assert_allclose(got["mse"], ref, rtol=OBJ_RTOL, atol=1e-12)
        sl = engine.hbvedu(*args, qobs=qobs, objective=objective, slab_steps=111)  # sums carried across time slabs
        np.testing.assert_allclose(sl["mse"], ref, rtol=OBJ_RTOL, atol=1e-12)
        pr = engine.hbvedu(*args, qobs=qobs, want_qsim=False, objective=objective, math="precise")
        np.testing.assert_allclose(pr["mse"], ref, rtol=OBJ_RTOL, atol=1e-12)
    finally:
        engine.VARIANT = 0


@pytest.mark.parametrize("objective", ["nse", "kge"])
def test_fused_objectives_other_models(objective):
    T, N = 800, 70
    f = synthetic.forcing(T)
    rng = np.random.default_rng(3)
    qobs = np.abs(rng.normal(2.0, 1.0, T)) + 0.1
    fn = REF_METRIC[objective]
    Pa = synthetic.random_params(ABCModel(), N)
    got = engine.abc(f["prec"], 1.0, Pa, qobs=qobs, want_qsim=False, objective=objective)["mse"]
    np.testing.assert_allclose(got, _columns(fn, qobs, oracle.abc(f["prec"], 1.0, Pa)), rtol=OBJ_RTOL)
    Pg = synthetic.random_params(GR4J(), N)
    got = engine.gr4j(f["prec"], f["etp"], 0.6, 0.7, Pg, qobs=qobs, want_qsim=False, objective=objective)["mse"]
    np.testing.assert_allclose(got, _columns(fn, qobs, oracle.gr4j(f["prec"], f["etp"], 0.6, 0.7, Pg)), rtol=1e-8)
    c = load_golden("ensemble_cemaneige")
    g = load_golden("ensemble_cemaneigegr4j")
    Pc = synthetic.random_params(CemaneigeGR4J(), N)
    qo = np.abs(rng.normal(2.0, 1.0, c["layer_prec"].shape[0])) + 0.1
    full = engine.cemaneigegr4j(c["layer_prec"], c["layer_mean_temp"], g["etp"], c["frac_solid"], (0, 0, 0.6, 0.7), Pc)["qsim"]
    got = engine.cemaneigegr4j(c["layer_prec"], c["layer_mean_temp"], g["etp"], c["frac_solid"], (0, 0, 0.6, 0.7), Pc,
                               qobs=qo, want_qsim=False, objective=objective)["mse"]
    np.testing.assert_allclose(got, _columns(fn, qo, full), rtol=1e-8)


def test_fused_objective_undefined_cases_raise_like_the_reference():
    f, P, qobs = _hbv(T=60, N=4)
    args = (f["temp"], f["prec"], f["month"] - 1, f["PE_m"], f["T_m"], (0, 100, 3, 10), P)
    with pytest.raises(RuntimeError, match="Nash-Sutcliffe"):
        engine.hbvedu(*args, qobs=np.full(60, 2.0), objective="nse")
    with pytest.raises(RuntimeError, match="mean of the observations"):
        engine.hbvedu(*args, qobs=np.r_[np.ones(30), -np.ones(30)], objective="kge")
    with pytest.raises(RuntimeError, match="standard deviation"):
        engine.hbvedu(*args, qobs=np.full(60, 2.0), objective="kge")
    with pytest.raises(ValueError):
        engine.hbvedu(*args, qobs=qobs, objective="rmse")


def test_multi_catchment_objectives_use_per_catchment_statistics():
    Cn, T, N = 3, 400, 64
    fs = [synthetic.forcing(T, seed=synthetic.SEED + c) for c in range(Cn)]
    stack = lambda k: np.stack([f[k] for f in fs])
    P = np.stack([engine.pack_params(synthetic.random_params(HBVEdu(), N, seed=50 + c)) for c in range(Cn)])
    rng = np.random.default_rng(9)
    qobs = np.abs(rng.normal(2.0 + np.arange(Cn)[:, None], 1.0, (Cn, T))) + 0.1
    for objective in ("nse", "kge"):
        got = engine.hbvedu_multi(stack("temp"), stack("prec"), stack("month") - 1, stack("PE_m"), stack("T_m"),
                                  (0, 100, 3, 10), P, qobs=qobs, want_qsim=False, objective=objective)["mse"]
        for c in range(Cn):
            one = engine.hbvedu(fs[c]["temp"], fs[c]["prec"], fs[c]["month"] - 1, fs[c]["PE_m"], fs[c]["T_m"],
                                (0, 100, 3, 10), P[c], qobs=qobs[c], want_qsim=False, objective=objective)["mse"]
            assert_bits_equal(got[c], one, f"{objective} catchment {c}")


def test_monte_carlo_runs_on_the_fused_objective():
    f = synthetic.forcing(600)
    qobs = f["prec"] * 0.3 + 0.1
    kw = dict(temp=f["temp"], prec=f["prec"], month=f["month"], PE_m=f["PE_m"], T_m=f["T_m"], soil_init=100)
    np.random.seed(7)
    res = monte_carlo(HBVEdu(), num=96, qobs=qobs, **kw)
    assert set(res) == {"params", "qsim", "mse"}  # the reference's keys (rrmpg/tools/monte_carlo.py:73-76)
    ref = oracle.hbvedu(f["temp"], f["prec"], f["month"] - 1, f["PE_m"], f["T_m"], (0, 100, 0, 0), res["params"])
    assert_close(res["qsim"], ref, "monte_carlo qsim")
    np.testing.assert_allclose(res["mse"], oracle.mse_columns(qobs, ref), rtol=OBJ_RTOL)
    np.random.seed(7)
    lean = monte_carlo(HBVEdu(), num=96, qobs=qobs, return_qsim=False, **kw)  # extension: nothing but mse[num] comes back
    assert set(lean) == {"params", "mse"}
    assert_bits_equal(lean["mse"], res["mse"], "mse with and without the discharge array")
    np.random.seed(7)
    kge = monte_carlo(HBVEdu(), num=96, qobs=qobs, return_qsim=False, objective="kge", **kw)["mse"]
    np.testing.assert_allclose(kge, _columns(ref_kge, qobs, ref), rtol=OBJ_RTOL)
    with pytest.raises(ValueError, match="Arrays must have the same size"):
        monte_carlo(HBVEdu(), num=4, qobs=qobs[:-1], **kw)
    assert engine._FUSED is None
    # a model whose simulate returns a tuple / another first output name
    c = load_golden("fixture_cemaneigegr4j")
    np.random.seed(3)
    r = monte_carlo(CemaneigeGR4J(), num=8, qobs=np.abs(c["expected"]) + 0.1, prec=c["prec"], mean_temp=c["mean_temp"],
                    min_temp=c["min_temp"], max_temp=c["max_temp"], etp=c["etp"], met_station_height=495,
                    altitudes=[550, 620, 700, 785, 920], s_init=0.6, r_init=0.7)
    np.testing.assert_allclose(r["mse"], _columns(ref_mse, np.abs(c["expected"]) + 0.1, r["qsim"]), rtol=OBJ_RTOL)


def test_hyst_fit_losses_come_from_the_kernel():
    from rrmpg_b200.models import _snowice
    g = load_golden("fixture_cemaneigehystgr4j")
    model = CemaneigeHystGR4J()
    prepared = model._prepare(g["prec"], g["mean_temp"], g["min_temp"], g["max_temp"], g["etp"], None, 495, 0, 0, 0, 0.6,
                              0.7, [550, 620, 700, 785, 920])
    prec, mean_temp, etp, frac, frac_ice, inits = prepared
    obs = np.abs(g["expected"]) + 0.1
    np.random.seed(0)
    P = model.get_random_params(24)
    X = engine.pack_params(P).T.copy()
    full = engine.snowice_gr4j(True, False, prec, mean_temp, etp, None, frac, inits, P)["qsim"]
    for metric, fn in (("mse", ref_mse), ("kge", ref_kge)):
        args = (obs, prec, mean_temp, frac, etp, frac_ice, tuple(inits), model._dtype, metric, True, False)
        np.testing.assert_allclose(_snowice._loss(X, *args), _columns(fn, obs, full), rtol=1e-8)


# ------------------------------------------------------------------ resume
def _split_run(fn, T, cut, **state_kw):
    """fn(t0, t1, **kw) -> result dict; runs [0, cut) then [cut, T) through state_out -> state_in."""
    a = fn(0, cut, return_state=True)
    b = fn(cut, T, state_in=a["state"], return_state=True)
    return a, b


@pytest.mark.parametrize("cut", [1, 200, 333])
def test_resume_hbvedu_is_bit_identical_to_one_run(cut):
    T, N = 700, 130
    f, P, _ = _hbv(T, N)
    inits = (2.0, 100, 3, 10)
    run = lambda t0, t1, **kw: engine.hbvedu(f["temp"][t0:t1], f["prec"][t0:t1], (f["month"] - 1)[t0:t1], f["PE_m"], f["T_m"],
                                             inits, P, return_storage=True, **kw)
    one = run(0, T, return_state=True)
    a, b = _split_run(run, T, cut)
    assert a["state"].shape == (4, N)
    for nm in ("qsim", "snow", "soil", "s1", "s2"):
        assert_bits_equal(np.concatenate([a[nm], b[nm]]), one[nm], f"hbvedu resumed at {cut}: {nm}")
    assert_bits_equal(b["state"], one["state"], "final state")
    for row, nm in enumerate(("snow", "soil", "s1", "s2")):   # documented layout: snow, soil, s1, s2
        assert_bits_equal(one["state"][row], one[nm][-1], f"state row {row} = {nm}[T-1]")
    # with time slabs on both sides of the hand-over, and in PRECISE math
    a2 = run(0, cut, return_state=True, slab_steps=64)
    b2 = run(cut, T, state_in=a2["state"], slab_steps=50)
    assert_bits_equal(np.concatenate([a2["qsim"], b2["qsim"]]), one["qsim"], "resume + time slabs")
    p1 = run(0, T, math="precise")
    pa = run(0, cut, return_state=True, math="precise")
    pb = run(cut, T, state_in=pa["state"], math="precise")
    assert_bits_equal(np.concatenate([pa["qsim"], pb["qsim"]]), p1["qsim"], "resume, precise")


@pytest.mark.parametrize("x4_hi", [2.9, 3.9, 9.5, 15.5])
def test_resume_gr4j_carries_the_unit_hydrograph_buffers(x4_hi):
    """The reference neither returns nor accepts UH1 / UH2 (gr4j_model.py:82-83): a series cannot be continued there.
    state_out / state_in hand them over: S, R, uh1[C1], uh2[C2] (rrb_state_rows)."""
    T, N, cut = 500, 96, 217
    f = synthetic.forcing(T)
    P = synthetic.random_params(GR4J(), N)
    P["x4"] = np.random.default_rng(1).uniform(0.6, x4_hi, N)
    run = lambda t0, t1, **kw: engine.gr4j(f["prec"][t0:t1], f["etp"][t0:t1], 0.6, 0.7, P, return_storage=True,
                                           x4_max=float(P["x4"].max()), **kw)
    one = run(0, T, return_state=True)
    a, b = _split_run(run, T, cut)
    rows = _lib.lib().rrb_state_rows(_lib.MODEL_GR4J, float(P["x4"].max()))
    assert a["state"].shape == (rows, N) and rows >= 2 + 3 + 7
    for nm in ("qsim", "s_store", "r_store"):
        assert_bits_equal(np.concatenate([a[nm], b[nm]]), one[nm], f"gr4j resumed, x4 <= {x4_hi}: {nm}")
    assert_bits_equal(one["state"][0], one["s_store"][-1], "row 0 = S")
    assert_bits_equal(one["state"][1], one["r_store"][-1], "row 1 = R")
    assert_close(np.concatenate([a["qsim"], b["qsim"]]), oracle.gr4j(f["prec"], f["etp"], 0.6, 0.7, P), "vs oracle")


def test_resume_abc_and_device_mode():
    torch = pytest.importorskip("torch")
    T, N, cut = 300, 64, 100
    f = synthetic.forcing(T)
    P = synthetic.random_params(ABCModel(), N)
    one = engine.abc(f["prec"], 2.0, P, return_storage=True)
    a = engine.abc(f["prec"][:cut], 2.0, P, return_storage=True, return_state=True)
    b = engine.abc(f["prec"][cut:], 2.0, P, return_storage=True, state_in=a["state"])
    for nm in ("qsim", "storage"):
        assert_bits_equal(np.concatenate([a[nm], b[nm]]), one[nm], "abc resumed: " + nm)
    # device mode: state tensors stay on the GPU
    dev = torch.device("cuda:0")
    t = lambda x, dt=torch.float64: torch.as_tensor(np.ascontiguousarray(x), dtype=dt, device=dev)
    fh, Ph, _ = _hbv(T, N)
    Pd = t(engine.pack_params(Ph))
    args = lambda t0, t1: (t(fh["temp"][t0:t1]), t(fh["prec"][t0:t1]), t((fh["month"] - 1)[t0:t1], torch.int8), t(fh["PE_m"]),
                           t(fh["T_m"]), (0, 100, 3, 10), Pd)
    one = engine.hbvedu(*args(0, T))["qsim"].cpu().numpy()
    a = engine.hbvedu(*args(0, cut), return_state=True)
    assert a["state"].is_cuda and tuple(a["state"].shape) == (4, N)
    b = engine.hbvedu(*args(cut, T), state_in=a["state"])
    assert_bits_equal(np.concatenate([a["qsim"].cpu().numpy(), b["qsim"].cpu().numpy()]), one, "device-mode resume")
    # the Cemaneige family has no resume: G_tresh is a mean over the whole series (cemaneige_model.py:80), a continued
    # run could not reproduce one long run -- the C ABI refuses the request
    import ctypes as C
    o = _lib.Opts()
    o.struct_size = C.sizeof(_lib.Opts)
    z = np.zeros((10, 1)); pr = np.array([[0.5, 3.0]]); q = np.zeros((10, 1)); st = np.zeros((4, 1))
    o.state_out = st.ctypes.data
    rc = _lib.lib().rrb_cemaneige_simulate(z.ctypes.data, z.ctypes.data, z.ctypes.data, 10, 1, 0.0, 0.0, pr.ctypes.data, 2, 1,
                                           q.ctypes.data, None, None, C.byref(o))
    assert rc == _lib.RRB_EUNSUPPORTED and "state_in" in _lib.last_error()


# ------------------------------------------------------------------ in-library multi-device sharding
def _sharded_equals_single(devices):
    T, N = 500, 1000 + 37
    f, P, qobs = _hbv(T, N)
    args = (f["temp"], f["prec"], f["month"] - 1, f["PE_m"], f["T_m"], (0, 100, 3, 10), P)
    one = engine.hbvedu(*args, return_storage=True, qobs=qobs, devices="one")
    many = engine.hbvedu(*args, return_storage=True, qobs=qobs, devices=devices, slab_steps=128)
    for nm in one:
        assert_bits_equal(many[nm], one[nm], f"hbvedu sharded over {devices}: {nm}")
    c = load_golden("ensemble_cemaneige")
    g = load_golden("ensemble_cemaneigegr4j")
    Pc = synthetic.random_params(CemaneigeGR4J(), 300)
    cargs = (c["layer_prec"], c["layer_mean_temp"], g["etp"], c["frac_solid"], (0, 0, 0.6, 0.7), Pc)
    one = engine.cemaneigegr4j(*cargs, return_storages=True, devices="one")
    many = engine.cemaneigegr4j(*cargs, return_storages=True, devices=devices)   # [T, L, N] storages: pitched rows
    for nm in one:
        assert_bits_equal(many[nm], one[nm], f"cemaneigegr4j sharded over {devices}: {nm}")
    Pg = synthetic.random_params(GR4J(), 257)
    one = engine.gr4j(f["prec"], f["etp"], 0.6, 0.7, Pg, devices="one")["qsim"]
    assert_bits_equal(engine.gr4j(f["prec"], f["etp"], 0.6, 0.7, Pg, devices=devices)["qsim"], one, "gr4j sharded")
    Pa = synthetic.random_params(ABCModel(), 130)
    one = engine.abc(f["prec"], 0.0, Pa, devices="one")["qsim"]
    assert_bits_equal(engine.abc(f["prec"], 0.0, Pa, devices=devices)["qsim"], one, "abc sharded")


def test_member_sharding_machinery_on_one_device():
    """rrb_opts.n_devices with the same device listed several times: the blocks run one after the other (context mutex)
    through the very code path of a multi-GPU node -- worker threads, per-block staging, pitched row copies."""
    _sharded_equals_single([0, 0, 0])


def test_member_sharding_over_the_visible_devices():
    n = _lib.device_count()
    if n < 2:
        pytest.skip(f"{n} CUDA device(s) visible: the multi-GPU form needs two")
    _sharded_equals_single(list(range(n)))
    _sharded_equals_single("all")
    # the drop-in call itself uses the whole node for a large ensemble (engine.DEVICES = 'auto')
    f, P, _ = _hbv(300, 2 * engine.AUTO_MEMBERS_PER_DEVICE)
    kw = dict(temp=f["temp"], prec=f["prec"], month=f["month"], PE_m=f["PE_m"], T_m=f["T_m"], soil_init=100)
    q = HBVEdu().simulate(params=P, **kw)
    idx = np.r_[0:8, P.size - 8:P.size]
    assert_close(q[:, idx], oracle.hbvedu(f["temp"], f["prec"], f["month"] - 1, f["PE_m"], f["T_m"], (0, 100, 0, 0), P[idx]),
                 "auto-sharded drop-in call")


def test_sharding_argument_errors():
    f, P, _ = _hbv(50, 64)
    args = (f["temp"], f["prec"], f["month"] - 1, f["PE_m"], f["T_m"], (0, 100, 3, 10), P)
    with pytest.raises(ValueError):
        engine.hbvedu(*args, devices=[0, 99])
    with pytest.raises(ValueError):
        engine.hbvedu(*args, devices=[0, 0], return_state=True)


# ------------------------------------------------------------------ catchment batches of the remaining models
def test_abc_cemaneige_and_snow_ice_multi_catchment_equal_a_loop_over_catchments(monkeypatch):
    """rrb_abc / cemaneige / cemaneigegr4jice / cemaneigehystgr4j / cemaneigehystgr4jice _simulate_multi: grid.y =
    catchment with per-catchment forcing, G_tresh, initial states (incl. sca_init) and glacier fractions -- bit-identical
    to one call per catchment; host mode through the chunked D2H ring, device mode with the fused objective."""
    import torch
    from rrmpg_b200.models import _snow_inputs, Cemaneige, CemaneigeGR4JIce, CemaneigeHystGR4J, CemaneigeHystGR4JIce
    monkeypatch.setenv("RRMPG_B200_MULTI_CHUNK_BYTES", str(3 << 20))  # several chunks through the ring
    Cn, T, N = 3, 400, 70
    fs = [synthetic.forcing(T, seed=700 + c) for c in range(Cn)]
    qobs = np.abs(np.random.default_rng(6).normal(1.5, 0.5, (Cn, T))) + 0.1
    dev = torch.device("cuda:0")
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=dev)
    # ---- ABC
    prec = np.stack([f["prec"] for f in fs])
    Pa = np.stack([engine.pack_params(synthetic.random_params(ABCModel(), N, seed=800 + c)) for c in range(Cn)])
    s0 = np.array([0.0, 5.0, 12.5])
    multi = engine.abc_multi(prec, s0, Pa, return_storage=True, qobs=qobs, objective="nse")
    for c in range(Cn):
        one = engine.abc(prec[c], s0[c], Pa[c], return_storage=True, qobs=qobs[c], objective="nse")
        for nm in one:
            assert_bits_equal(multi[nm][c], one[nm], f"abc catchment {c} {nm}")
    pair = engine.abc_multi(prec, s0, Pa)["qsim"]                      # discharge only: the two-members-per-thread kernel
    assert_bits_equal(pair, multi["qsim"], "abc_multi pair kernel")
    r = engine.abc_multi(t(prec), s0, t(Pa), qobs=t(qobs), want_qsim=False, objective="nse")
    assert_bits_equal(r["mse"].cpu().numpy(), multi["mse"], "abc device-mode objective")
    # ---- layer arrays: catchments at different heights, one without precipitation (G_tresh = 0)
    alts = [550, 620, 700, 785, 920]
    lay = [_snow_inputs.to_layers(f["prec"] * (c != 1), f["temp"] - 2 * c, f["min_temp"] - 2 * c, f["max_temp"] - 2 * c,
                                  400 + 60 * c, np.array(alts, dtype=float)) for c, f in enumerate(fs)]
    lp, lt, fr = (np.stack([l[k] for l in lay]) for k in range(3))
    etp = np.stack([f["etp"] for f in fs])
    # ---- Cemaneige
    Pc = np.stack([engine.pack_params(synthetic.random_params(Cemaneige(), N, seed=810 + c)) for c in range(Cn)])
    ini2 = np.array([[0.0, 0.0], [20.0, -1.0], [3.0, 0.0]])
    multi = engine.cemaneige_multi(lp, lt, fr, ini2, Pc, return_storages=True, qobs=qobs)
    for c in range(Cn):
        one = engine.cemaneige(lp[c], lt[c], fr[c], ini2[c, 0], ini2[c, 1], Pc[c], return_storages=True, qobs=qobs[c])
        for nm in one:
            assert_bits_equal(multi[nm][c], one[nm], f"cemaneige catchment {c} {nm}")
    # ---- the three snow-ice couplings
    fice = np.array([[0.1, 0.2, 0.0, 0.3, 0.05], [0.0] * 5, [0.5, 0.1, 0.1, 0.1, 0.2]])
    for cls, hyst, ice in ((CemaneigeGR4JIce, False, True), (CemaneigeHystGR4J, True, False), (CemaneigeHystGR4JIce, True, True)):
        Ps = np.stack([engine.pack_params(synthetic.random_params(cls(), N, seed=820 + c)) for c in range(Cn)])
        ini = (np.array([[0, 0, 0.6, 0.7], [10, -0.5, 0.2, 0.9], [0, 0, 0.0, 0.1]], dtype=float) if not hyst else
               np.array([[0, 0, 0.0, 0.6, 0.7], [10, -0.5, 0.4, 0.2, 0.9], [0, 0, 1.0, 0.0, 0.1]], dtype=float))
        x4 = float(Ps[..., (4 if hyst else 2) + 3].max())
        multi = engine.snowice_gr4j_multi(hyst, ice, lp, lt, etp, fice if ice else None, fr, ini, Ps, return_storages=True,
                                          qobs=qobs, objective="kge", x4_max=x4)
        for c in range(Cn):
            one = engine.snowice_gr4j(hyst, ice, lp[c], lt[c], etp[c], fice[c] if ice else None, fr[c], ini[c], Ps[c],
                                      return_storages=True, qobs=qobs[c], objective="kge", x4_max=x4)
            assert set(one) == set(multi)
            for nm in one:
                assert_bits_equal(multi[nm][c], one[nm], f"{cls.__name__} catchment {c} {nm}")
        r = engine.snowice_gr4j_multi(hyst, ice, t(lp), t(lt), t(etp), t(fice) if ice else None, t(fr), ini, t(Ps),
                                      qobs=t(qobs), want_qsim=False, objective="kge", x4_max=x4)
        assert set(r) == {"mse"}
        assert_bits_equal(r["mse"].cpu().numpy(), multi["mse"], f"{cls.__name__} device-mode objective")
    # one-step series: sca[t-1] at t = 0 wraps to sca[T-1] = sca_init when T == 1 (cemaneigehyst_model.py:126)
    Ps = np.stack([engine.pack_params(synthetic.random_params(CemaneigeHystGR4J(), 8, seed=830 + c)) for c in range(Cn)])
    ini = np.array([[5, 0, 0.3, 0.6, 0.7], [5, 0, 0.9, 0.6, 0.7], [5, 0, 0.0, 0.6, 0.7]], dtype=float)
    m1 = engine.snowice_gr4j_multi(True, False, lp[:, :1], lt[:, :1], etp[:, :1], None, fr[:, :1], ini, Ps, return_storages=True,
                                   x4_max=10.0)
    for c in range(Cn):
        o1 = engine.snowice_gr4j(True, False, lp[c, :1], lt[c, :1], etp[c, :1], None, fr[c, :1], ini[c], Ps[c],
                                 return_storages=True, x4_max=10.0)
        for nm in o1:
            assert_bits_equal(m1[nm][c], o1[nm], f"T = 1, catchment {c} {nm}")
