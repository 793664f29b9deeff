"""FAST-path transcendentals evaluated on the host (same header, same tables as the kernels) vs glibc."""
import ctypes as C

import numpy as np
import pytest

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


def _hbv_pow_step(soil, FC, Beta, liquid, sp):
    a = [np.ascontiguousarray(v, np.float64) for v in (soil, FC, Beta, liquid, sp)]
    pe = np.empty_like(a[0]); sn = np.empty_like(a[0])
    _lib.lib().rrb_host_hbv_pow_step(*[v.ctypes.data for v in a], a[0].size, pe.ctypes.data, sn.ctypes.data)
    return pe, sn


def test_hbv_soil_chain_step_twin_vs_libm():
    """Round-2 soil chain (rr_math.cuh hbv_pow_step_twin = the device sequence of hbv_fast2_kernel): hoisted log2(FC),
    512-entry log2 / 1024-entry exp2 tables, fused shift, late table scale.  prec_eff = liquid (soil/FC)^Beta within
    8e-15 relative of glibc over the default bounds (hbvedu.py:47-60), 1e-13 over the whole admitted range
    soil/FC in [2^-15, 2^15), |Beta| < 32; soil_new = soil_partial - prec_eff to the rounding of that subtraction."""
    rng = np.random.default_rng(11)
    n = 400_000
    FC = rng.uniform(100, 200, n); Beta = rng.uniform(1, 7, n)
    soil = FC * rng.uniform(0.05, 1.5, n); liquid = rng.gamma(0.8, 6.0, n); sp = soil + liquid
    pe, sn = _hbv_pow_step(soil, FC, Beta, liquid, sp)
    ref = liquid * np.power(soil / FC, Beta)
    rel = np.abs(pe - ref) / ref
    assert rel.max() < 8e-15 and rel.mean() < 1e-15, (rel.max(), rel.mean())  # incl. the rounding of soil/FC itself (x Beta)
    assert np.max(np.abs(sn - (sp - ref)) / np.abs(sp)) < 4e-15
    # the whole admitted operand range
    FC = 10 ** rng.uniform(-3, 6, n); Beta = rng.uniform(-31.9, 31.9, n)
    soil = FC * 2 ** rng.uniform(-14.99, 14.99, n)
    pe, sn = _hbv_pow_step(soil, FC, Beta, np.ones(n), np.zeros(n))
    ref = np.power(soil / FC, Beta)
    assert np.all(np.isfinite(pe)) and (np.abs(pe - ref) / ref).max() < 1e-13
    assert np.array_equal(sn, -pe)  # fma(-scale, w, 0) = -(scale w)
    # mantissa-bin and table-index edges: soil at powers of two and just below them, soil == FC
    FC = np.full(8, 150.0); soil = np.array([128.0, np.nextafter(128.0, 0), 256.0, np.nextafter(256.0, 0), 150.0, 64.0, 1.0, 1e-2])
    pe, _ = _hbv_pow_step(soil, FC, np.full(8, 3.3), np.ones(8), np.zeros(8))
    ref = np.power(soil / FC, 3.3)
    assert (np.abs(pe - ref) / ref).max() < 4e-15
    # no liquid water: prec_eff is +0 and the soil moisture passes through
    pe, sn = _hbv_pow_step([120.0], [150.0], [2.0], [0.0], [119.5])
    assert pe[0] == 0.0 and sn[0] == 119.5


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


def test_third_order_corrections_of_the_seed_units():
    """rr_math.cuh: rcp_sane / rsqrt4_sane / pow35_sane refine a MUFU.RCP64H / MUFU.RSQ64H seed (which only sees and
    only produces the high 32 bits of a double: ~2^-20 relative) with ONE third-order correction.  The same arithmetic
    in numpy, with seeds truncated to their high word, stays within a few ulp over the operand ranges of the kernels."""
    rng = np.random.default_rng(5)
    hi_only = lambda a: (np.ascontiguousarray(a).view(np.uint64) & np.uint64(0xFFFFFFFF00000000)).view(np.float64)
    seed_rsq = lambda x: hi_only(1.0 / np.sqrt(hi_only(x)))
    seed_rcp = lambda x: hi_only(1.0 / hi_only(x))
    ulp = 2.0 ** -52
    # v^(-1/4) for v = 1 + u^4 >= 1: y (1 + e/4 + 5 e^2/32), e = 1 - v y^4
    v = 1.0 + rng.random(200000) ** 4 * 10.0 ** rng.uniform(-8, 40, 200000)
    a = seed_rsq(v)
    y = a * seed_rsq(a)
    assert np.max(np.abs(y * v ** 0.25 - 1)) < 2.0 ** -18          # the double seed
    y2 = y * y
    e = 1.0 - v * (y2 * y2)
    y = y + y * (e * (e * 0.15625 + 0.25))
    assert np.max(np.abs(y * v ** 0.25 - 1)) < 4 * ulp
    # w^3.5 = w^4 w^(-1/2): y (1 + e/2 + 3 e^2/8), e = 1 - w y^2
    w = 10.0 ** rng.uniform(-30, 9, 200000)
    y0 = seed_rsq(w)
    e = 1.0 - (w * y0) * y0
    y = y0 + (y0 * e) * (e * 0.375 + 0.5)
    w2 = w * w
    assert np.max(np.abs((w2 * w2) * y / w ** 3.5 - 1)) < 6 * ulp
    # 1/den for den >= 2: y (1 + e + e^2), e = 1 - den y
    den = 2.0 + 10.0 ** rng.uniform(-6, 20, 200000)
    y0 = seed_rcp(den)
    e = 1.0 - den * y0
    y = y0 + y0 * (e * e + e)
    assert np.max(np.abs(y * den - 1)) < 4 * ulp


def test_gr4j_fast_step_algebra_matches_the_reference_recurrence():
    """Host twin of Gr4jMember::step_fast (rr_gr4j.cuh) in numpy -- the branch-free formulation (wet / dry through the
    sign of P - E, tanh as m / (m + 2) with m = expm1(2a), both production-store branches as num m / (2 + c m),
    S (1 - y) as one product, 0.9 / 0.1 folded into the ordinates) -- against the oracle's reference-order recurrence.
    Exact libm functions stand in for the seed-unit sequences (covered by the test above)."""
    import oracle
    from rrmpg_b200 import synthetic
    from rrmpg_b200.models import GR4J
    T, N = 600, 300
    f = synthetic.forcing(T, seed=31)
    P = synthetic.random_params(GR4J(), N, seed=32)
    x1, x2, x3, x4 = (P[k].astype(np.float64) for k in ("x1", "x2", "x3", "x4"))
    C1, C2 = 3, 7
    sc1 = lambda t: np.where(t <= 0, 0.0, np.where(t < x4, (np.maximum(t, 0) / x4) ** 2.5, 1.0))
    def sc2(t):
        r = np.where(t <= 0, 0.0, 0.5 * (np.maximum(t, 0) / x4) ** 2.5)
        r = np.where((t > x4) & (t < 2 * x4), 1 - 0.5 * np.abs(2 - t / x4) ** 2.5, r)
        return np.where(t >= 2 * x4, 1.0, r)
    o1 = [0.9 * (sc1(j) - sc1(j - 1)) for j in range(1, C1 + 1)]
    o2 = [0.1 * (sc2(j) - sc2(j - 1)) for j in range(1, C2 + 1)]
    u1 = [np.zeros(N) for _ in range(C1)]
    u2 = [np.zeros(N) for _ in range(C2)]
    S, R = 0.6 * x1, 0.7 * x3
    inv_x1, inv_x3, k49 = 1.0 / x1, 1.0 / x3, (4.0 / 9.0) / x1
    q = np.empty((T, N))
    for t in range(T):
        d = f["prec"][t] - f["etp"][t]
        wet = d >= 0
        arg = abs(d)
        sr = S * inv_x1
        m = np.expm1(2.0 * arg * inv_x1)
        sgn = 1.0 if wet else -1.0
        c = sgn * sr + (1.0 if wet else 2.0)
        base = x1 if wet else S + S
        num = base - S * sr
        frac = (num * m) / (c * m + 2.0)
        S = S + sgn * frac
        u = S * k49
        perc = S - S * (1.0 + (u * u) ** 2) ** -0.25
        S = S - perc
        p_r = perc + ((arg - frac) if wet else 0.0)
        for j in range(C1 - 1):
            u1[j] = o1[j] * p_r + u1[j + 1]
        u1[C1 - 1] = o1[C1 - 1] * p_r
        for j in range(C2 - 1):
            u2[j] = o2[j] * p_r + u2[j + 1]
        u2[C2 - 1] = o2[C2 - 1] * p_r
        w = R * inv_x3
        gw = x2 * (w ** 4 / np.sqrt(np.where(w > 0, w, 1.0)))
        R = np.maximum(0.0, (R + u1[0]) + gw)
        v = R * inv_x3
        q_r = R - R * (1.0 + (v * v) ** 2) ** -0.25
        R = R - q_r
        q[t] = q_r + np.maximum(0.0, u2[0] + gw)
    ref = oracle.gr4j(f["prec"], f["etp"], 0.6, 0.7, P)
    assert np.allclose(q, ref, rtol=1e-10, atol=1e-12)
    assert np.max(np.abs(q - ref) / (np.abs(ref) + 1e-6)) < 1e-11


def test_hbv_fast_snow_routine_is_bitwise_the_reference_one_for_finite_rain():
    """rr_hbvedu.cu (FAST): melt = min(snow, m) serves max(0, snow - m) = snow - melt and the liquid water
    prec + melt; the cold branch is the same two additions with -prec in place of melt.  Elementwise, on random and
    adversarial operands, both formulations give the same bits (hbvedu_model.py:87-96)."""
    rng = np.random.default_rng(9)
    n = 400000
    snow = np.where(rng.random(n) < 0.3, 0.0, rng.gamma(1.0, 20.0, n))
    prec = np.where(rng.random(n) < 0.5, 0.0, rng.gamma(0.8, 6.0, n))
    temp = rng.normal(2.0, 8.0, n)
    T_t = rng.uniform(-3, 3, n)
    DD = rng.uniform(0.5, 5, n)
    # adversarial: exact ties and zeros
    snow[:1000] = DD[:1000] * (temp[:1000] - T_t[:1000])
    temp[1000:2000] = T_t[1000:2000]
    prec[2000:3000] = 0.0
    snow[3000:4000] = -rng.random(1000)          # a negative pack (negative snow_init)
    m = DD * (temp - T_t)
    cold = temp < T_t
    # reference (numba semantics: max(0, x) = x if x > 0 else 0; min(a, b) = b if b < a else a)
    d = snow - m
    ref_snow = np.where(cold, snow + prec, np.where(d > 0, d, 0.0))
    ref_liq = np.where(cold, 0.0, prec + np.where(m < snow, m, snow))
    # FAST formulation
    melt = np.where(m < snow, m, snow)
    sel = np.where(np.signbit(temp - T_t), -prec, melt)       # cold read off the sign of temp - T_t
    got_snow = snow - sel
    got_liq = prec + sel
    same = lambda a, b: np.array_equal(a.view(np.uint64), b.view(np.uint64))
    assert same(got_snow, ref_snow)
    assert same(got_liq, ref_liq)


def test_cemaneige_contract_absorbs_the_quotient_of_a_vanishing_pack():
    """rr_cemaneige.cuh, contract step: for a pack below 2^-960 the unchecked Markstein quotient G / G_tresh may have
    wrong low bits, but it stays below 2^-900 (G_tresh >= 2^-60) and 0.9 ratio + 0.1 rounds to 0.1 for every ratio
    below 2^-58 -- exactly what the reference computes from the exact tiny quotient (cemaneige_model.py:109-115)."""
    r = np.concatenate([2.0 ** -np.arange(58, 1075, dtype=np.float64), [0.0, 5e-324]])
    assert np.all(0.9 * r + 0.1 == 0.1)
    assert 0.9 * 2.0 ** -52 + 0.1 != 0.1          # ... and not for quotients that matter


def test_hbv_fast_step_algebra_matches_the_reference_recurrence():
    """Host twin of hbv_fast_kernel (rr_hbvedu.cu) in numpy: FAST forcing packing (dT * PEm), hoisted reciprocals and
    linear-store coefficients, the select-based snow routine, pow skipped where liquid water is zero, sign-bit
    max(0, s1 - L) -- against the oracle's reference-order recurrence (hbvedu_model.py:84-127)."""
    import oracle
    from rrmpg_b200 import synthetic
    from rrmpg_b200.models import HBVEdu
    T, N = 800, 400
    f = synthetic.forcing(T, seed=41)
    P = synthetic.random_params(HBVEdu(), N, seed=42)
    g = lambda k: P[k].astype(np.float64)
    T_t, DD, FC, Beta, C, PWP, K_0, K_1, K_2, K_p, L = (g(k) for k in
                                                       ("T_t", "DD", "FC", "Beta", "C", "PWP", "K_0", "K_1", "K_2", "K_p", "L"))
    m0 = f["month"] - 1
    dTPEm = (f["temp"] - f["T_m"][m0]) * f["PE_m"][m0]      # FAST packing of the pack kernel
    PEm = f["PE_m"][m0]
    inv_FC, inv_PWP, c1, c2 = 1.0 / FC, 1.0 / PWP, 1.0 - K_1 - K_p, 1.0 - K_2
    snow, soil, s1, s2 = np.zeros(N), np.full(N, 100.0), np.full(N, 3.0), np.full(N, 10.0)
    q = np.zeros((T, N))
    for t in range(1, T):
        dtt = f["temp"][t] - T_t
        m = DD * dtt
        melt = np.where(m < snow, m, snow)
        sel = np.where(np.signbit(dtt), -f["prec"][t], melt)
        snow = snow - sel
        liquid = f["prec"][t] + sel
        pe = C * dTPEm[t] + PEm[t]
        ea = np.where(soil > PWP, pe, pe * (soil * inv_PWP))
        x = s1 - L
        oK = np.where(np.signbit(x), 0.0, x) * K_0
        s2_new = s1 * K_p + s2 * c2
        s1_new = s1 * c1 - oK
        soil_new = (soil + liquid) - ea
        wet = liquid != 0
        prec_eff = np.where(wet, liquid * np.power(soil * inv_FC, Beta, where=wet, out=np.ones(N)), 0.0)
        soil, s1, s2 = soil_new - prec_eff, s1_new + prec_eff, s2_new
        q[t] = s2 * K_2 + (s1 * K_1 + oK)
    ref = oracle.hbvedu(f["temp"], f["prec"], m0, f["PE_m"], f["T_m"], (0, 100, 3, 10), P)
    assert np.allclose(q, ref, rtol=1e-10, atol=1e-12)
    assert np.max(np.abs(q - ref) / (np.abs(ref) + 1e-6)) < 1e-12


@pytest.mark.parametrize("mode,n", [(0, 4_000_000), (1, 4_000_000), (2, 2_000_000)])
def test_division_by_a_loop_invariant_is_the_ieee_quotient(mode, n):
    """rr_common.cuh div_by_invariant / the Cemaneige contract step: the Markstein sequence (reciprocal computed once,
    two FMA residual corrections) equals a / b bit for bit over the operand ranges the kernels admit without a
    range check; restated in the oracle with exact fma()."""
    import oracle
    assert oracle.check_invariant_division(n, seed=12345 + mode, mode=mode) == 0


def test_cemaneige_contract_step_twin_is_bit_identical_to_the_reference_routine():
    """The contract step of cema_kernel (unconditional, unchecked G / G_tresh; selects for potential melt and ratio)
    restated in C with exact fma() in the oracle library, against run_cemaneige's restatement: every bit of outflow, G and
    eTG over long series, in-contract edge members (Kf = 0, CTG in {0, 1}, huge Kf, cold and warm starts) and packs that
    decay towards zero over decades without snowfall."""
    import oracle
    from rrmpg_b200 import synthetic
    from rrmpg_b200.models import Cemaneige, _snow_inputs
    same = lambda a, b: np.array_equal(np.ascontiguousarray(a).view(np.uint64), np.ascontiguousarray(b).view(np.uint64))
    for seed, T, alts, warm in ((1, 3000, [550, 620, 700, 785, 920], 0.0), (2, 1500, [], 0.0), (3, 2500, [1600, 2400, 3100], 0.0),
                                (4, 14610, [500.0, 900.0], 25.0)):
        f = synthetic.forcing(T, seed=100 + seed)
        temp = f["temp"] + warm                     # warm = 25: it never snows, an initial pack melts away for 40 years
        lp, lt, fr, _ = _snow_inputs.to_layers(f["prec"], temp, f["min_temp"] + warm, f["max_temp"] + warm, 480.0,
                                               np.array(alts, dtype=float))
        P = synthetic.random_params(Cemaneige(), 48, seed=200 + seed)
        P["Kf"][:4] = [0.0, 1e-9, 1e3, 1e6]
        P["CTG"][4:8] = [0.0, 1.0, 1e-12, 0.999999999]
        for g0, e0 in ((0.0, 0.0), (35.0, -2.0), (1e-300, 0.0)):
            ref = oracle.cemaneige(lp, lt, fr, g0, e0, P, return_storages=True)
            got = oracle.cemaneige_contract_twin(lp, lt, fr, g0, e0, P)
            for a, b, nm in zip(got, ref, ("outflow", "G", "eTG")):
                assert same(a, b), f"seed {seed} g0={g0} {nm}"
