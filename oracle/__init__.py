"""CPU parity oracle for rrmpg_b200 -- TEST INFRASTRUCTURE, not product code.

ctypes front-end to ``oracle/liboracle.so`` (built from ``oracle/rr_oracle.c`` by
``oracle/Makefile``).  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this package; ``rrmpg_b200``
never does.

Each function mirrors the signature of the reference numba kernel it restates
(``rrmpg/models/*_model.py``) but takes the parameter sets as a float64 ``[N, k]``
matrix (the packed view of the reference's structured record array) and returns the
ensemble result in the wrappers' ``[T, N]`` / ``[T, L, N]`` layout
(``rrmpg/models/hbvedu.py:191-209``).

Parity status: PINNED.  ``tests/golden/make_golden.py`` checks this oracle bit-for-bit
against the live numba reference (random ensembles) and against the reference's four
golden fixtures (``test/test_models.py:142-268``); the resulting vectors are committed
under ``tests/golden/`` and re-checked by ``tests/test_oracle.py`` on every run.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")
_lib = None

_dp = C.POINTER(C.c_double)
_bp = C.POINTER(C.c_int8)


def build(force=False):
    """Compile liboracle.so with gcc (see oracle/Makefile)."""
    src = os.path.join(_HERE, "rr_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "liboracle.so"])
    return _SO


def reference_package():
    """The unmodified reference installed under oracle/_ref by oracle/build_ref.py (git-ignored; travels to the GPU
    box with the snapshot).  Returns the imported `rrmpg` package, or None when it (or numba) is not available."""
    import importlib
    import sys
    ref = os.path.join(_HERE, "_ref")
    if not os.path.isdir(os.path.join(ref, "rrmpg")):
        return None
    if ref not in sys.path:
        sys.path.insert(0, ref)
    try:
        import numba  # noqa: F401
        pkg = importlib.import_module("rrmpg")
        importlib.import_module("rrmpg.models")
    except Exception:
        return None
    return pkg if os.path.abspath(pkg.__file__).startswith(ref) else None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.oracle_num_threads.restype = C.c_int
        _lib.oracle_gr4j_batch.restype = C.c_int
        _lib.oracle_cemaneigegr4j_batch.restype = C.c_int
        _lib.oracle_snowice_gr4j_batch.restype = C.c_int
        _lib.oracle_check_invariant_division.restype = C.c_int64
        _lib.oracle_check_invariant_division.argtypes = [C.c_int64, C.c_uint64, C.c_int]
    return _lib


def num_threads():
    return int(lib().oracle_num_threads())


def _d(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def pack_params(params):
    """Structured record array (reference ``_dtype``) or [N,k] matrix -> float64 [N,k]."""
    params = np.asarray(params)
    if params.dtype.names:
        params = np.atleast_1d(params)
        k = len(params.dtype.names)
        out = np.empty((params.size, k), np.float64)
        for j, name in enumerate(params.dtype.names):
            out[:, j] = params[name]
        return out
    return np.ascontiguousarray(np.atleast_2d(params), dtype=np.float64)


def abc(prec, initial_state, params, return_storage=False, nthreads=0):
    """run_abcmodel over an ensemble (rrmpg/models/abcmodel_model.py:16-60)."""
    prec = _f64(prec); P = pack_params(params); T, N = prec.size, P.shape[0]
    q = np.zeros((T, N)); s = np.zeros((T, N)) if return_storage else None
    lib().oracle_abc_batch(_d(prec), C.c_int64(T), C.c_double(initial_state), _d(P), C.c_int64(N),
                           _d(q), _d(s), C.c_int(nthreads))
    return (q, s) if return_storage else q


def hbvedu(temp, prec, month0, PE_m, T_m, inits, params, return_storage=False, nthreads=0):
    """run_hbvedu over an ensemble (rrmpg/models/hbvedu_model.py:16-129).

    ``month0`` is the 0-based int8 month index the numba kernel receives.
    ``inits`` = (snow_init, soil_init, s1_init, s2_init).
    """
    temp = _f64(temp); prec = _f64(prec); PE_m = _f64(PE_m); T_m = _f64(T_m)
    month0 = np.ascontiguousarray(month0, dtype=np.int8)
    inits = _f64(inits); P = pack_params(params); T, N = prec.size, P.shape[0]
    q = np.zeros((T, N))
    st = [np.zeros((T, N)) for _ in range(4)] if return_storage else [None] * 4
    lib().oracle_hbvedu_batch(_d(temp), _d(prec), month0.ctypes.data_as(_bp), _d(PE_m), _d(T_m),
                              C.c_int64(T), _d(inits), _d(P), C.c_int64(N), _d(q),
                              _d(st[0]), _d(st[1]), _d(st[2]), _d(st[3]), C.c_int(nthreads))
    return (q, *st) if return_storage else q


def gr4j(prec, etp, s_init, r_init, params, return_storage=False, nthreads=0):
    """run_gr4j over an ensemble (rrmpg/models/gr4j_model.py:16-157)."""
    prec = _f64(prec); etp = _f64(etp); P = pack_params(params); T, N = prec.size, P.shape[0]
    q = np.zeros((T, N))
    st = [np.zeros((T, N)) for _ in range(2)] if return_storage else [None] * 2
    rc = lib().oracle_gr4j_batch(_d(prec), _d(etp), C.c_int64(T), C.c_double(s_init),
                                 C.c_double(r_init), _d(P), C.c_int64(N), _d(q), _d(st[0]),
                                 _d(st[1]), C.c_int(nthreads))
    if rc:
        raise RuntimeError("oracle_gr4j: unit hydrograph length out of range")
    return (q, *st) if return_storage else q


def cemaneige(prec, mean_temp, frac_solid, snow_pack_init, thermal_state_init, params,
              return_storages=False, nthreads=0):
    """run_cemaneige over an ensemble (rrmpg/models/cemaneige_model.py:16-127).

    prec / mean_temp / frac_solid are the preprocessed ``[T, L]`` arrays.
    """
    prec = _f64(prec); mean_temp = _f64(mean_temp); frac_solid = _f64(frac_solid)
    P = pack_params(params); (T, L), N = prec.shape, P.shape[0]
    out = np.zeros((T, N))
    G = np.zeros((T, L, N)) if return_storages else None
    E = np.zeros((T, L, N)) if return_storages else None
    lib().oracle_cemaneige_batch(_d(prec), _d(mean_temp), _d(frac_solid), C.c_int64(T),
                                 C.c_int64(L), C.c_double(snow_pack_init),
                                 C.c_double(thermal_state_init), _d(P), C.c_int64(P.shape[1]),
                                 C.c_int64(N), _d(out), _d(G), _d(E), C.c_int(nthreads))
    return (out, G, E) if return_storages else out


def cemaneigegr4j(prec, mean_temp, etp, frac_solid, inits, params, return_storages=False,
                  nthreads=0):
    """run_cemaneigegr4j over an ensemble (rrmpg/models/cemaneigegr4j_model.py:17-64).

    ``inits`` = (snow_pack_init, thermal_state_init, s_init, r_init).
    """
    prec = _f64(prec); mean_temp = _f64(mean_temp); frac_solid = _f64(frac_solid); etp = _f64(etp)
    inits = _f64(inits); P = pack_params(params); (T, L), N = prec.shape, P.shape[0]
    q = np.zeros((T, N))
    G = np.zeros((T, L, N)) if return_storages else None
    E = np.zeros((T, L, N)) if return_storages else None
    s = np.zeros((T, N)) if return_storages else None
    r = np.zeros((T, N)) if return_storages else None
    rc = lib().oracle_cemaneigegr4j_batch(_d(prec), _d(mean_temp), _d(etp), _d(frac_solid),
                                          C.c_int64(T), C.c_int64(L), _d(inits), _d(P),
                                          C.c_int64(N), _d(q), _d(G), _d(E), _d(s), _d(r),
                                          C.c_int(nthreads))
    if rc:
        raise RuntimeError("oracle_cemaneigegr4j: unit hydrograph length out of range")
    return (q, G, E, s, r) if return_storages else q


def snowice_gr4j(hyst, ice, prec, mean_temp, etp, frac_ice, frac_solid, inits, params, return_storages=False,
                 nthreads=0):
    """The snow(+hysteresis)(+ice)+GR4J couplings over an ensemble.

    hyst=0, ice=1: run_cemaneigegr4jice (rrmpg/models/cemaneigegr4jice_model.py:16-93), records of 7 fields;
    hyst=1, ice=0: run_cemaneigehystgr4j (cemaneigehystgr4j_model.py:17-79), 8 fields;
    hyst=1, ice=1: run_cemaneigehystgr4jice (cemaneigehystgr4jice_model.py:18-104), 9 fields.
    ``inits`` = (snow_pack_init, thermal_state_init, sca_init, s_init, r_init).
    Returns qsim or (qsim, G, eTG, s_store, r_store, sca, icemelt, snowmelt, rain[T,L]).
    """
    prec = _f64(prec); mean_temp = _f64(mean_temp); frac_solid = _f64(frac_solid); etp = _f64(etp)
    inits = _f64(inits); P = pack_params(params); (T, L), N = prec.shape, P.shape[0]
    fice = _f64(frac_ice) if frac_ice is not None else np.zeros(L)
    assert P.shape[1] == 6 + (2 if hyst else 0) + (1 if ice else 0) and inits.size == 5
    q = np.zeros((T, N))
    G = E = S = s = r = im = sm = None
    if return_storages:
        G, E, S = (np.zeros((T, L, N)) for _ in range(3))
        s, r, im, sm = (np.zeros((T, N)) for _ in range(4))
    rc = lib().oracle_snowice_gr4j_batch(C.c_int(int(hyst)), C.c_int(int(ice)), _d(prec), _d(mean_temp), _d(etp),
                                         _d(fice), _d(frac_solid), C.c_int64(T), C.c_int64(L), _d(inits), _d(P),
                                         C.c_int64(N), _d(q), _d(G), _d(E), _d(S), _d(s), _d(r), _d(im), _d(sm),
                                         C.c_int(nthreads))
    if rc:
        raise RuntimeError("oracle_snowice_gr4j: unit hydrograph length out of range")
    if not return_storages:
        return q
    rain = prec - prec * frac_solid  # cemaneigehyst_model.py:98-99 (member independent)
    return q, G, E, s, r, S, im, sm, rain


def extrapolate_precipitation(prec, altitudes, met_station_height):
    """rrmpg/models/cemaneige_utils.py:101-158."""
    prec = _f64(prec); alt = _f64(altitudes); out = np.zeros((prec.size, alt.size))
    lib().oracle_extrapolate_precipitation(_d(prec), C.c_int64(prec.size), _d(alt),
                                           C.c_int64(alt.size), C.c_double(met_station_height),
                                           _d(out))
    return out


def extrapolate_temperature(min_temp, mean_temp, max_temp, altitudes, met_station_height):
    """rrmpg/models/cemaneige_utils.py:161-208."""
    a, b, c = _f64(min_temp), _f64(mean_temp), _f64(max_temp); alt = _f64(altitudes)
    o = [np.zeros((a.size, alt.size)) for _ in range(3)]
    lib().oracle_extrapolate_temperature(_d(a), _d(b), _d(c), C.c_int64(a.size), _d(alt),
                                         C.c_int64(alt.size), C.c_double(met_station_height),
                                         _d(o[0]), _d(o[1]), _d(o[2]))
    return tuple(o)


def calculate_solid_fraction(prec, altitudes, mean_temp, min_temp, max_temp):
    """rrmpg/models/cemaneige_utils.py:16-98 (prec only supplies the shape)."""
    alt = _f64(altitudes); a, b, c = _f64(mean_temp), _f64(min_temp), _f64(max_temp)
    out = np.zeros(a.shape)
    lib().oracle_solid_fraction(_d(alt), C.c_int64(alt.size), _d(a), _d(b), _d(c),
                                C.c_int64(a.shape[0]), _d(out))
    return out


def mse_columns(qobs, qsim):
    """Per-member calc_mse (rrmpg/tools/monte_carlo.py:70-71)."""
    qobs = _f64(qobs); qsim = _f64(qsim); T, N = qsim.shape
    out = np.zeros(N)
    lib().oracle_mse_columns(_d(qobs), _d(qsim), C.c_int64(T), C.c_int64(N), _d(out))
    return out


def check_invariant_division(n, seed=1, mode=0):
    """Mismatches of the kernels' division-by-an-invariant sequence against IEEE division over n operand pairs."""
    return int(lib().oracle_check_invariant_division(int(n), int(seed), int(mode)))


def cemaneige_contract_twin(prec, mean_temp, frac_solid, snow_pack_init, thermal_state_init, params):
    """CPU twin of the CUDA contract snow step (test helper; see rr_oracle.c).  Returns (outflow, G, eTG)."""
    prec = _f64(prec); mean_temp = _f64(mean_temp); frac_solid = _f64(frac_solid)
    P = pack_params(params); (T, L), N = prec.shape, P.shape[0]
    out, G, E = np.zeros((T, N)), np.zeros((T, L, N)), np.zeros((T, L, N))
    lib().oracle_cemaneige_contract_twin(_d(prec), _d(mean_temp), _d(frac_solid), C.c_int64(T), C.c_int64(L),
                                         C.c_double(snow_pack_init), C.c_double(thermal_state_init), _d(P),
                                         C.c_int64(P.shape[1]), C.c_int64(N), _d(out), _d(G), _d(E))
    return out, G, E
