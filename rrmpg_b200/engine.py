"""Batched ensemble simulations: the array-level API over the C ABI.

One call here = one pass of the hot path over a whole ensemble: it replaces the reference's
``for i in range(params.size): run_<model>(..., params[i])`` loops
(``rrmpg/models/hbvedu.py:199-209`` and siblings) by a single library call.

Two calling modes, chosen by the argument types:

* **numpy** arrays: host mode.  The library uploads forcing + parameters, runs the kernel over
  time slabs and streams finished output rows back while the next slab computes.  Outputs are
  numpy arrays (pinned host memory when large).
* **torch CUDA tensors** (float64, contiguous): device mode.  Nothing leaves the GPU; the work
  is enqueued on torch's current stream and the outputs are CUDA tensors.

Every function returns a dict: ``{'qsim': [T,N], <storage name>: ..., 'mse': [N]}`` (keys
present only when requested).
"""
import ctypes as C

import numpy as np

from . import _lib

_MATH = {"fast": _lib.MATH_FAST, "precise": _lib.MATH_PRECISE,
         _lib.MATH_FAST: _lib.MATH_FAST, _lib.MATH_PRECISE: _lib.MATH_PRECISE}

DEFAULT_MATH = "fast"
_OBJECTIVE = {"mse": _lib.OBJ_MSE, "nse": _lib.OBJ_NSE, "kge": _lib.OBJ_KGE,
              _lib.OBJ_MSE: _lib.OBJ_MSE, _lib.OBJ_NSE: _lib.OBJ_NSE, _lib.OBJ_KGE: _lib.OBJ_KGE}
# Host-mode calls of the single-catchment models shard the ensemble over several GPUs inside the library (one worker
# thread per device, rrb_opts.n_devices).  "auto": every visible device, for ensembles that give each at least
# AUTO_MEMBERS_PER_DEVICE members -- unless the process is one rank of a multi-process job (torchrun: RANK / LOCAL_RANK
# set), where each rank drives its own GPU.  "one" / None: the current device only; "all", a count or a list force it.
import os as _os
DEVICES = _os.environ.get("RRMPG_B200_DEVICES", "one" if ("LOCAL_RANK" in _os.environ or "RANK" in _os.environ) else "auto")
if DEVICES.isdigit():
    DEVICES = int(DEVICES)
AUTO_MEMBERS_PER_DEVICE = 32768
VARIANT = 0  # rrb_opts.variant of every call made through this module (kernel A/B timing and variant parity tests)


class fused:
    """Fuse a per-member objective into the next ensemble call made through a model's ``simulate``:

        with engine.fused(qobs, objective="mse", want_qsim=False) as f:
            model.simulate(params=P, ...)        # returns None for qsim when want_qsim=False
        f.values                                  # [N] objective, accumulated in the kernel's registers

    The drop-in ``simulate`` signatures are the reference's and have no ``qobs`` argument; this is how
    ``rrmpg_b200.tools.monte_carlo`` asks for the objective of the very launch that simulates the ensemble
    (rrmpg/tools/monte_carlo.py:64-73 loops calc_mse over the columns afterwards).  Not re-entrant / thread-safe.
    """

    def __init__(self, qobs, objective="mse", want_qsim=True):
        self.qobs, self.objective, self.want_qsim = qobs, objective, bool(want_qsim)
        self.values = None
        self.calls = 0

    def __enter__(self):
        global _FUSED
        if _FUSED is not None:
            raise RuntimeError("engine.fused() contexts do not nest")
        _FUSED = self
        return self

    def __exit__(self, *exc):
        global _FUSED
        _FUSED = None
        return False


_FUSED = None


def _want(flag):
    """want_qsim / want_outflow of a call under an engine.fused() context."""
    return bool(flag) and (_FUSED is None or _FUSED.want_qsim)


def _is_torch(a):
    return a is not None and type(a).__module__.split(".")[0] == "torch"


def pack_params(params):
    """Structured record array of a model ``_dtype`` (or [N,k] matrix) -> C-contiguous float64 [N,k]."""
    if _is_torch(params):
        return params
    params = np.asarray(params)
    if params.dtype.names:
        params = np.ascontiguousarray(np.atleast_1d(params))
        k = len(params.dtype.names)
        if params.dtype.itemsize == 8 * k and all(params.dtype[n] == np.float64 for n in params.dtype.names):
            return params.view(np.float64).reshape(params.size, k)  # zero-copy
        out = np.empty((params.size, k), np.float64)
        for j, name in enumerate(params.dtype.names):
            out[:, j] = params[name]
        return out
    return np.ascontiguousarray(np.atleast_2d(params), dtype=np.float64)


class _Call:
    """Collects arguments of one library call in either mode."""

    def __init__(self, arrays, math, device, block, slab_steps, qobs, x4_max=0.0, objective="mse", devices=None,
                 state_in=None, return_state=False, fusable=True):
        self.torch_mode = any(_is_torch(a) for a in arrays)
        self.opts = _lib.Opts()
        self.opts.struct_size = C.sizeof(_lib.Opts)
        self.opts.math = _MATH[math]
        self.opts.block = int(block)
        self.opts.slab_steps = int(slab_steps)
        self.opts.variant = int(VARIANT)
        self.opts.x4_max = float(x4_max)
        self.keep = []
        if self.torch_mode:
            import torch
            self.torch = torch
            devs = {a.device for a in arrays if _is_torch(a)}
            if len(devs) != 1 or next(iter(devs)).type != "cuda":
                raise ValueError("device mode needs every tensor on the same CUDA device")
            self.dev = next(iter(devs))
            self.opts.mem = _lib.MEM_DEVICE
            self.opts.device = self.dev.index if self.dev.index is not None else torch.cuda.current_device()
            self.opts.stream = torch.cuda.current_stream(self.dev).cuda_stream
        else:
            self.opts.mem = _lib.MEM_HOST
            self.opts.device = -1 if device is None else int(device)
        self.mse = None
        self.state = None
        self._state_in, self._return_state = state_in, bool(return_state)
        if qobs is None and _FUSED is not None and fusable:
            qobs, objective = _FUSED.qobs, _FUSED.objective
            _FUSED.calls += 1
        if objective not in _OBJECTIVE:
            raise ValueError(f"objective must be one of {sorted(k for k in _OBJECTIVE if isinstance(k, str))}")
        self.opts.objective = _OBJECTIVE[objective]
        if qobs is not None:
            q = self.f64(qobs)
            self.opts.qobs = _lib.ptr(q)
            self._qobs_len = q.shape[-1]
            if self.opts.objective != _lib.OBJ_MSE:
                self._obs_stats(qobs)
        if not self.torch_mode:
            self._devices(devices)
        elif devices not in (None, "one"):
            raise ValueError("devices= needs host (numpy) arrays: device tensors live on one GPU")

    def _obs_stats(self, qobs):
        """(np.mean, np.std) of every observed series, computed like rrmpg/utils/metrics.py:62-70,164-173 and with the
        reference's errors for the cases in which NSE / KGE are not defined."""
        q = qobs.detach().cpu().numpy() if _is_torch(qobs) else np.asarray(qobs, dtype=np.float64)
        q = np.atleast_2d(q)
        mean, std = np.mean(q, axis=1), np.std(q, axis=1)
        if self.opts.objective == _lib.OBJ_NSE and np.any(np.sum((q - mean[:, None]) ** 2, axis=1) == 0):
            raise RuntimeError("The Nash-Sutcliffe-Efficiency coefficient is not defined for the case, that all values in "
                               "the observations are equal. Maybe you should use the Mean-Squared-Error instead.")
        if self.opts.objective == _lib.OBJ_KGE:
            if np.any(mean == 0):
                raise RuntimeError("KGE not definied if the mean of the observations equals 0.")
            if np.any(std == 0):
                raise RuntimeError("KGE not definied if the standard deviation of the observations equals 0.")
        stats = np.ascontiguousarray(np.stack([mean, std], axis=1), dtype=np.float64)
        self.keep.append(stats)
        self.opts.obs_stats = stats.ctypes.data

    def _devices(self, devices):
        """rrb_opts.n_devices / devices: None = DEVICES (module default), "all", an int count or a list of ordinals."""
        if devices is None:
            devices = DEVICES
        if devices is None or devices == "one":
            return
        if isinstance(devices, str):
            if devices not in ("all", "auto"):
                raise ValueError("devices must be None, 'one', 'all', 'auto', a count or a list of device ordinals")
            self._auto_devices = devices == "auto"
            devices = list(range(_lib.device_count()))
        elif isinstance(devices, (int, np.integer)):
            devices = list(range(int(devices)))
        devices = [int(d) for d in devices]
        if len(devices) > 1:
            arr = np.ascontiguousarray(devices, dtype=np.int32)
            self.keep.append(arr)
            self.opts.n_devices = arr.size
            self.opts.devices = arr.ctypes.data
        elif len(devices) == 1:
            self.opts.device = devices[0]

    def limit_devices(self, N):
        """'auto': shard only ensembles large enough to fill several GPUs (>= AUTO_MEMBERS_PER_DEVICE members each)."""
        if getattr(self, "_auto_devices", False) and self.opts.n_devices > 1:
            n = max(1, min(int(self.opts.n_devices), int(N) // AUTO_MEMBERS_PER_DEVICE))
            self.opts.n_devices = n if n > 1 else 0

    def want_state(self, model, N, x4_max=0.0):
        """state_in / state_out of the resumable models (rrb_state_rows gives the row count of the layout)."""
        if self._state_in is None and not self._return_state:
            return
        rows = int(_lib.lib().rrb_state_rows(model, float(x4_max)))
        if rows < 0:
            raise ValueError("no state layout for this model / x4_max")
        if self.opts.n_devices > 1:
            raise ValueError("state_in / return_state are not supported together with several devices")
        if self._state_in is not None:
            st = self.f64(self._state_in, (rows, N))
            self.opts.state_in = _lib.ptr(st)
        if self._return_state:
            self.state = self.empty((rows, N))
            self.opts.state_out = _lib.ptr(self.state)

    def f64(self, a, shape=None):
        if self.torch_mode:
            if not _is_torch(a):
                a = self.torch.as_tensor(np.asarray(a, dtype=np.float64), device=self.dev)
            if a.dtype != self.torch.float64:
                raise TypeError("device mode needs float64 tensors")
            a = a.contiguous()
        else:
            a = np.ascontiguousarray(a, dtype=np.float64)
        if shape is not None and tuple(a.shape) != tuple(shape):
            raise ValueError(f"array of shape {tuple(a.shape)}, expected {tuple(shape)}")
        self.keep.append(a)
        return a

    def i8(self, a):
        if self.torch_mode:
            if not _is_torch(a):
                a = self.torch.as_tensor(np.asarray(a, dtype=np.int8), device=self.dev)
            if a.dtype != self.torch.int8:
                raise TypeError("device mode needs an int8 month tensor")
            a = a.contiguous()
        else:
            a = np.ascontiguousarray(a, dtype=np.int8)
        self.keep.append(a)
        return a

    def month0(self, a):
        """0-based int8 month index; host arrays are range-checked (the packer indexes PE_m / T_m with it)."""
        a = self.i8(a)
        if not self.torch_mode and a.size and (a.min() < 0 or a.max() > 11):
            raise ValueError("month0 must be the 0-based month index in [0, 11]")
        return a

    def host_f64(self, a, n):
        a = np.ascontiguousarray(np.asarray(a, dtype=np.float64).reshape(-1))
        if a.size != n:
            raise ValueError(f"expected {n} values, got {a.size}")
        self.keep.append(a)
        return a

    def empty(self, shape, given=None):
        if given is not None:
            if tuple(given.shape) != tuple(shape):
                raise ValueError(f"out array of shape {tuple(given.shape)}, expected {tuple(shape)}")
            if self.torch_mode != _is_torch(given):
                raise TypeError("out array must live where the inputs live")
            if self.torch_mode:
                if given.dtype != self.torch.float64 or not given.is_contiguous():
                    raise TypeError("out tensor must be contiguous float64")
            elif given.dtype != np.float64 or not given.flags.c_contiguous:
                raise TypeError("out array must be C-contiguous float64")
            return given
        if self.torch_mode:
            return self.torch.empty(tuple(shape), dtype=self.torch.float64, device=self.dev)
        return _lib.host_empty(shape)

    def want_mse(self, N, T):
        if self.opts.qobs:
            if self._qobs_len != T:
                raise ValueError("Arrays must have the same size.")  # rrmpg/utils/metrics.py:127-128
            self.mse = self.empty((N,))
            self.opts.mse = _lib.ptr(self.mse)


def _run(call, T, N, fn, *args):
    """Call the library unless the problem is empty (T == 0 or N == 0: nothing to simulate)."""
    if T > 0 and N > 0:
        _lib.check(fn(*args))
    elif call.mse is not None and N > 0:
        call.mse[...] = float("nan")  # np.mean over an empty series (rrmpg/utils/metrics.py:131)


def _result(call, names, arrays):
    out = {n: a for n, a in zip(names, arrays) if a is not None}
    if call.mse is not None:
        out["mse"] = call.mse          # the fused objective: MSE, NSE or KGE per member
    if call.state is not None:
        out["state"] = call.state      # [rrb_state_rows, N] stores after the last timestep
    if _FUSED is not None and call.mse is not None:
        _FUSED.values = call.mse
        out.setdefault(names[0], None)  # 'qsim' / 'outflow' not materialised (fused(want_qsim=False))
    return out


def abc(prec, initial_state, params, return_storage=False, qobs=None, want_qsim=True,
        math=DEFAULT_MATH, device=None, block=0, slab_steps=0, out=None, objective="mse", devices=None, state_in=None, return_state=False):
    """ABC model ensemble (run_abcmodel, rrmpg/models/abcmodel_model.py:16-60)."""
    P0 = pack_params(params)
    c = _Call([prec, P0], math, device, block, slab_steps, qobs, objective=objective, devices=devices, state_in=state_in, return_state=return_state)
    prec = c.f64(prec); P = c.f64(P0)
    T, N = prec.shape[0], P.shape[0]
    if P.shape[1] != 3:
        raise ValueError("ABC parameter records have 3 fields (a, b, c)")
    out = out or {}
    q = c.empty((T, N), out.get("qsim")) if _want(want_qsim) else None
    s = c.empty((T, N), out.get("storage")) if return_storage else None
    c.want_mse(N, T)
    c.limit_devices(N)
    c.want_state(_lib.MODEL_ABC, N)
    _run(c, T, N, _lib.lib().rrb_abc_simulate, _lib.ptr(prec), T, float(initial_state), _lib.ptr(P), N,
         _lib.ptr(q), _lib.ptr(s), C.byref(c.opts))
    return _result(c, ["qsim", "storage"], [q, s])


def hbvedu(temp, prec, month0, PE_m, T_m, inits, params, return_storage=False, qobs=None,
           want_qsim=True, math=DEFAULT_MATH, device=None, block=0, slab_steps=0, out=None, objective="mse", devices=None, state_in=None, return_state=False):
    """HBV-Edu ensemble (run_hbvedu, rrmpg/models/hbvedu_model.py:16-129).

    ``month0`` is the 0-based int8 month index; ``inits`` = (snow, soil, s1, s2).
    """
    P0 = pack_params(params)
    c = _Call([temp, prec, month0, PE_m, T_m, P0], math, device, block, slab_steps, qobs, objective=objective, devices=devices, state_in=state_in, return_state=return_state)
    temp = c.f64(temp); prec = c.f64(prec); month0 = c.month0(month0)
    PE_m = c.f64(PE_m, (12,)); T_m = c.f64(T_m, (12,)); P = c.f64(P0)
    inits = c.host_f64(inits, 4)
    T, N = prec.shape[0], P.shape[0]
    if temp.shape[0] != T or month0.shape[0] != T:
        raise ValueError("temp, prec and month must have the same length")
    if P.shape[1] != 11:
        raise ValueError("HBVEdu parameter records have 11 fields")
    out = out or {}
    q = c.empty((T, N), out.get("qsim")) if _want(want_qsim) else None
    names = ["snow", "soil", "s1", "s2"]
    st = [c.empty((T, N), out.get(n)) for n in names] if return_storage else [None] * 4
    c.want_mse(N, T)
    c.limit_devices(N)
    c.want_state(_lib.MODEL_HBVEDU, N)
    _run(c, T, N, _lib.lib().rrb_hbvedu_simulate,
         _lib.ptr(temp), _lib.ptr(prec), _lib.ptr(month0), _lib.ptr(PE_m), _lib.ptr(T_m), T,
         _lib.ptr(inits), _lib.ptr(P), N, _lib.ptr(q), *[_lib.ptr(a) for a in st], C.byref(c.opts))
    return _result(c, ["qsim"] + names, [q] + st)


def snowice_gr4j(hyst, ice, prec, mean_temp, etp, frac_ice, frac_solid, inits, params, return_storages=False,
                 qobs=None, want_qsim=True, math=DEFAULT_MATH, device=None, block=0, slab_steps=0, out=None, x4_max=0.0, objective="mse", devices=None):
    """The snow(+hysteresis)(+ice)+GR4J couplings of the reference over an ensemble.

    hyst=False, ice=True : CemaneigeGR4JIce     (run_cemaneigegr4jice, cemaneigegr4jice_model.py:16-93), 7 fields,
                           inits = (snow_pack_init, thermal_state_init, s_init, r_init)
    hyst=True,  ice=False: CemaneigeHystGR4J    (run_cemaneigehystgr4j, cemaneigehystgr4j_model.py:17-79), 8 fields
    hyst=True,  ice=True : CemaneigeHystGR4JIce (run_cemaneigehystgr4jice, cemaneigehystgr4jice_model.py:18-104), 9
                           inits = (snow_pack_init, thermal_state_init, sca_init, s_init, r_init) for both Hyst models
    Result keys: qsim, G, eTG, s_store, r_store, then sca (hyst), icemelt (ice), snowmelt (hyst and ice).
    """
    if not (hyst or ice):
        raise ValueError("use cemaneigegr4j for the plain coupling")
    P0 = pack_params(params)
    c = _Call([prec, mean_temp, etp, frac_solid, P0], math, device, block, slab_steps, qobs, x4_max, objective=objective, devices=devices)
    prec = c.f64(prec); mean_temp = c.f64(mean_temp, prec.shape); frac_solid = c.f64(frac_solid, prec.shape)
    etp = c.f64(etp); P = c.f64(P0)
    inits = c.host_f64(inits, 5 if hyst else 4)
    if prec.ndim != 2:
        raise ValueError("layer arrays must be [T, L]")
    (T, L), N = prec.shape, P.shape[0]
    fice = c.f64(frac_ice, (L,)) if ice else None
    if etp.shape[0] != T:
        raise ValueError("etp must have the same length as the layer arrays")
    k = 6 + (2 if hyst else 0) + (1 if ice else 0)
    if P.shape[1] != k:
        raise ValueError(f"parameter records of this model have {k} fields")
    out = out or {}
    st = return_storages
    q = c.empty((T, N), out.get("qsim")) if _want(want_qsim) else None
    G = c.empty((T, L, N), out.get("G")) if st else None
    E = c.empty((T, L, N), out.get("eTG")) if st else None
    s = c.empty((T, N), out.get("s_store")) if st else None
    r = c.empty((T, N), out.get("r_store")) if st else None
    sca = c.empty((T, L, N), out.get("sca")) if (st and hyst) else None
    im = c.empty((T, N), out.get("icemelt")) if (st and ice) else None
    sm = c.empty((T, N), out.get("snowmelt")) if (st and hyst and ice) else None
    c.want_mse(N, T)
    c.limit_devices(N)
    p_ = _lib.ptr
    L_ = _lib.lib()
    if hyst and ice:
        _run(c, T, N, L_.rrb_cemaneigehystgr4jice_simulate, p_(prec), p_(mean_temp), p_(etp), p_(fice), p_(frac_solid),
             T, L, p_(inits), p_(P), N, p_(q), p_(G), p_(E), p_(s), p_(r), p_(sca), p_(im), p_(sm), C.byref(c.opts))
    elif hyst:
        _run(c, T, N, L_.rrb_cemaneigehystgr4j_simulate, p_(prec), p_(mean_temp), p_(etp), p_(frac_solid), T, L,
             p_(inits), p_(P), N, p_(q), p_(G), p_(E), p_(s), p_(r), p_(sca), C.byref(c.opts))
    else:
        _run(c, T, N, L_.rrb_cemaneigegr4jice_simulate, p_(prec), p_(mean_temp), p_(etp), p_(fice), p_(frac_solid), T, L,
             p_(inits), p_(P), N, p_(q), p_(G), p_(E), p_(s), p_(r), p_(im), C.byref(c.opts))
    return _result(c, ["qsim", "G", "eTG", "s_store", "r_store", "sca", "icemelt", "snowmelt"],
                   [q, G, E, s, r, sca, im, sm])


def hbvedu_multi(temp, prec, month0, PE_m, T_m, inits, params, return_storage=False, qobs=None, want_qsim=True,
                 math=DEFAULT_MATH, device=None, block=0, out=None, objective="mse"):
    """HBV-Edu for C independent catchments with N members each, one launch.

    temp, prec, month0: [C, T]; PE_m, T_m: [C, 12]; inits: (4,) or [C, 4]; params: [C, N, 11] (or a [C, N]
    record array); qobs: [C, T].  Returns {'qsim': [C, T, N], storages..., 'mse': [C, N]}.
    Equivalent to looping ``hbvedu`` over the catchments (tests/test_parity_gpu.py).
    """
    if not _is_torch(params):
        params = np.asarray(params)
        if params.dtype.names:
            Cn, Nn = params.shape
            params = pack_params(params.reshape(-1)).reshape(Cn, Nn, -1)
    c = _Call([temp, prec, month0, PE_m, T_m, params], math, device, block, 0, None, objective=objective, devices="one", fusable=False)
    temp = c.f64(temp); prec = c.f64(prec); month0 = c.month0(month0); PE_m = c.f64(PE_m); T_m = c.f64(T_m)
    P = c.f64(params)
    if temp.ndim != 2 or P.ndim != 3 or P.shape[2] != 11:
        raise ValueError("expected temp/prec/month0 [C, T] and params [C, N, 11]")
    (Cc, T), N = temp.shape, P.shape[1]
    for a, shp in ((prec, (Cc, T)), (month0, (Cc, T)), (PE_m, (Cc, 12)), (T_m, (Cc, 12))):
        if tuple(a.shape) != shp:
            raise ValueError(f"array of shape {tuple(a.shape)}, expected {shp}")
    if P.shape[0] != Cc:
        raise ValueError("params must have one [N, 11] block per catchment")
    ini = np.asarray(inits, dtype=np.float64)
    ini = np.ascontiguousarray(np.broadcast_to(ini.reshape(-1, 4), (Cc, 4)))
    c.keep.append(ini)
    _multi_objective(c, qobs, Cc, T, N)
    out = out or {}
    q = c.empty((Cc, T, N), out.get("qsim")) if _want(want_qsim) else None
    names = ["snow", "soil", "s1", "s2"]
    st = [c.empty((Cc, T, N), out.get(n)) for n in names] if return_storage else [None] * 4
    if Cc > 0 and T > 0:
        _lib.check(_lib.lib().rrb_hbvedu_simulate_multi(
            _lib.ptr(temp), _lib.ptr(prec), _lib.ptr(month0), _lib.ptr(PE_m), _lib.ptr(T_m), Cc, T, _lib.ptr(ini),
            _lib.ptr(P), N, _lib.ptr(q), *[_lib.ptr(a) for a in st], C.byref(c.opts)))
    return _result(c, ["qsim"] + names, [q] + st)


def _x4_hint(P, col):
    """Largest x4 of a host parameter matrix (device tensors: let the library reduce)."""
    if _is_torch(P) or P.shape[0] == 0:
        return 0.0
    return float(np.max(P[:, col]))


def gr4j(prec, etp, s_init, r_init, params, return_storage=False, qobs=None, want_qsim=True,
         math=DEFAULT_MATH, device=None, block=0, slab_steps=0, out=None, x4_max=0.0, objective="mse", devices=None, state_in=None, return_state=False):
    """GR4J ensemble (run_gr4j, rrmpg/models/gr4j_model.py:16-157); every member is simulated."""
    P0 = pack_params(params)
    c = _Call([prec, etp, P0], math, device, block, slab_steps, qobs, x4_max, objective=objective, devices=devices, state_in=state_in, return_state=return_state)
    prec = c.f64(prec); etp = c.f64(etp); P = c.f64(P0)
    T, N = prec.shape[0], P.shape[0]
    if etp.shape[0] != T:
        raise ValueError("prec and etp must have the same length")
    if P.shape[1] != 4:
        raise ValueError("GR4J parameter records have 4 fields (x1, x2, x3, x4)")
    out = out or {}
    q = c.empty((T, N), out.get("qsim")) if _want(want_qsim) else None
    names = ["s_store", "r_store"]
    st = [c.empty((T, N), out.get(n)) for n in names] if return_storage else [None] * 2
    c.want_mse(N, T)
    c.limit_devices(N)
    if state_in is not None or return_state:
        xm = x4_max if x4_max and x4_max > 0 else _x4_hint(P, 3)
        if not xm > 0:
            raise ValueError("device mode: pass x4_max together with state_in / return_state (it fixes the state layout)")
        c.opts.x4_max = float(xm)
        c.want_state(_lib.MODEL_GR4J, N, xm)
    _run(c, T, N, _lib.lib().rrb_gr4j_simulate, _lib.ptr(prec), _lib.ptr(etp), T, float(s_init), float(r_init),
         _lib.ptr(P), N, _lib.ptr(q), _lib.ptr(st[0]), _lib.ptr(st[1]), C.byref(c.opts))
    return _result(c, ["qsim"] + names, [q] + st)


def cemaneige(prec, mean_temp, frac_solid, snow_pack_init, thermal_state_init, params,
              return_storages=False, qobs=None, want_outflow=True, math=DEFAULT_MATH, device=None,
              block=0, slab_steps=0, out=None, objective="mse", devices=None):
    """Cemaneige ensemble (run_cemaneige, rrmpg/models/cemaneige_model.py:16-127).

    ``prec``, ``mean_temp``, ``frac_solid`` are the preprocessed [T, L] layer arrays.  ``params`` may
    be Cemaneige records (CTG, Kf) or any record type whose first two fields are (CTG, Kf).
    """
    P0 = pack_params(params)
    c = _Call([prec, mean_temp, frac_solid, P0], math, device, block, slab_steps, qobs, objective=objective, devices=devices)
    prec = c.f64(prec); mean_temp = c.f64(mean_temp, prec.shape); frac_solid = c.f64(frac_solid, prec.shape)
    P = c.f64(P0)
    if prec.ndim != 2:
        raise ValueError("layer arrays must be [T, L]")
    (T, L), N = prec.shape, P.shape[0]
    if P.shape[1] < 2:
        raise ValueError("Cemaneige parameter records start with (CTG, Kf)")
    out = out or {}
    q = c.empty((T, N), out.get("outflow")) if _want(want_outflow) else None
    names = ["G", "eTG"]
    st = [c.empty((T, L, N), out.get(n)) for n in names] if return_storages else [None] * 2
    c.want_mse(N, T)
    c.limit_devices(N)
    _run(c, T, N, _lib.lib().rrb_cemaneige_simulate,
         _lib.ptr(prec), _lib.ptr(mean_temp), _lib.ptr(frac_solid), T, L, float(snow_pack_init),
         float(thermal_state_init), _lib.ptr(P), P.shape[1], N, _lib.ptr(q), _lib.ptr(st[0]),
         _lib.ptr(st[1]), C.byref(c.opts))
    return _result(c, ["outflow"] + names, [q] + st)


def cemaneigegr4j(prec, mean_temp, etp, frac_solid, inits, params, return_storages=False, qobs=None,
                  want_qsim=True, math=DEFAULT_MATH, device=None, block=0, slab_steps=0, out=None,
                  x4_max=0.0, objective="mse", devices=None):
    """Cemaneige + GR4J ensemble (run_cemaneigegr4j, rrmpg/models/cemaneigegr4j_model.py:17-64).

    ``inits`` = (snow_pack_init, thermal_state_init, s_init, r_init).
    """
    P0 = pack_params(params)
    c = _Call([prec, mean_temp, etp, frac_solid, P0], math, device, block, slab_steps, qobs, x4_max, objective=objective, devices=devices)
    prec = c.f64(prec); mean_temp = c.f64(mean_temp, prec.shape); frac_solid = c.f64(frac_solid, prec.shape)
    etp = c.f64(etp); P = c.f64(P0)
    inits = c.host_f64(inits, 4)
    if prec.ndim != 2:
        raise ValueError("layer arrays must be [T, L]")
    (T, L), N = prec.shape, P.shape[0]
    if etp.shape[0] != T:
        raise ValueError("etp must have the same length as the layer arrays")
    if P.shape[1] != 6:
        raise ValueError("CemaneigeGR4J parameter records have 6 fields")
    out = out or {}
    q = c.empty((T, N), out.get("qsim")) if _want(want_qsim) else None
    G = c.empty((T, L, N), out.get("G")) if return_storages else None
    E = c.empty((T, L, N), out.get("eTG")) if return_storages else None
    s = c.empty((T, N), out.get("s_store")) if return_storages else None
    r = c.empty((T, N), out.get("r_store")) if return_storages else None
    c.want_mse(N, T)
    c.limit_devices(N)
    _run(c, T, N, _lib.lib().rrb_cemaneigegr4j_simulate,
         _lib.ptr(prec), _lib.ptr(mean_temp), _lib.ptr(etp), _lib.ptr(frac_solid), T, L, _lib.ptr(inits),
         _lib.ptr(P), N, _lib.ptr(q), _lib.ptr(G), _lib.ptr(E), _lib.ptr(s), _lib.ptr(r), C.byref(c.opts))
    return _result(c, ["qsim", "G", "eTG", "s_store", "r_store"], [q, G, E, s, r])


def _multi_objective(c, qobs, Cc, T, N):
    """qobs [C, T] -> the fused objective [C, N] of a catchment batch (per-catchment statistics for NSE / KGE)."""
    if qobs is not None:
        q_ = c.f64(qobs, (Cc, T))
        c.opts.qobs = _lib.ptr(q_)
        c.mse = c.empty((Cc, N))
        c.opts.mse = _lib.ptr(c.mse)
        if c.opts.objective != _lib.OBJ_MSE:
            c._obs_stats(qobs)


def _multi_common(c, params, width, qobs, Cc, T):
    P = c.f64(params)
    if P.ndim != 3 or P.shape[2] != width or P.shape[0] != Cc:
        raise ValueError(f"params must be [C, N, {width}] with one block per catchment")
    N = P.shape[1]
    _multi_objective(c, qobs, Cc, T, N)
    return P, N


def _records_to_matrix(params):
    if not _is_torch(params):
        params = np.asarray(params)
        if params.dtype.names:
            Cn, Nn = params.shape
            params = pack_params(params.reshape(-1)).reshape(Cn, Nn, -1)
    return params


def gr4j_multi(prec, etp, inits, params, return_storage=False, qobs=None, want_qsim=True, math=DEFAULT_MATH,
               device=None, block=0, out=None, x4_max=0.0, objective="mse"):
    """GR4J for C independent catchments with N members each, one launch.

    prec, etp: [C, T]; inits: (2,) or [C, 2] = (s_init, r_init); params: [C, N, 4] (or a [C, N] record array);
    qobs: [C, T].  Returns {'qsim': [C, T, N], 's_store', 'r_store', 'mse': [C, N]}; bit-identical to looping
    ``gr4j`` over the catchments."""
    params = _records_to_matrix(params)
    c = _Call([prec, etp, params], math, device, block, 0, None, x4_max, objective=objective, devices="one", fusable=False)
    prec = c.f64(prec); etp = c.f64(etp, prec.shape)
    if prec.ndim != 2:
        raise ValueError("expected prec / etp [C, T]")
    Cc, T = prec.shape
    P, N = _multi_common(c, params, 4, qobs, Cc, T)
    ini = np.ascontiguousarray(np.broadcast_to(np.asarray(inits, dtype=np.float64).reshape(-1, 2), (Cc, 2)))
    c.keep.append(ini)
    out = out or {}
    q = c.empty((Cc, T, N), out.get("qsim")) if _want(want_qsim) else None
    st = [c.empty((Cc, T, N), out.get(n)) for n in ("s_store", "r_store")] if return_storage else [None, None]
    if Cc > 0 and T > 0 and N > 0:
        _lib.check(_lib.lib().rrb_gr4j_simulate_multi(_lib.ptr(prec), _lib.ptr(etp), Cc, T, _lib.ptr(ini), _lib.ptr(P), N,
                                                      _lib.ptr(q), _lib.ptr(st[0]), _lib.ptr(st[1]), C.byref(c.opts)))
    return _result(c, ["qsim", "s_store", "r_store"], [q] + st)


def cemaneigegr4j_multi(prec, mean_temp, etp, frac_solid, inits, params, return_storages=False, qobs=None,
                        want_qsim=True, math=DEFAULT_MATH, device=None, block=0, out=None, x4_max=0.0, objective="mse"):
    """Cemaneige + GR4J for C independent catchments with N members each, one launch.

    prec, mean_temp, frac_solid: [C, T, L] layer arrays (``snow_layers`` per catchment); etp: [C, T];
    inits: (4,) or [C, 4] = (snow_pack_init, thermal_state_init, s_init, r_init); params: [C, N, 6].
    Returns {'qsim': [C, T, N], 'G', 'eTG': [C, T, L, N], 's_store', 'r_store', 'mse': [C, N]}; bit-identical to
    looping ``cemaneigegr4j`` over the catchments."""
    params = _records_to_matrix(params)
    c = _Call([prec, mean_temp, etp, frac_solid, params], math, device, block, 0, None, x4_max, objective=objective, devices="one", fusable=False)
    prec = c.f64(prec); mean_temp = c.f64(mean_temp, prec.shape); frac_solid = c.f64(frac_solid, prec.shape)
    if prec.ndim != 3:
        raise ValueError("layer arrays must be [C, T, L]")
    Cc, T, L = prec.shape
    etp = c.f64(etp, (Cc, T))
    P, N = _multi_common(c, params, 6, qobs, Cc, T)
    ini = np.ascontiguousarray(np.broadcast_to(np.asarray(inits, dtype=np.float64).reshape(-1, 4), (Cc, 4)))
    c.keep.append(ini)
    out = out or {}
    q = c.empty((Cc, T, N), out.get("qsim")) if _want(want_qsim) else None
    if return_storages:
        G, E = c.empty((Cc, T, L, N), out.get("G")), c.empty((Cc, T, L, N), out.get("eTG"))
        s, r = c.empty((Cc, T, N), out.get("s_store")), c.empty((Cc, T, N), out.get("r_store"))
    else:
        G = E = s = r = None
    if Cc > 0 and T > 0 and N > 0:
        _lib.check(_lib.lib().rrb_cemaneigegr4j_simulate_multi(
            _lib.ptr(prec), _lib.ptr(mean_temp), _lib.ptr(etp), _lib.ptr(frac_solid), Cc, T, L, _lib.ptr(ini), _lib.ptr(P),
            N, _lib.ptr(q), _lib.ptr(G), _lib.ptr(E), _lib.ptr(s), _lib.ptr(r), C.byref(c.opts)))
    return _result(c, ["qsim", "G", "eTG", "s_store", "r_store"], [q, G, E, s, r])


def abc_multi(prec, inits, params, return_storage=False, qobs=None, want_qsim=True, math=DEFAULT_MATH, device=None,
              block=0, out=None, objective="mse"):
    """ABC model for C independent catchments with N members each, one launch.

    prec: [C, T]; inits: scalar or [C] initial storages; params: [C, N, 3] (or a [C, N] record array); qobs: [C, T].
    Returns {'qsim': [C, T, N], 'storage', 'mse': [C, N]}; bit-identical to looping ``abc`` over the catchments."""
    params = _records_to_matrix(params)
    c = _Call([prec, params], math, device, block, 0, None, objective=objective, devices="one", fusable=False)
    prec = c.f64(prec)
    if prec.ndim != 2:
        raise ValueError("expected prec [C, T]")
    Cc, T = prec.shape
    P, N = _multi_common(c, params, 3, qobs, Cc, T)
    ini = np.ascontiguousarray(np.broadcast_to(np.asarray(inits, dtype=np.float64).reshape(-1), (Cc,)))
    c.keep.append(ini)
    out = out or {}
    q = c.empty((Cc, T, N), out.get("qsim")) if _want(want_qsim) else None
    st = c.empty((Cc, T, N), out.get("storage")) if return_storage else None
    if Cc > 0 and T > 0 and N > 0:
        _lib.check(_lib.lib().rrb_abc_simulate_multi(_lib.ptr(prec), Cc, T, _lib.ptr(ini), _lib.ptr(P), N, _lib.ptr(q),
                                                     _lib.ptr(st), C.byref(c.opts)))
    return _result(c, ["qsim", "storage"], [q, st])


def cemaneige_multi(prec, mean_temp, frac_solid, inits, params, return_storages=False, qobs=None, want_outflow=True,
                    math=DEFAULT_MATH, device=None, block=0, out=None, objective="mse"):
    """Cemaneige snow routine for C independent catchments with N members each, one launch.

    prec, mean_temp, frac_solid: [C, T, L]; inits: (2,) or [C, 2] = (snow_pack_init, thermal_state_init);
    params: [C, N, k >= 2] whose first two fields are (CTG, Kf).  Returns {'outflow': [C, T, N], 'G', 'eTG':
    [C, T, L, N], 'mse': [C, N]}; bit-identical to looping ``cemaneige`` over the catchments."""
    params = _records_to_matrix(params)
    c = _Call([prec, mean_temp, frac_solid, params], math, device, block, 0, None, objective=objective, devices="one",
              fusable=False)
    prec = c.f64(prec); mean_temp = c.f64(mean_temp, prec.shape); frac_solid = c.f64(frac_solid, prec.shape)
    if prec.ndim != 3:
        raise ValueError("layer arrays must be [C, T, L]")
    Cc, T, L = prec.shape
    width = int(params.shape[2]) if np.ndim(params) == 3 else 2
    if width < 2:
        raise ValueError("Cemaneige parameter records start with (CTG, Kf)")
    P, N = _multi_common(c, params, width, qobs, Cc, T)
    ini = np.ascontiguousarray(np.broadcast_to(np.asarray(inits, dtype=np.float64).reshape(-1, 2), (Cc, 2)))
    c.keep.append(ini)
    out = out or {}
    q = c.empty((Cc, T, N), out.get("outflow")) if _want(want_outflow) else None
    G, E = (c.empty((Cc, T, L, N), out.get("G")), c.empty((Cc, T, L, N), out.get("eTG"))) if return_storages else (None, None)
    if Cc > 0 and T > 0 and N > 0:
        _lib.check(_lib.lib().rrb_cemaneige_simulate_multi(
            _lib.ptr(prec), _lib.ptr(mean_temp), _lib.ptr(frac_solid), Cc, T, L, _lib.ptr(ini), _lib.ptr(P), width, N,
            _lib.ptr(q), _lib.ptr(G), _lib.ptr(E), C.byref(c.opts)))
    return _result(c, ["outflow", "G", "eTG"], [q, G, E])


def snowice_gr4j_multi(hyst, ice, prec, mean_temp, etp, frac_ice, frac_solid, inits, params, return_storages=False,
                       qobs=None, want_qsim=True, math=DEFAULT_MATH, device=None, block=0, out=None, x4_max=0.0,
                       objective="mse"):
    """The snow(+hysteresis)(+ice)+GR4J couplings for C independent catchments with N members each, one launch.

    Layer arrays [C, T, L]; etp [C, T]; frac_ice [C, L] (ice models); inits (4,) / [C, 4] without and (5,) / [C, 5] with
    hysteresis (the order of ``snowice_gr4j``); params [C, N, 7 | 8 | 9].  Bit-identical to looping ``snowice_gr4j``."""
    if not (hyst or ice):
        raise ValueError("use cemaneigegr4j_multi for the plain coupling")
    k = 6 + (2 if hyst else 0) + (1 if ice else 0)
    nin = 5 if hyst else 4
    params = _records_to_matrix(params)
    arrays = [prec, mean_temp, etp, frac_solid, params] + ([frac_ice] if ice else [])
    c = _Call(arrays, math, device, block, 0, None, x4_max, objective=objective, devices="one", fusable=False)
    prec = c.f64(prec); mean_temp = c.f64(mean_temp, prec.shape); frac_solid = c.f64(frac_solid, prec.shape)
    if prec.ndim != 3:
        raise ValueError("layer arrays must be [C, T, L]")
    Cc, T, L = prec.shape
    etp = c.f64(etp, (Cc, T))
    fice = c.f64(frac_ice, (Cc, L)) if ice else None
    P, N = _multi_common(c, params, k, qobs, Cc, T)
    ini = np.ascontiguousarray(np.broadcast_to(np.asarray(inits, dtype=np.float64).reshape(-1, nin), (Cc, nin)))
    c.keep.append(ini)
    out = out or {}
    e3 = lambda name: c.empty((Cc, T, N), out.get(name))
    e4 = lambda name: c.empty((Cc, T, L, N), out.get(name))
    q = e3("qsim") if _want(want_qsim) else None
    G = E = s = r = sca = im = sm = None
    if return_storages:
        G, E, s, r = e4("G"), e4("eTG"), e3("s_store"), e3("r_store")
        if hyst:
            sca = e4("sca")
        if ice:
            im = e3("icemelt")
        if hyst and ice:
            sm = e3("snowmelt")
    p_, L_ = _lib.ptr, _lib.lib()
    if Cc > 0 and T > 0 and N > 0:
        if hyst and ice:
            rc = L_.rrb_cemaneigehystgr4jice_simulate_multi(p_(prec), p_(mean_temp), p_(etp), p_(fice), p_(frac_solid), Cc, T, L,
                                                            p_(ini), p_(P), N, p_(q), p_(G), p_(E), p_(s), p_(r), p_(sca), p_(im),
                                                            p_(sm), C.byref(c.opts))
        elif hyst:
            rc = L_.rrb_cemaneigehystgr4j_simulate_multi(p_(prec), p_(mean_temp), p_(etp), p_(frac_solid), Cc, T, L, p_(ini),
                                                         p_(P), N, p_(q), p_(G), p_(E), p_(s), p_(r), p_(sca), C.byref(c.opts))
        else:
            rc = L_.rrb_cemaneigegr4jice_simulate_multi(p_(prec), p_(mean_temp), p_(etp), p_(fice), p_(frac_solid), Cc, T, L,
                                                        p_(ini), p_(P), N, p_(q), p_(G), p_(E), p_(s), p_(r), p_(im),
                                                        C.byref(c.opts))
        _lib.check(rc)
    return _result(c, ["qsim", "G", "eTG", "s_store", "r_store", "sca", "icemelt", "snowmelt"], [q, G, E, s, r, sca, im, sm])


def snow_layers(prec, mean_temp, min_temp, max_temp, met_station_height, altitudes=(), device=None):
    """Layer preprocessing of the Cemaneige family on the GPU: station series ``[T]`` -> ``(layer_prec,
    layer_mean_temp, frac_solid)``, each ``[T, L]`` -- what ``extrapolate_precipitation``,
    ``extrapolate_temperature`` and ``calculate_solid_fraction`` (rrmpg/models/cemaneige_utils.py:16-208) do on
    the host, bit-identical (the per-layer ``exp`` gradient is evaluated here with libm like numba does).
    numpy in -> numpy out, torch CUDA tensors in -> tensors out (no host round trip of the series).
    ``altitudes`` empty: one layer at the station height, series passed through (cemaneige.py:209-217)."""
    import math
    alts = np.asarray(altitudes, dtype=np.float64).reshape(-1)
    h = float(met_station_height)
    if alts.size == 0:
        z, flags0 = np.array([h]), 0
    else:
        z, flags0 = alts, _lib.LAYER_SHIFT_TEMP
    L = z.size
    fac, dt, fl = np.ones(L), np.zeros(L), np.zeros(L, dtype=np.int32)
    for l in range(L):
        f = flags0
        if alts.size:
            if z[l] <= 4000:                                   # cemaneige_utils.py:136-138
                fac[l] = math.exp(float((z[l] - h) * 0.0004)); f |= _lib.LAYER_SCALE_PREC
            elif h <= 4000:                                    # :142-144
                fac[l] = math.exp(float((4000 - h) * 0.0004)); f |= _lib.LAYER_SCALE_PREC
            dt[l] = (z[l] - h) * -0.0065                       # :201
        if not (z[l] < 1500):                                  # :58
            f |= _lib.LAYER_HIGH
        fl[l] = f
    c = _Call([prec, mean_temp, min_temp, max_temp], DEFAULT_MATH, device, 0, 0, None, devices="one", fusable=False)
    prec = c.f64(prec); mean_temp = c.f64(mean_temp, prec.shape); min_temp = c.f64(min_temp, prec.shape)
    max_temp = c.f64(max_temp, prec.shape)
    if prec.ndim != 1:
        raise ValueError("station series must be one-dimensional")
    T = prec.shape[0]
    outs = [c.empty((T, L)) for _ in range(3)]
    if T > 0:
        _lib.check(_lib.lib().rrb_snow_layers(_lib.ptr(prec), _lib.ptr(mean_temp), _lib.ptr(min_temp), _lib.ptr(max_temp),
                                              T, L, _lib.ptr(fac), _lib.ptr(dt), _lib.ptr(fl), *[_lib.ptr(o) for o in outs],
                                              C.byref(c.opts)))
    return tuple(outs)
