"""ctypes binding of ``librrmpg_b200.so`` (the C ABI declared in ``include/rrmpg_b200.h``).

The library is the only compute path of this package: there is no CPU or PyTorch fallback.
If the shared object is missing, or no CUDA device is visible when a simulation is requested,
the call raises ``RuntimeError`` instead of silently computing somewhere else.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RRMPG_B200_LIB", os.path.join(_HERE, "librrmpg_b200.so"))  # override: kernel experiments

RRB_OK, RRB_EINVAL, RRB_ECUDA, RRB_EUNSUPPORTED, RRB_ENOMEM = range(5)
MEM_HOST, MEM_DEVICE = 0, 1
MATH_FAST, MATH_PRECISE = 0, 1
OBJ_MSE, OBJ_NSE, OBJ_KGE = 0, 1, 2
MODEL_ABC, MODEL_HBVEDU, MODEL_GR4J = 0, 1, 2
MAX_LAYERS = 16
MAX_X4 = 64.0

_dp = C.POINTER(C.c_double)
_bp = C.POINTER(C.c_int8)


class Opts(C.Structure):
    """``struct rrb_opts`` (include/rrmpg_b200.h)."""
    _fields_ = [("struct_size", C.c_int32), ("device", C.c_int32), ("mem", C.c_int32),
                ("math", C.c_int32), ("stream", C.c_void_p), ("block", C.c_int32),
                ("variant", C.c_int32), ("x4_max", C.c_double), ("qobs", C.c_void_p),
                ("mse", C.c_void_p), ("slab_steps", C.c_int64),
                ("objective", C.c_int32), ("n_devices", C.c_int32), ("devices", C.c_void_p),
                ("obs_stats", C.c_void_p), ("state_in", C.c_void_p), ("state_out", C.c_void_p),
                ("out_row_pitch", C.c_int64)]


_SIGS = {
    "rrb_version": (C.c_int, []),
    "rrb_device_count": (C.c_int, []),
    "rrb_init": (C.c_int, [C.c_int]),
    "rrb_init_devices": (C.c_int, [C.c_void_p, C.c_int]),
    "rrb_state_rows": (C.c_int, [C.c_int, C.c_double]),
    "rrb_shutdown": (C.c_int, []),
    "rrb_last_error": (C.c_char_p, []),
    "rrb_synchronize": (C.c_int, [C.c_int]),
    "rrb_host_alloc": (C.c_void_p, [C.c_size_t]),
    "rrb_host_free": (None, [C.c_void_p]),
    "rrb_host_pool_trim": (None, []),
    "rrb_abc_simulate": (C.c_int, [C.c_void_p, C.c_int64, C.c_double, C.c_void_p, C.c_int64,
                                   C.c_void_p, C.c_void_p, C.POINTER(Opts)]),
    "rrb_hbvedu_simulate": (C.c_int, [C.c_void_p] * 5 + [C.c_int64, C.c_void_p, C.c_void_p, C.c_int64]
                            + [C.c_void_p] * 5 + [C.POINTER(Opts)]),
    "rrb_hbvedu_simulate_multi": (C.c_int, [C.c_void_p] * 5 + [C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64]
                                  + [C.c_void_p] * 5 + [C.POINTER(Opts)]),
    "rrb_gr4j_simulate": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_double, C.c_double,
                                    C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.POINTER(Opts)]),
    "rrb_cemaneige_simulate": (C.c_int, [C.c_void_p] * 3 + [C.c_int64, C.c_int64, C.c_double, C.c_double,
                                                            C.c_void_p, C.c_int64, C.c_int64]
                               + [C.c_void_p] * 3 + [C.POINTER(Opts)]),
    "rrb_cemaneigegr4j_simulate": (C.c_int, [C.c_void_p] * 4 + [C.c_int64, C.c_int64, C.c_void_p,
                                                                C.c_void_p, C.c_int64]
                                   + [C.c_void_p] * 5 + [C.POINTER(Opts)]),
    "rrb_cemaneigegr4jice_simulate": (C.c_int, [C.c_void_p] * 5 + [C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64]
                                      + [C.c_void_p] * 6 + [C.POINTER(Opts)]),
    "rrb_cemaneigehystgr4j_simulate": (C.c_int, [C.c_void_p] * 4 + [C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64]
                                       + [C.c_void_p] * 6 + [C.POINTER(Opts)]),
    "rrb_cemaneigehystgr4jice_simulate": (C.c_int, [C.c_void_p] * 5 + [C.c_int64, C.c_int64, C.c_void_p, C.c_void_p,
                                                                        C.c_int64] + [C.c_void_p] * 8 + [C.POINTER(Opts)]),
    "rrb_gr4j_simulate_multi": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64]
                                + [C.c_void_p] * 3 + [C.POINTER(Opts)]),
    "rrb_cemaneigegr4j_simulate_multi": (C.c_int, [C.c_void_p] * 4 + [C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p,
                                                                      C.c_int64] + [C.c_void_p] * 5 + [C.POINTER(Opts)]),
    "rrb_abc_simulate_multi": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                                         C.c_void_p, C.POINTER(Opts)]),
    "rrb_cemaneige_simulate_multi": (C.c_int, [C.c_void_p] * 3 + [C.c_int64] * 3 + [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64]
                                     + [C.c_void_p] * 3 + [C.POINTER(Opts)]),
    "rrb_cemaneigegr4jice_simulate_multi": (C.c_int, [C.c_void_p] * 5 + [C.c_int64] * 3 + [C.c_void_p, C.c_void_p, C.c_int64]
                                            + [C.c_void_p] * 6 + [C.POINTER(Opts)]),
    "rrb_cemaneigehystgr4j_simulate_multi": (C.c_int, [C.c_void_p] * 4 + [C.c_int64] * 3 + [C.c_void_p, C.c_void_p, C.c_int64]
                                             + [C.c_void_p] * 6 + [C.POINTER(Opts)]),
    "rrb_cemaneigehystgr4jice_simulate_multi": (C.c_int, [C.c_void_p] * 5 + [C.c_int64] * 3 + [C.c_void_p, C.c_void_p, C.c_int64]
                                                + [C.c_void_p] * 8 + [C.POINTER(Opts)]),
    "rrb_snow_layers": (C.c_int, [C.c_void_p] * 4 + [C.c_int64, C.c_int64] + [C.c_void_p] * 6 + [C.POINTER(Opts)]),
    "rrb_host_fast_pow": (None, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "rrb_host_fast_exp2m1": (None, [C.c_void_p, C.c_int64, C.c_void_p]),
    "rrb_host_hbv_pow_step": (None, [C.c_void_p] * 5 + [C.c_int64, C.c_void_p, C.c_void_p]),
}

_lib = None


def lib():
    """Load the shared library (once).  Raises RuntimeError when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C rrmpg_b200/csrc`.  rrmpg_b200 has no CPU fallback.")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(handle, name)  # AttributeError if the .so does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def exported_symbols():
    return sorted(_SIGS)


def last_error():
    return lib().rrb_last_error().decode("utf-8", "replace")


def check(rc):
    if rc != RRB_OK:
        msg = last_error()
        if rc == RRB_EINVAL:
            raise ValueError(f"rrmpg_b200: {msg}")
        if rc == RRB_ENOMEM:
            raise MemoryError(f"rrmpg_b200: {msg}")
        raise RuntimeError(f"rrmpg_b200: {msg}")


def device_count():
    return int(lib().rrb_device_count())


def require_gpu():
    if device_count() < 1:
        raise RuntimeError("rrmpg_b200: no CUDA device visible; this engine has no CPU fallback")


# ------------------------------------------------------------------------------------------
# pinned host arrays (so the D2H of multi-GB discharge arrays runs at PCIe speed)
# ------------------------------------------------------------------------------------------
class _PinnedBlock:
    def __init__(self, nbytes):
        self._lib = lib()
        self.ptr = self._lib.rrb_host_alloc(nbytes)
        if not self.ptr:
            raise MemoryError(last_error())
        self.__array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (self.ptr, False),
                                    "version": 3}

    def __del__(self):
        try:
            self._lib.rrb_host_free(self.ptr)
        except Exception:  # interpreter shutdown
            pass


PINNED_MIN_BYTES = 1 << 20
# include/rrmpg_b200.h: RRB_LAYER_*
LAYER_SCALE_PREC, LAYER_SHIFT_TEMP, LAYER_HIGH = 1, 2, 4


def host_empty(shape, dtype=np.float64):
    """Uninitialised C-order host array; pinned (pooled cudaHostAlloc) when it is large."""
    shape = tuple(int(s) for s in np.atleast_1d(shape))
    nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
    if nbytes >= PINNED_MIN_BYTES:
        try:
            block = _PinnedBlock(nbytes)
            return np.asarray(block).view(dtype).reshape(shape)
        except MemoryError:
            pass
    return np.empty(shape, dtype)


def ptr(a):
    """Raw address of a numpy array / torch tensor / None."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    return a.data_ptr()  # torch tensor
