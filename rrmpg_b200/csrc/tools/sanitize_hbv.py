"""Small HBV-Edu launches over every FAST launch shape (one / two members per thread, one CTA per SM, capped and uncapped
build, hbv_rot_kernel, fused objectives, time slabs, a flagged member) -- the workload for compute-sanitizer.
usage: compute-sanitizer --tool {memcheck,racecheck,synccheck} python sanitize_hbv.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", ".."))
os.environ["RRMPG_B200_HBV_ROT_SMS"] = "3"
from rrmpg_b200 import engine, synthetic
from rrmpg_b200.models import HBVEdu

T = 301
f = synthetic.forcing(T)
qobs = np.abs(np.random.default_rng(5).normal(2.0, 1.0, T))
for pairs in (7, 5, 13):
    N = 64 * pairs * 3 - 6
    np.random.seed(pairs)
    P = HBVEdu().get_random_params(N)
    P["FC"][5] = 1e-200   # outside the FAST contract: its block goes to the PRECISE kernel
    args = (f["temp"], f["prec"], f["month"] - 1, f["PE_m"], f["T_m"], (2.0, 100, 3, 10), P)
    for v in (0, 1, 2, 3, 5):
        engine.VARIANT = v
        for kw in (dict(), dict(qobs=qobs, objective="kge", want_qsim=False), dict(qobs=qobs, slab_steps=97)):
            with np.errstate(all="ignore"):
                r = engine.hbvedu(*args, **kw)
            assert all(np.isfinite(np.delete(np.asarray(x), 5, axis=-1)).all() for x in r.values() if x is not None)
    print("pairs per CTA", pairs, "ok", flush=True)
engine.VARIANT = 0
