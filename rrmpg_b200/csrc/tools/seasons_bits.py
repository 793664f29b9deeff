"""Are the season short cuts of hbv_fast2_loop bit-neutral?  Runs a set of HBV-Edu cases with the library named by
RRMPG_B200_LIB and stores the raw results; run it once per build and compare the files (development aid).
usage: RRMPG_B200_LIB=<.so> seasons_bits.py out.npz      |      seasons_bits.py --compare a.npz b.npz"""
import os, sys
import numpy as np
if sys.argv[1] == "--compare":
    a, b = np.load(sys.argv[2]), np.load(sys.argv[3])
    bad = [k for k in a.files if not np.array_equal(a[k].view(np.int64), b[k].view(np.int64))]
    print(f"{len(a.files)} arrays compared:", "all bit-identical" if not bad else f"DIFFER: {bad}")
    sys.exit(1 if bad else 0)
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", ".."))
from rrmpg_b200 import engine, synthetic
from rrmpg_b200.models import HBVEdu
out = {}
for case, (T, N, hourly, seed) in enumerate([(14610, 4096, False, 1), (8760, 2048, True, 2), (3001, 1000, False, 3)]):
    f = synthetic.forcing(T, seed=synthetic.SEED + case, hourly=hourly)
    np.random.seed(seed)
    P = HBVEdu().get_random_params(N)
    if case == 2:                      # wide thresholds, a member without degree-day melt, negative-zero rain and temperatures
        P["T_t"] = np.random.default_rng(7).uniform(-6, 6, N)
        P["DD"][::7] = 0.0
        f["prec"][::5] = -0.0
        f["temp"][::11] = -0.0
    qobs = np.abs(np.random.default_rng(5).normal(2.0, 1.0, T))
    args = (f["temp"], f["prec"], f["month"] - 1, f["PE_m"], f["T_m"], (5.0 * (case == 1), 100, 3, 10), P)
    for v in (1, 2):
        engine.VARIANT = v
        r = engine.hbvedu(*args, return_storage=True, qobs=qobs, objective="kge")
        for k, x in r.items():
            out[f"case{case}_v{v}_{k}"] = np.asarray(x)
np.savez(sys.argv[1], **out)
print("saved", len(out), "arrays")
