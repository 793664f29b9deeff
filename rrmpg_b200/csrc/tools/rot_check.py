"""hbv_rot_kernel (rrb_opts.variant = 3) against hbv_fast2_kernel<2> (variant 2): bit-identical discharge, objectives and
carried states over every shape of the schedule (RRMPG_B200_HBV_ROT_SMS pretends a small GPU).  Development tool."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", ".."))
from rrmpg_b200 import engine, synthetic
from rrmpg_b200.models import HBVEdu

def case(T, N, seed):
    f = synthetic.forcing(T)
    np.random.seed(seed)
    P = HBVEdu().get_random_params(N)
    return f, P

def run(variant, *a, **k):
    engine.VARIANT = variant
    try:
        return engine.hbvedu(*a, **k)
    finally:
        engine.VARIANT = 0

bad = 0
for sms, pairs_per_cta in ((4, 7), (4, 6), (4, 5), (3, 3), (2, 8), (2, 1), (4, 11), (2, 14), (2, 16), (3, 13)):
    os.environ["RRMPG_B200_HBV_ROT_SMS"] = str(sms)
    N = 64 * pairs_per_cta * sms - 6          # ragged last pair
    T = 3001
    f, P = case(T, N, seed=sms * 100 + pairs_per_cta)
    qobs = np.abs(np.random.default_rng(5).normal(2.0, 1.0, T))
    args = (f["temp"], f["prec"], f["month"] - 1, f["PE_m"], f["T_m"], (2.0, 100, 3, 10), P)
    for kw in (dict(), dict(qobs=qobs), dict(qobs=qobs, objective="kge"), dict(qobs=qobs, want_qsim=False),
               dict(slab_steps=1100), dict(qobs=qobs, objective="nse", slab_steps=777)):
        a = run(2, *args, **kw)
        for v in (3, 5):
            b = run(v, *args, **kw)
            for k in a:
                if a[k] is None:
                    continue
                same = np.array_equal(np.asarray(a[k]).view(np.int64), np.asarray(b[k]).view(np.int64))
                if not same:
                    bad += 1
                    d = np.abs(np.asarray(a[k]) - np.asarray(b[k]))
                    print(f"MISMATCH variant={v} sms={sms} P={pairs_per_cta} {kw.keys()} {k}: max|d|={np.nanmax(d):.3e} at {np.unravel_index(np.nanargmax(d), d.shape)}")
    print(f"sms={sms} pairs/CTA={pairs_per_cta} N={N}: checked", flush=True)
print("rot_check:", "FAILED" if bad else "all bit-identical")
sys.exit(1 if bad else 0)
