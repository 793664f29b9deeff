"""BASELINE config 5, the share of one GPU of eight: 128 catchments x 4096 members x 87 660 hourly steps of HBV-Edu with the
fused objective (no discharge array), device mode.  usage: config5_share.py [variant ...]   (rrb_opts.variant, 0 = library)"""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", ".."))
import numpy as np, torch
from rrmpg_b200 import engine, synthetic
from rrmpg_b200.models import HBVEdu
dev = torch.device("cuda:0"); C, N, T = 128, 4096, 87660
fs = [synthetic.forcing(T, seed=20260101 + c, hourly=True) for c in range(4)]
st = lambda k, dt=np.float64: np.stack([fs[c % 4][k] for c in range(C)]).astype(dt)
t = lambda a, dt=torch.float64: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=dev)
P = t(np.stack([engine.pack_params(synthetic.random_params(HBVEdu(), N, seed=100 + c)) for c in range(C)]))
temp, prec, month0 = t(st("temp")), t(st("prec")), t(st("month") - 1, torch.int8)
pe, tm = t(st("PE_m")), t(st("T_m"))
qobs = t(np.abs(np.random.default_rng(1).normal(1, 0.5, (C, T))))
ref = {}
for variant in [int(v) for v in sys.argv[1:]] or [0]:
    engine.VARIANT = variant
    for obj in ("mse", "kge"):
        fn = lambda: engine.hbvedu_multi(temp, prec, month0, pe, tm, (0, 100, 3, 10), P, qobs=qobs, want_qsim=False, objective=obj)
        for _ in range(2): r = fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); fn(); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 2
        same = ""
        if obj in ref:
            same = " (bit-identical to the first variant)" if torch.equal(ref[obj], r["mse"]) else " (DIFFERS from the first variant)"
        else:
            ref[obj] = r["mse"].clone()
        print(f"config 5 share, variant {variant}, objective {obj}: {ms:.1f} ms, {C*N*T/ms/1e6:.1f} G member-steps/s{same}", flush=True)
