"""Latency of one small-ensemble call (the fit() regime): host-mode fused-MSE calls with N = 165 members."""
import sys, os, time
import numpy as np
root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, root)
from rrmpg_b200 import engine, synthetic
from rrmpg_b200.models import HBVEdu, GR4J
N = int(sys.argv[1]) if len(sys.argv) > 1 else 165
f = synthetic.forcing(14610)
qobs = np.abs(np.random.default_rng(0).normal(1, 0.3, 14610))
P = engine.pack_params(synthetic.random_params(HBVEdu(), N))
m0 = (f["month"] - 1).astype(np.int8)
fn = lambda: engine.hbvedu(f["temp"], f["prec"], m0, f["PE_m"], f["T_m"], (0, 100, 3, 10), P, qobs=qobs, want_qsim=False)
for _ in range(3): fn()
t0 = time.perf_counter()
for _ in range(20): fn()
print(f"HBVEdu N={N} fused-MSE host call: {(time.perf_counter() - t0) / 20 * 1e3:.3f} ms per call")
Pg = engine.pack_params(synthetic.random_params(GR4J(), N))
fn = lambda: engine.gr4j(f["prec"], f["etp"], 0.6, 0.7, Pg, qobs=qobs, want_qsim=False)
for _ in range(3): fn()
t0 = time.perf_counter()
for _ in range(20): fn()
print(f"GR4J   N={N} fused-MSE host call: {(time.perf_counter() - t0) / 20 * 1e3:.3f} ms per call")
