"""Monte-Carlo calibration of HBV-Edu with the drop-in classes -- the workflow of the reference's tutorial
(docs/source/examples/model_api_example.rst in kratzert/RRMPG) with `rrmpg` replaced by `rrmpg_b200`.

    python examples/monte_carlo_hbvedu.py [num_members]

Needs a B200 (there is no CPU fallback).  Forcing is the seeded synthetic 40-year daily series of the benchmark;
"observations" are the discharge of a hidden parameter set plus noise.
"""
import sys
import time
import os

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from rrmpg_b200 import synthetic                      # noqa: E402
from rrmpg_b200.models import HBVEdu                  # noqa: E402  (was: from rrmpg.models import HBVEdu)
from rrmpg_b200.tools import monte_carlo              # noqa: E402  (was: from rrmpg.tools import monte_carlo)
from rrmpg_b200.utils import calc_nse                 # noqa: E402

num = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
f = synthetic.forcing(synthetic.T_DAILY_40Y)
forcing = dict(temp=f["temp"], prec=f["prec"], month=f["month"], PE_m=f["PE_m"], T_m=f["T_m"], soil_init=100, s1_init=3, s2_init=10)

# synthetic truth
np.random.seed(7)
truth = HBVEdu()
truth.set_params(truth.get_random_params(1)[0])
qobs = truth.simulate(**forcing).ravel() * np.random.default_rng(1).normal(1.0, 0.05, f["prec"].shape[0])

# 1. Monte-Carlo search: one launch over the sampled parameter matrix, per-member MSE fused into the kernel
model = HBVEdu()
t0 = time.perf_counter()
res = monte_carlo(model, num=num, qobs=qobs, **forcing)
dt = time.perf_counter() - t0
best = int(np.nanargmin(res["mse"]))
print(f"monte_carlo: {num} members x {qobs.size} steps in {dt:.2f} s ({num * qobs.size / dt / 1e9:.2f} G member-steps/s end to end, "
      f"qsim {res['qsim'].nbytes / 1e9:.1f} GB returned); best MSE {res['mse'][best]:.4f}, "
      f"NSE {calc_nse(qobs, res['qsim'][:, best]):.4f}")

# 2. refine with the population-vectorised differential evolution (one launch per generation)
t0 = time.perf_counter()
fit = model.fit(qobs, f["temp"], f["prec"], f["month"], f["PE_m"], f["T_m"], soil_init=100, s1_init=3, s2_init=10)
print(f"fit: {fit.nfev} model evaluations in {fit.nit} generations, {time.perf_counter() - t0:.2f} s, MSE {fit.fun:.4f}")
model.set_params(dict(zip(model.get_parameter_names(), fit.x)))
print(f"NSE of the calibrated model: {calc_nse(qobs, model.simulate(**forcing).ravel()):.4f}")
