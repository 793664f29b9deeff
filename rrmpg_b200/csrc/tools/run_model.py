"""Run one model's device-mode kernel a few times (for ncu).  usage: run_model.py <model> [N] [math]"""
import sys, os
import numpy as np, torch
root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, root)
from rrmpg_b200 import engine, synthetic
from rrmpg_b200.models import ABCModel, HBVEdu, GR4J, Cemaneige, CemaneigeGR4J, _snow_inputs
model = sys.argv[1]; N = int(sys.argv[2]) if len(sys.argv) > 2 else 65536; math = sys.argv[3] if len(sys.argv) > 3 else "fast"
dev = torch.device("cuda:0"); T = 14610
f = synthetic.forcing(T)
t = lambda a, dt=torch.float64: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=dev)
buf = torch.empty((T, N), dtype=torch.float64, device=dev)
prec, etp = t(f["prec"]), t(f["etp"])
if model == "gr4j":
    P = t(engine.pack_params(synthetic.random_params(GR4J(), N)))
    fn = lambda: engine.gr4j(prec, etp, 0.6, 0.7, P, out={"qsim": buf}, math=math, x4_max=2.9)
elif model in ("cemaneige", "cemaneigegr4j"):
    lp, lt, fr, L = _snow_inputs.to_layers(f["prec"], f["temp"], f["min_temp"], f["max_temp"], synthetic.MET_STATION_HEIGHT,
                                           np.array(synthetic.ALTITUDES))
    lp, lt, fr = t(lp), t(lt), t(fr)
    if model == "cemaneige":
        P = t(engine.pack_params(synthetic.random_params(Cemaneige(), N)))
        fn = lambda: engine.cemaneige(lp, lt, fr, 0.0, 0.0, P, out={"outflow": buf}, math=math)
    else:
        P = t(engine.pack_params(synthetic.random_params(CemaneigeGR4J(), N)))
        fn = lambda: engine.cemaneigegr4j(lp, lt, etp, fr, (0, 0, 0.6, 0.7), P, out={"qsim": buf}, math=math, x4_max=2.9)
elif model == "hyst":
    from rrmpg_b200.models import CemaneigeHystGR4J
    lp, lt, fr, L = _snow_inputs.to_layers(f["prec"], f["temp"], f["min_temp"], f["max_temp"], synthetic.MET_STATION_HEIGHT,
                                           np.array(synthetic.ALTITUDES))
    lp, lt, fr = t(lp), t(lt), t(fr)
    Pm = synthetic.random_params(CemaneigeHystGR4J(), N)
    P = t(engine.pack_params(Pm))
    fn = lambda: engine.snowice_gr4j(True, False, lp, lt, etp, None, fr, (0, 0, 0, 0.6, 0.7), P, out={"qsim": buf}, math=math,
                                     x4_max=float(Pm["x4"].max()))
elif model == "abc":
    P = t(engine.pack_params(synthetic.random_params(ABCModel(), N)))
    fn = lambda: engine.abc(prec, 0.0, P, out={"qsim": buf}, math=math)
else:
    P = t(engine.pack_params(synthetic.random_params(HBVEdu(), N)))
    temp, month0, pe, tm = t(f["temp"]), t(f["month"] - 1, torch.int8), t(f["PE_m"]), t(f["T_m"])
    fn = lambda: engine.hbvedu(temp, prec, month0, pe, tm, (0, 100, 3, 10), P, out={"qsim": buf}, math=math)
for _ in range(3):
    fn()
torch.cuda.synchronize()
