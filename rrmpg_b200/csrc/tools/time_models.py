"""Device-resident timing of every model kernel (development aid; bench.py is the official harness)."""
import sys, os, time
import numpy as np, torch
root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, root)
from rrmpg_b200 import engine, synthetic
from rrmpg_b200.models import ABCModel, HBVEdu, GR4J, Cemaneige, CemaneigeGR4J
from rrmpg_b200.models import _snow_inputs

dev = torch.device("cuda:0")
T = 14610
N = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
math = sys.argv[2] if len(sys.argv) > 2 else "fast"
f = synthetic.forcing(T)
t = lambda a, dt=torch.float64: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=dev)

def timeit(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

def report(name, ms, nbytes=8):
    print(f"{name:28s} N={N:8d} {ms:8.3f} ms  {N*T/ms/1e6:8.1f} G member-steps/s  {nbytes*N*T/ms/1e6:7.0f} GB/s", flush=True)

buf = torch.empty((T, N), dtype=torch.float64, device=dev)
report("cudaMemset (fill_)", timeit(lambda: buf.fill_(1.0)))
src = torch.empty((T // 2, N), dtype=torch.float64, device=dev)
report("copy_ half (r+w bytes)", timeit(lambda: buf[:T // 2].copy_(src)))
out = {"qsim": buf}
P = t(engine.pack_params(synthetic.random_params(ABCModel(), N)))
prec = t(f["prec"])
report("ABC", timeit(lambda: engine.abc(prec, 0.0, P, out=out, math=math)))
P = t(engine.pack_params(synthetic.random_params(HBVEdu(), N)))
temp, month0, pe, tm = t(f["temp"]), t(f["month"] - 1, torch.int8), t(f["PE_m"]), t(f["T_m"])
report("HBVEdu", timeit(lambda: engine.hbvedu(temp, prec, month0, pe, tm, (0, 100, 3, 10), P, out=out, math=math)))
if N <= 131072:  # all five outputs: 40 B per member-step, 38 GB at 65 536 members
    st = {k: torch.empty((T, N), dtype=torch.float64, device=dev) for k in ("snow", "soil", "s1", "s2")}
    st["qsim"] = buf
    report("HBVEdu + 4 storages (40 B)", timeit(lambda: engine.hbvedu(temp, prec, month0, pe, tm, (0, 100, 3, 10), P, return_storage=True,
                                                                  out=st, math=math)), nbytes=40)
    del st
P = t(engine.pack_params(synthetic.random_params(GR4J(), N)))
etp = t(f["etp"])
report("GR4J", timeit(lambda: engine.gr4j(prec, etp, 0.6, 0.7, P, out=out, math=math, x4_max=2.9)))
Pl = synthetic.random_params(GR4J(), N); Pl["x4"] = np.random.default_rng(3).uniform(0.5, 10.0, N)
P = t(engine.pack_params(Pl))
report("GR4J x4<=10 (long UH class)", timeit(lambda: engine.gr4j(prec, etp, 0.6, 0.7, P, out=out, math=math, x4_max=10.0)))
lp, lt, fr, L = _snow_inputs.to_layers(f["prec"], f["temp"], f["min_temp"], f["max_temp"], synthetic.MET_STATION_HEIGHT,
                                       np.array(synthetic.ALTITUDES))
lp, lt, fr = t(lp), t(lt), t(fr)
P = t(engine.pack_params(synthetic.random_params(Cemaneige(), N)))
out2 = {"outflow": buf}
report("Cemaneige L=5", timeit(lambda: engine.cemaneige(lp, lt, fr, 0.0, 0.0, P, out=out2, math=math)))
P = t(engine.pack_params(synthetic.random_params(CemaneigeGR4J(), N)))
report("CemaneigeGR4J L=5", timeit(lambda: engine.cemaneigegr4j(lp, lt, etp, fr, (0, 0, 0.6, 0.7), P, out=out, math=math, x4_max=2.9)))
# snow-ice couplings
from rrmpg_b200.models import CemaneigeGR4JIce, CemaneigeHystGR4J, CemaneigeHystGR4JIce
fice = t(np.full(L, 0.1))
for name, cls, hyst, ice in (("CemaneigeGR4JIce", CemaneigeGR4JIce, False, True), ("CemaneigeHystGR4J", CemaneigeHystGR4J, True, False),
                             ("CemaneigeHystGR4JIce", CemaneigeHystGR4JIce, True, True)):
    Pm = synthetic.random_params(cls(), N)
    x4 = float(np.max(Pm["x4"]))
    P = t(engine.pack_params(Pm))
    ini = (0, 0, 0, 0.6, 0.7) if hyst else (0, 0, 0.6, 0.7)
    report(f"{name} L=5 x4<={x4:.1f}", timeit(lambda: engine.snowice_gr4j(hyst, ice, lp, lt, etp, fice if ice else None, fr, ini, P,
                                                                       out=out, math=math, x4_max=x4)))
