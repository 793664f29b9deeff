"""CTA-size sweep of the GR4J / Cemaneige kernels for one ensemble size (development aid): is ONE CTA per SM -- warps spread
evenly over the four sub-partitions -- better than the hardware's placement of small CTAs?  usage: time_blocks.py N"""
import sys, os
import numpy as np, torch
root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, root)
from rrmpg_b200 import engine, synthetic
from rrmpg_b200.models import GR4J, Cemaneige, CemaneigeGR4J
from rrmpg_b200.models import _snow_inputs

dev = torch.device("cuda:0")
T = 14610
N = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
f = synthetic.forcing(T)
t = lambda a, dt=torch.float64: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=dev)

def timeit(fn, reps=4):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

buf = torch.empty((T, N), dtype=torch.float64, device=dev)
out = {"qsim": buf}
prec, etp = t(f["prec"]), t(f["etp"])
one = 32 * -(-(-(-N // 32)) // 148)
P = t(engine.pack_params(synthetic.random_params(GR4J(), N)))
for b in (0, 64, 128, 256, one):
    try:
        ms = timeit(lambda: engine.gr4j(prec, etp, 0.6, 0.7, P, out=out, x4_max=2.9, block=b))
        print(f"GR4J N={N} block={b}: {ms:.3f} ms {N*T/ms/1e6:.1f} G/s", flush=True)
    except Exception as e:
        print(f"GR4J N={N} block={b}: {e}")
lp, lt, fr, L = _snow_inputs.to_layers(f["prec"], f["temp"], f["min_temp"], f["max_temp"], synthetic.MET_STATION_HEIGHT,
                                       np.array(synthetic.ALTITUDES))
lp, lt, fr = t(lp), t(lt), t(fr)
P = t(engine.pack_params(synthetic.random_params(Cemaneige(), N)))
for b in (0, 64, 128, 256, one):
    try:
        ms = timeit(lambda: engine.cemaneige(lp, lt, fr, 0.0, 0.0, P, out={"outflow": buf}, block=b))
        print(f"Cemaneige N={N} block={b}: {ms:.3f} ms {N*T/ms/1e6:.1f} G/s", flush=True)
    except Exception as e:
        print(f"Cemaneige N={N} block={b}: {e}")
P = t(engine.pack_params(synthetic.random_params(CemaneigeGR4J(), N)))
for b in (0, 64, 128, 192, 256, 320):
    try:
        ms = timeit(lambda: engine.cemaneigegr4j(lp, lt, etp, fr, (0, 0, 0.6, 0.7), P, out=out, x4_max=2.9, block=b))
        print(f"CemaneigeGR4J N={N} block={b}: {ms:.3f} ms {N*T/ms/1e6:.1f} G/s", flush=True)
    except Exception as e:
        print(f"CemaneigeGR4J N={N} block={b}: {e}")
