"""BASELINE.json configs 3, 4 and 5 at their stated sizes (bench.py measures configs[1]; these are the other
GPU configurations: full-size parity spot checks against the oracle plus timings kept under profiles/).
Test infrastructure (it uses the oracle as the checker), not collected by pytest; run on the GPU box:

    python tests/baseline_configs.py --config 3                                  # GR4J, 1M members, 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port P \
           tests/baseline_configs.py --config 4 5      # CemaneigeGR4J 256k members / HBVEdu 1024 x 4096 hourly, sharded

One process per GPU.  Config 4 shards the ensemble by contiguous member block, config 5 by catchment block; rank 0
owns the forcing and sends it once (broadcast / scatter); no collective touches the results.  Timing: CUDA events
around the device-mode library call (forcing pack + ensemble kernel), max over ranks, after two warm-up calls.
Prints one JSON line per configuration on rank 0.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "member-timesteps/sec"


def _events(torch, fn, reps):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def _peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"])
    return 6650.0


def config3(torch, rdist, dev, rank, world, reps):
    from rrmpg_b200 import engine, synthetic
    from rrmpg_b200.models import GR4J
    import oracle
    T, N = synthetic.T_DAILY_40Y, 1048576
    f = synthetic.forcing(T)
    P = synthetic.random_params(GR4J(), N)
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=dev)
    prec, etp, dP = t(f["prec"]), t(f["etp"]), t(engine.pack_params(P))
    out = {"qsim": torch.empty((T, N), dtype=torch.float64, device=dev)}       # 122.6 GB, device resident
    ms = _events(torch, lambda: engine.gr4j(prec, etp, 0.6, 0.7, dP, out=out, x4_max=2.9), reps)
    idx = np.r_[0:4, N - 4:N]
    ref = oracle.gr4j(f["prec"], f["etp"], 0.6, 0.7, P[idx])
    ok = bool(np.allclose(out["qsim"][:, torch.as_tensor(idx, device=dev)].cpu().numpy(), ref, rtol=1e-10, atol=1e-12))
    return {"config": "GR4J 1M-member ensemble, 40-year daily, 1xB200 (BASELINE.json configs[2])", "members": N,
            "timesteps": T, "ms": ms, "value": N * T / ms * 1e3, "bytes_per_member_step": 8,
            "roofline_frac": 8 * N * T / ms * 1e3 / 1e9 / _peak(), "parity_spot_check": ok}


def config4(torch, rdist, dev, rank, world, reps):
    from rrmpg_b200 import engine, synthetic
    from rrmpg_b200.models import CemaneigeGR4J, _snow_inputs
    import oracle
    T, N = synthetic.T_DAILY_40Y, 262144
    series = None
    if rank == 0:  # rank 0 owns the station series and does the member-independent layer preprocessing
        f = synthetic.forcing(T)
        lp, lt, fr, _ = _snow_inputs.to_layers(f["prec"], f["temp"], f["min_temp"], f["max_temp"],
                                               synthetic.MET_STATION_HEIGHT, np.array(synthetic.ALTITUDES))
        series = {"layer_prec": lp, "layer_temp": lt, "frac_solid": fr, "etp": f["etp"]}
    s = rdist.broadcast_series(series, src=0, device=dev)                        # the path's one collective
    P = synthetic.random_params(CemaneigeGR4J(), N)                              # same seeded global draw on every rank
    lo, hi = rdist.member_block(N, rank, world)
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=dev)
    lp, lt, fr, etp, dP = t(s["layer_prec"]), t(s["layer_temp"]), t(s["frac_solid"]), t(s["etp"]), t(engine.pack_params(P[lo:hi]))
    out = {"qsim": torch.empty((T, hi - lo), dtype=torch.float64, device=dev)}
    ms = _events(torch, lambda: engine.cemaneigegr4j(lp, lt, etp, fr, (0, 0, 0.6, 0.7), dP, out=out, x4_max=2.9), reps)
    ms = rdist.max_over_ranks(ms, dev)
    idx = np.r_[0:4, (hi - lo) - 4:(hi - lo)]
    ref = oracle.cemaneigegr4j(s["layer_prec"], s["layer_temp"], s["etp"], s["frac_solid"], (0, 0, 0.6, 0.7), P[lo:hi][idx])
    ok = bool(np.allclose(out["qsim"][:, torch.as_tensor(idx, device=dev)].cpu().numpy(), ref, rtol=1e-10, atol=1e-12))
    ok = rdist.max_over_ranks(0.0 if ok else 1.0, dev) == 0.0
    return {"config": f"CemaneigeGR4J 262144 members, 5 layers, 40-year daily, member blocks over {world} GPU(s) "
                      "(BASELINE.json configs[3])", "members": N, "members_per_gpu": hi - lo, "timesteps": T, "ms": ms,
            "value": N * T / ms * 1e3, "bytes_per_member_step": 8,
            "roofline_frac": 8 * (hi - lo) * T / ms * 1e3 / 1e9 / _peak(), "parity_spot_check": ok}


def config5(torch, rdist, dev, rank, world, reps):
    import torch.distributed as dist
    from rrmpg_b200 import engine, synthetic
    from rrmpg_b200.models import HBVEdu
    import oracle
    T, Cn, N = synthetic.T_HOURLY_10Y, 1024, 4096
    c_lo, c_hi = rdist.member_block(Cn, rank, world)                             # contiguous catchment block
    nc = c_hi - c_lo

    def scatter(make, dtype, width):
        """rank 0 builds the [Cn, width] array; every rank receives its catchment rows (one scatter)."""
        if world == 1:
            return torch.as_tensor(make(), device=dev)
        mine = torch.empty((nc, width), dtype=dtype, device=dev)
        parts = None
        if rank == 0:
            full = torch.as_tensor(make(), device=dev)
            parts = [full[slice(*rdist.member_block(Cn, r, world))].contiguous() for r in range(world)]
        assert Cn % world == 0, "equal catchment blocks"
        dist.scatter(mine, parts, src=0)
        return mine

    fs = [synthetic.forcing(T, seed=synthetic.SEED + c, hourly=True) for c in range(Cn)] if rank == 0 else None
    temp = scatter(lambda: np.stack([f["temp"] for f in fs]), torch.float64, T)
    prec = scatter(lambda: np.stack([f["prec"] for f in fs]), torch.float64, T)
    m0 = scatter(lambda: np.stack([f["month"] - 1 for f in fs]).astype(np.int8), torch.int8, T)
    qobs = scatter(lambda: np.abs(np.random.default_rng(11).normal(0.05, 0.02, (Cn, T))), torch.float64, T)
    PE = scatter(lambda: np.stack([f["PE_m"] / 24.0 for f in fs]), torch.float64, 12)
    TM = scatter(lambda: np.stack([f["T_m"] for f in fs]), torch.float64, 12)
    del fs
    P = np.stack([engine.pack_params(synthetic.random_params(HBVEdu(), N, seed=1000 + c)) for c in range(c_lo, c_hi)])
    dP = torch.as_tensor(P, device=dev)
    call = lambda: engine.hbvedu_multi(temp, prec, m0, PE, TM, (0, 100, 3, 10), dP, qobs=qobs, want_qsim=False)
    ms = rdist.max_over_ranks(_events(torch, call, reps), dev)
    mse = call()["mse"].cpu().numpy()
    c = nc - 1
    idx = np.r_[0, N - 1, 17, 2048]
    ref = oracle.hbvedu(temp[c].cpu().numpy(), prec[c].cpu().numpy(), m0[c].cpu().numpy(), PE[c].cpu().numpy(),
                        TM[c].cpu().numpy(), (0, 100, 3, 10), P[c][idx])
    ref_mse = ((qobs[c].cpu().numpy()[:, None] - ref) ** 2).mean(axis=0)
    ok = bool(np.allclose(mse[c, idx], ref_mse, rtol=1e-9)) and bool(np.isfinite(mse).all())
    ok = rdist.max_over_ranks(0.0 if ok else 1.0, dev) == 0.0
    return {"config": f"HBVEdu {Cn} catchments x {N} members, 10-year hourly, catchment blocks over {world} GPU(s), fused "
                      "per-member MSE instead of the 2.94 TB discharge array (BASELINE.json configs[4])",
            "catchments": Cn, "catchments_per_gpu": nc, "members_per_catchment": N, "timesteps": T, "ms": ms,
            "value": Cn * N * T / ms * 1e3, "bytes_per_member_step": 0, "roofline_frac": None, "parity_spot_check": ok}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, nargs="+", default=[4, 5])
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    import torch
    from rrmpg_b200 import _lib, distributed as rdist
    rank, local_rank, world = rdist.init_process_group()
    _lib.require_gpu()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    runners = {3: config3, 4: config4, 5: config5}
    for c in args.config:
        if c == 3 and world > 1:
            continue  # a single-GPU configuration
        r = runners[c](torch, rdist, dev, rank, world, args.reps)
        r.update(metric=METRIC, unit="member-timesteps/s", n_gpus=world, math="fast", data="synthetic")
        if rank == 0:
            print(json.dumps(r), flush=True)
        rdist.barrier()
    rdist.shutdown()
    return 0


if __name__ == "__main__":
    sys.exit(main())
