#!/usr/bin/env python
"""bench.py -- member-timesteps/sec of the HBV-Edu ensemble hot path on N B200s.

    python bench.py --gpus 1 --steps 5 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
           --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the CPU reference arm (oracle port, all host threads)

Workload (BASELINE.json configs[1]): HBVEdu, 65 536 ensemble members per GPU, 40-year daily
synthetic forcing (T = 14 610), parameters uniform in the model's default bounds.  Weak scaling:
every rank owns one contiguous member block of the global ensemble; rank 0 broadcasts the forcing
once per step (NCCL), there is no other communication.

One "step" = one pass of the hot path over the whole ensemble:
  value : inputs resident in HBM; pack forcing + ensemble kernel through the C ABI in device mode
          (torch tensors), timed with CUDA events on the launching stream, max over ranks.
  e2e   : the call a user makes -- HBVEdu.simulate(numpy in, numpy out): H2D of forcing+params
          and D2H of the full [T, N] discharge inside the timed region.
One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "member-timesteps/sec"
UNIT = "member-timesteps/s"
T_STEPS = 14610
MEMBERS_PER_GPU = 65536
BYTES_PER_MEMBER_STEP = 8  # one fp64 qsim store; forcing/params amortise to ~0
CPU_SAMPLE_MEMBERS = 16384
# kernels of ours per device-mode step: forcing pack + ensemble kernel (+ in FAST mode the PRECISE kernel queued
# behind it as the fallback for non-finite rain, which exits at once on this workload)
LAUNCHES_PER_STEP = {"fast": 3, "precise": 2}
# The binding roof of the HBV kernel is instruction issue, not HBM: per member-timestep it
# executes 61.6 warp-instructions of which ~30 are fp64 (ncu, profiles/r01_ncu_full_hbv_v9_summary.txt), an fp64 warp
# instruction holds a sub-partition's issue port for 2 cycles (16 fp64 lanes), every other one for 1.
HBV_FAST_WARP_INSTR, HBV_FAST_FP64_INSTR = 61.6, 30.0
SM_COUNT, SUBPARTITIONS = 148, 4


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.lines, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 6:
                continue
            try:
                clk, mxv = float(parts[0]), float(parts[1])
            except ValueError:
                continue
            mx = mxv
            if t0 - 0.05 <= ts <= t1 + 0.05:
                sm.append(clk)
                for n, v in zip(names, parts[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def workload(n_members_total):
    from rrmpg_b200 import synthetic
    from rrmpg_b200.models import HBVEdu
    f = synthetic.forcing(T_STEPS)
    P = synthetic.random_params(HBVEdu(), n_members_total)
    return f, P


def cpu_baseline_run(f, P, reps=3):
    """Oracle port (C restatement of run_hbvedu) on all host threads, bounded member sample."""
    import oracle
    n = min(CPU_SAMPLE_MEMBERS, P.shape[0])
    Ps = oracle.pack_params(P[:n])
    month0 = (f["month"] - 1).astype(np.int8)
    cores = oracle.num_threads()
    best = float("inf")
    for _ in range(reps + 1):  # first repetition warms page tables / the thread pool
        t0 = time.perf_counter()
        oracle.hbvedu(f["temp"], f["prec"], month0, f["PE_m"], f["T_m"], (0, 100, 3, 10), Ps, nthreads=cores)
        dt = time.perf_counter() - t0
        best = min(best, dt)
    return {"value": n * T_STEPS / best, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{n} of {MEMBERS_PER_GPU} members x {T_STEPS} steps, oracle/rr_oracle.c (C restatement of "
                      f"rrmpg run_hbvedu), {cores} pthreads, best of {reps}"}, best


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (oracle port; the Python+numba
    reference cannot travel to the GPU box), all host threads, same metric/config as our arm."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return 0
    f, P = workload(CPU_SAMPLE_MEMBERS)
    import oracle
    cores = oracle.num_threads()
    Ps = oracle.pack_params(P)
    month0 = (f["month"] - 1).astype(np.int8)

    def step():
        oracle.hbvedu(f["temp"], f["prec"], month0, f["PE_m"], f["T_m"], (0, 100, 3, 10), Ps, nthreads=cores)

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = CPU_SAMPLE_MEMBERS * T_STEPS * args.steps / dt
    sample = (f"each step = {CPU_SAMPLE_MEMBERS} members x {T_STEPS} steps (bounded sample of the "
              f"{MEMBERS_PER_GPU}-member workload; member-timesteps/s does not depend on N)")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(args.gpus),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


def config_dict(n_gpus, members=MEMBERS_PER_GPU):
    which = "BASELINE.json configs[1]" if members == MEMBERS_PER_GPU else "configs[1] at a non-default ensemble size"
    return {"workload": f"HBVEdu {members} members per GPU x {T_STEPS} daily steps ({which}), "
                        "qsim-only output [T, N] fp64",
            "members_per_gpu": members, "timesteps": T_STEPS, "global_members": members * n_gpus,
            "parallelism": f"member-block x{n_gpus}" if n_gpus > 1 else "single GPU",
            "math": "fast",
            "l2": f"no explicit flush: every step streams its {members * T_STEPS * 8 / 1e9:.2f} GB discharge array through the 126 MB L2"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--math", default="fast", choices=["fast", "precise"])
    ap.add_argument("--members", type=int, default=MEMBERS_PER_GPU, help="members per GPU")
    ap.add_argument("--block", type=int, default=0)
    ap.add_argument("--variant", type=int, default=0, help="rrb_opts.variant (kernel A/B timing; 0 = library default)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)
    if args.impl == "reference":
        return run_reference(args)

    import torch
    from rrmpg_b200 import _lib, distributed as rdist, engine
    from rrmpg_b200.models import HBVEdu

    engine.VARIANT = args.variant
    rank, local_rank, world = rdist.init_process_group()
    _lib.require_gpu()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    members = args.members
    n_total = members * world
    lo, hi = rdist.member_block(n_total, rank, world)

    # ---- inputs: rank 0 owns the forcing; parameters are sliced from the seeded global draw
    f, P_all = workload(n_total)
    P = P_all[lo:hi]
    del P_all
    month0 = (f["month"] - 1).astype(np.int8)
    fmat_host, layout = rdist.pack_forcing({"temp": f["temp"], "prec": f["prec"]})
    if rank != 0:
        fmat_host = np.zeros_like(fmat_host)  # filled by the broadcast
    fmat = torch.as_tensor(fmat_host, device=dev)
    d_month = torch.as_tensor(month0, device=dev)
    d_pe = torch.as_tensor(f["PE_m"], device=dev)
    d_tm = torch.as_tensor(f["T_m"], device=dev)
    d_params = torch.as_tensor(engine.pack_params(P), device=dev)
    out = {"qsim": torch.empty((T_STEPS, hi - lo), dtype=torch.float64, device=dev)}
    inits = (0.0, 100.0, 3.0, 10.0)

    def step_device():
        rdist.broadcast_forcing(fmat, src=0)  # the path's single collective (no-op at N=1)
        engine.hbvedu(fmat[1], fmat[0], d_month, d_pe, d_tm, inits, d_params, out=out, math=args.math,
                      block=args.block)  # pack_forcing sorts names: row 0 = prec, row 1 = temp

    for _ in range(args.warmup):
        step_device()
    torch.cuda.synchronize()
    rdist.barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    torch.cuda.synchronize()
    t_wall0 = time.perf_counter()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for a, b in ev:
        a.record()
        step_device()
        b.record()
    stop.record()
    torch.cuda.synchronize()
    rdist.barrier()
    t_wall1 = time.perf_counter()
    total_ms = rdist.max_over_ranks(start.elapsed_time(stop), dev)
    step_ms = [a.elapsed_time(b) for a, b in ev]
    kernel_ms = rdist.max_over_ranks(float(np.mean(step_ms)), dev)
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    value = n_total * T_STEPS * args.steps / (total_ms * 1e-3)

    # parity spot check of the timed output (columns against the oracle) -- not timed
    parity = None
    if rank == 0:
        import oracle
        idx = np.r_[0:8, (hi - lo) - 8:(hi - lo)]
        ref = oracle.hbvedu(f["temp"], f["prec"], month0, f["PE_m"], f["T_m"], inits, P[idx])
        got = out["qsim"][:, torch.as_tensor(idx, device=dev)].cpu().numpy()
        parity = bool(np.allclose(got, ref, rtol=1e-10, atol=1e-12))

    # ---- e2e: the public API with host buffers
    e2e = None
    if not args.no_e2e:
        model = HBVEdu()
        kw = dict(temp=f["temp"], prec=f["prec"], month=f["month"], PE_m=f["PE_m"], T_m=f["T_m"],
                  snow_init=0, soil_init=100, s1_init=3, s2_init=10, params=P)
        engine.DEFAULT_MATH = args.math
        for _ in range(2):
            q = model.simulate(**kw)
            del q
        rdist.barrier()
        t0 = time.perf_counter()
        k_e2e = max(2, min(args.steps, 5))
        for _ in range(k_e2e):
            q = model.simulate(**kw)
            del q
        dt = rdist.max_over_ranks(time.perf_counter() - t0, dev)
        h2d = 2 * T_STEPS * 8 + T_STEPS + 2 * 12 * 8 + (hi - lo) * 11 * 8
        d2h = T_STEPS * (hi - lo) * 8
        e2e = {"value": n_total * T_STEPS * k_e2e / dt, "unit": UNIT, "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "ms_per_step": dt / k_e2e * 1e3, "steps": k_e2e,
               "api": "rrmpg_b200.models.HBVEdu.simulate(numpy) -> numpy [T, N] (pinned), D2H pipelined per time slab"}

    if rank != 0:
        rdist.shutdown()
        return 0

    peak, peak_src = measured_peaks()
    bytes_per_launch = BYTES_PER_MEMBER_STEP * (hi - lo) * T_STEPS
    achieved = bytes_per_launch / (kernel_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "hbv_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as fh:
            traffic = json.load(fh).get("dram_bytes_per_launch")
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "kernel": "rrb::hbv_fast_kernel<qsim-only>" if args.math == "fast" else "rrb::hbv_precise_kernel<qsim-only>",
                "kernel_ms": kernel_ms,
                "note": "algorithmic bytes = 8 B x members x timesteps per launch; duration = CUDA-event time of one "
                        "step (forcing pack kernel + ensemble kernel + the idle fallback launch; pack and fallback "
                        "are <0.3% of it, see profiles/ launch list)"}
    issue = None
    if args.math == "fast":
        slots = 2 * HBV_FAST_FP64_INSTR + (HBV_FAST_WARP_INSTR - HBV_FAST_FP64_INSTR)
        mhz = (clocks or {}).get("sm_mhz") or 1965.0
        ceiling = SM_COUNT * SUBPARTITIONS * mhz * 1e6 * 32 / slots
        warps = (hi - lo + 31) // 32
        per_sp = warps / (SM_COUNT * SUBPARTITIONS)
        issue = {"bound": "issue (fp64 = 2 slots)", "issue_slots_per_member_step": slots,
                 "ceiling_member_steps_per_s": ceiling, "achieved": (hi - lo) * T_STEPS / (kernel_ms * 1e-3),
                 "frac": (hi - lo) * T_STEPS / (kernel_ms * 1e-3) / ceiling,
                 "load_balance_limit": per_sp / float(int(per_sp) + (per_sp > int(per_sp))),
                 "note": "second roofline, explains the HBM fraction: instruction counts from the committed ncu capture; "
                         "load_balance_limit = mean / max warps per SM sub-partition for this ensemble size"}
    cpu = None
    if world == 1 and not args.no_cpu:
        cpu, _ = cpu_baseline_run(f, P)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": config_dict(world, members), "roofline": roofline,
            "issue_roofline": issue, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": LAUNCHES_PER_STEP[args.math] * args.steps, "clocks": clocks,
            "parity_spot_check": parity}
    line["config"]["math"] = args.math
    line["config"]["members_per_gpu"] = members
    print(json.dumps(line), flush=True)
    rdist.shutdown()
    return 0


if __name__ == "__main__":
    sys.exit(main())
