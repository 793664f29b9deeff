#!/usr/bin/env python
"""bench.py -- member-timesteps/sec of the HBV-Edu ensemble hot path on N B200s.

    python bench.py --gpus 1 --steps 20 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
           --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # RRMPG's own numba path on the box's host cores (oracle/_ref)

Workload (BASELINE.json configs[1]): HBVEdu, 65 536 ensemble members per GPU, 40-year daily synthetic forcing
(T = 14 610), parameters uniform in the model's default bounds.  Weak scaling: every rank owns one contiguous member
block of the global ensemble; rank 0 broadcasts the forcing once per step (NCCL, on a side stream, double-buffered
against the previous step's kernel); there is no other communication.

One "step" = one pass of the hot path over the whole ensemble:
  value : inputs resident in HBM; pack forcing + ensemble kernel through the C ABI in device mode (torch tensors),
          timed with CUDA events on the launching stream, max over ranks.
  e2e   : the call a user makes -- HBVEdu.simulate(numpy in, numpy out) for the GLOBAL ensemble, H2D of forcing +
          params and D2H of the full [T, N] discharge inside the timed region.  At N > 1 it is ONE call in ONE process
          (rank 0): the library shards the members over the N devices itself (rrb_opts.n_devices, worker threads).
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
NORTH_STAR_MEMBERS = 1048576  # BASELINE.json north_star: 1M members, strong scaling over the GPUs
BYTES_PER_MEMBER_STEP = 8  # one fp64 qsim store; forcing/params amortise to ~0
CPU_SAMPLE_MEMBERS = 16384   # oracle port (C, pthreads)
REF_MEMBERS_PER_WORKER = 384  # numba reference: members per worker process and step (about 0.3 s of CPU work)
# kernels of ours per device-mode step: forcing pack + FAST ensemble kernel + the PRECISE kernel queued behind it as the
# fallback for flagged CTAs / non-finite rain, which exits at once on this workload
LAUNCHES_PER_STEP = {"fast": 3, "precise": 2}
# Second roofline of the HBV kernel: register-file operand delivery.  Executed warp instructions per
# member-timestep of hbv_fast2_kernel<1 member/thread, qsim only> (one CTA of 14 warps per SM) on this forcing
# (ncu source page, profiles/r02_ncu_full_hbv_v5_summary.txt, tools/ncu_instr_mix.py): 52.69 in total, of which 8.87 DFMA with
# three plain register operands, 0.86 DFMA with a uniform / immediate / reuse-cached operand and 14.03 other fp64-pipe
# instructions (DADD / DMUL / DSETP / I2F).  Measured on the B200 (profiles/r02_fp64_probe_v2.txt): a DFMA with three distinct
# register operands issues every 3 cycles per SM sub-partition, the other fp64 instructions every 2, the rest ~1.
HBV_INSTR = {"total": 52.69, "dfma": 8.87, "fp64_other": 0.86 + 14.03}
SM_COUNT, SUBPARTITIONS = 148, 4


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def load_profile_json(name):
    path = os.path.join(ROOT, "profiles", name)
    if os.path.exists(path):
        with open(path) as fh:
            return json.load(fh)
    return {}


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

    def wait_for_samples(self, n=2, timeout=3.0):
        """nvidia-smi needs a moment to start (longer with eight ranks on the box): do not enter the timed region
        before it delivers."""
        t0 = time.perf_counter()
        while self.proc is not None and len(self.lines) < n and time.perf_counter() - t0 < timeout:
            time.sleep(0.02)

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons, near = [], None, set(), []
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
            near.append((min(abs(ts - t0), abs(ts - t1)), clk))
            if t0 - 0.05 <= ts <= t1 + 0.05:
                sm.append(clk)
                for n, v in zip(names, parts[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
        out = {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
               "samples": len(sm)}
        if not sm and near:  # the timed region fell between two samples: report the closest one and say so
            out["sm_mhz"] = float(min(near)[1])
            out["note"] = "no sample inside the timed region; sm_mhz is the sample closest to it"
        return out


def workload(n_members_total):
    from rrmpg_b200 import synthetic
    from rrmpg_b200.models import HBVEdu
    f = synthetic.forcing(T_STEPS)
    P = synthetic.random_params(HBVEdu(), n_members_total)
    return f, P


# ------------------------------------------------------------------------------------------------------------------
# CPU arms.  kind "reference": the UNMODIFIED RRMPG package installed under oracle/_ref (oracle/build_ref.py), i.e.
# numba's run_hbvedu (rrmpg/models/hbvedu_model.py:16) inside HBVEdu.simulate's member loop (hbvedu.py:199-209) --
# single-threaded as shipped, and on all host cores through one worker process per core (the reference has no
# multi-core path of its own).  kind "port": the C restatement (oracle/rr_oracle.c) on pthreads, used when the
# reference package or numba is missing on the box.
# ------------------------------------------------------------------------------------------------------------------
_REF_STATE = {}


def _ref_worker(job):
    """One worker process: RRMPG's HBVEdu.simulate over `members` parameter sets (as shipped: a Python loop over
    numba calls)."""
    seed, members = job
    st = _REF_STATE
    if "model" not in st:
        import oracle
        pkg = oracle.reference_package()
        from rrmpg.models import HBVEdu as RefHBVEdu  # noqa: resolves to oracle/_ref
        st["model"] = RefHBVEdu()
        assert os.path.abspath(pkg.__file__).startswith(os.path.join(ROOT, "oracle", "_ref"))
    f = st["forcing"]
    np.random.seed(seed)
    P = st["model"].get_random_params(members)
    t0 = time.perf_counter()
    q = st["model"].simulate(temp=f["temp"], prec=f["prec"], month=f["month"], PE_m=f["PE_m"], T_m=f["T_m"],
                             snow_init=0, soil_init=100, s1_init=3, s2_init=10, params=P)
    return members * q.shape[0], time.perf_counter() - t0


class ReferenceArm:
    """Times the reference on the host cores.  step() = every worker simulates REF_MEMBERS_PER_WORKER members."""

    def __init__(self, f):
        import multiprocessing as mp
        import oracle
        self.kind = "reference" if oracle.reference_package() is not None else "port"
        self.f = f
        self.cores = len(os.sched_getaffinity(0))
        if self.kind == "reference":
            _REF_STATE["forcing"] = f
            self.pool = mp.get_context("fork").Pool(self.cores)
            self.members = REF_MEMBERS_PER_WORKER * self.cores
            self.pool.map(_ref_worker, [(k, 8) for k in range(self.cores)], chunksize=1)  # numba JIT per worker, untimed
        else:
            self.cores = oracle.num_threads()
            from rrmpg_b200 import synthetic
            from rrmpg_b200.models import HBVEdu
            self.Ps = oracle.pack_params(synthetic.random_params(HBVEdu(), CPU_SAMPLE_MEMBERS))
            self.month0 = (f["month"] - 1).astype(np.int8)
            self.members = CPU_SAMPLE_MEMBERS

    def step(self, k=0):
        if self.kind == "reference":
            self.pool.map(_ref_worker, [(1000 + 64 * k + w, REF_MEMBERS_PER_WORKER) for w in range(self.cores)], chunksize=1)
        else:
            import oracle
            f = self.f
            oracle.hbvedu(f["temp"], f["prec"], self.month0, f["PE_m"], f["T_m"], (0, 100, 3, 10), self.Ps, nthreads=self.cores)
        return self.members * T_STEPS

    def single_thread(self):
        """As shipped: one process, one thread."""
        if self.kind != "reference":
            return None
        _ref_worker((1, 8))
        n, dt = _ref_worker((2, 256))
        return n / dt

    def describe(self, value, extra=""):
        if self.kind == "reference":
            sample = (f"each step = {self.cores} worker processes x {REF_MEMBERS_PER_WORKER} members x {T_STEPS} steps of the "
                      f"UNMODIFIED rrmpg.models.HBVEdu.simulate (numba run_hbvedu, oracle/_ref; JIT warmed up first); "
                      f"member-timesteps/s does not depend on N{extra}")
        else:
            sample = (f"each step = {CPU_SAMPLE_MEMBERS} members x {T_STEPS} steps of oracle/rr_oracle.c (C restatement of rrmpg "
                      f"run_hbvedu; the reference package / numba is not available on this box), {self.cores} pthreads{extra}")
        return {"value": value, "unit": UNIT, "cores": self.cores, "kind": self.kind, "sample": sample}

    def close(self):
        if self.kind == "reference":
            self.pool.close()
            self.pool.join()


def cpu_baseline_run(f, reps=3):
    """cpu_baseline of our arm (rank 0, N = 1): a bounded sample, about 10-20 s of CPU work in total."""
    arm = ReferenceArm(f)
    single = arm.single_thread()
    best = 0.0
    arm.step(0)
    for k in range(reps):
        t0 = time.perf_counter()
        n = arm.step(k + 1)
        best = max(best, n / (time.perf_counter() - t0))
    out = arm.describe(best, f"; best of {reps} steps")
    if single is not None:
        out["single_thread_value"] = single
        out["single_thread_note"] = "as shipped: the reference has no multi-core path, njit holds the GIL (1 process, 1 thread)"
    arm.close()
    return out


def run_reference(args):
    """--impl reference: same metric / config as our arm, measured on the box's host cores (rank 0 only)."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return 0
    f, _ = workload(8)
    arm = ReferenceArm(f)
    single = arm.single_thread()
    for k in range(args.warmup):
        arm.step(k)
    t0 = time.perf_counter()
    done = 0
    for k in range(args.steps):
        done += arm.step(args.warmup + k)
    dt = time.perf_counter() - t0
    value = done / dt
    cpu = arm.describe(value)
    if single is not None:
        cpu["single_thread_value"] = single
    arm.close()
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(args.gpus), "cpu_baseline": cpu,
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


# ------------------------------------------------------------------------------------------------------------------
def device_timed(args, torch, rdist, engine, dev, f, month0, P, members, steps, warmup, sampler_rank0=False, local_rank=0):
    """K device-resident steps of one rank's member block; returns (total_ms max over ranks, mean step ms max over
    ranks, out tensor, clocks).  The forcing broadcast of step k+1 runs on a side stream while step k computes."""
    rank = int(os.environ.get("RANK", 0))
    fmat_host, _ = rdist.pack_forcing({"temp": f["temp"], "prec": f["prec"]})  # sorted names: row 0 = prec, row 1 = temp
    if rank != 0:
        fmat_host = np.zeros_like(fmat_host)  # filled by the broadcast
    fbuf = [torch.as_tensor(fmat_host, device=dev), torch.as_tensor(fmat_host, device=dev).clone()]
    d_month = torch.as_tensor(month0, device=dev)
    d_pe = torch.as_tensor(f["PE_m"], device=dev)
    d_tm = torch.as_tensor(f["T_m"], device=dev)
    d_params = torch.as_tensor(engine.pack_params(P), device=dev)
    out = {"qsim": torch.empty((T_STEPS, members), dtype=torch.float64, device=dev)}
    inits = (0.0, 100.0, 3.0, 10.0)
    main = torch.cuda.current_stream(dev)
    side = torch.cuda.Stream(dev)
    ready = [torch.cuda.Event(), torch.cuda.Event()]   # forcing of the step is in fbuf[b]
    used = [torch.cuda.Event(), torch.cuda.Event()]    # the step that read fbuf[b] has finished

    def prefetch(b):
        """The path's single collective: rank 0's forcing block to every rank, into buffer b (no-op at N = 1)."""
        with torch.cuda.stream(side):
            side.wait_event(used[b])
            rdist.broadcast_forcing(fbuf[b], src=0)
            ready[b].record(side)

    used[0].record(main)
    used[1].record(main)
    prefetch(0)
    state = {"k": 0}

    def step_device():
        b = state["k"] & 1
        state["k"] += 1
        prefetch(b ^ 1)  # next step's forcing travels while this step computes
        main.wait_event(ready[b])
        engine.hbvedu(fbuf[b][1], fbuf[b][0], d_month, d_pe, d_tm, inits, d_params, out=out, math=args.math, block=args.block)
        used[b].record(main)

    for _ in range(warmup):
        step_device()
    torch.cuda.synchronize()
    rdist.barrier()
    sampler = ClockSampler(local_rank) if sampler_rank0 else None
    if sampler:
        sampler.wait_for_samples()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    torch.cuda.synchronize()
    rdist.barrier()
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
    kernel_ms = rdist.max_over_ranks(float(np.mean([a.elapsed_time(b) for a, b in ev])), dev)
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    return total_ms, kernel_ms, out["qsim"], clocks


def parity_columns(torch, rdist, dev, qsim, f, month0, P):
    """Every rank checks the first / last 8 columns of ITS OWN block against the oracle; all ranks must agree."""
    import oracle
    n = qsim.shape[1]
    idx = np.unique(np.r_[0:min(8, n), max(0, n - 8):n])
    ref = oracle.hbvedu(f["temp"], f["prec"], month0, f["PE_m"], f["T_m"], (0.0, 100.0, 3.0, 10.0), P[idx])
    got = qsim[:, torch.as_tensor(idx, device=dev)].cpu().numpy()
    ok = bool(np.allclose(got, ref, rtol=1e-10, atol=1e-12) and np.array_equal(np.isnan(got), np.isnan(ref)))
    return rdist.max_over_ranks(0.0 if ok else 1.0, dev) == 0.0


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
    ap.add_argument("--no-north-star", action="store_true")
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
    month0 = (f["month"] - 1).astype(np.int8)

    total_ms, kernel_ms, qsim, clocks = device_timed(args, torch, rdist, engine, dev, f, month0, P, hi - lo, args.steps,
                                                     args.warmup, sampler_rank0=(rank == 0), local_rank=local_rank)
    value = n_total * T_STEPS * args.steps / (total_ms * 1e-3)
    parity = parity_columns(torch, rdist, dev, qsim, f, month0, P)  # not timed
    del qsim
    torch.cuda.empty_cache()

    # ---- north star (BASELINE.json): 1M members in total, strong scaling over the GPUs, device timed
    north = None
    if not args.no_north_star and members == MEMBERS_PER_GPU:
        per = NORTH_STAR_MEMBERS // world
        need = per * T_STEPS * 8 + (2 << 30)
        free = torch.cuda.mem_get_info(dev)[0]
        fits = rdist.max_over_ranks(0.0 if free >= need else 1.0, dev) == 0.0
        if fits:
            from rrmpg_b200 import synthetic
            Pn_all = synthetic.random_params(HBVEdu(), NORTH_STAR_MEMBERS)
            nlo, nhi = rdist.member_block(NORTH_STAR_MEMBERS, rank, world)
            ns_steps = 3
            ns_total, _, q_ns, _ = device_timed(args, torch, rdist, engine, dev, f, month0, Pn_all[nlo:nhi], nhi - nlo, ns_steps, 2)
            ns_parity = parity_columns(torch, rdist, dev, q_ns, f, month0, Pn_all[nlo:nhi])
            del q_ns, Pn_all
            torch.cuda.empty_cache()
            peak, _ = measured_peaks()
            ns_value = NORTH_STAR_MEMBERS * T_STEPS * ns_steps / (ns_total * 1e-3)
            north = {"workload": f"HBVEdu {NORTH_STAR_MEMBERS} members in total ({per} per GPU, strong scaling) x {T_STEPS} steps, "
                                 "qsim [T, N] device resident", "value": ns_value, "unit": UNIT, "n_gpus": world,
                     "ms_per_step": ns_total / ns_steps, "steps": ns_steps,
                     "hbm_roofline_frac_per_gpu": 8.0 * ns_value / world / 1e9 / peak, "parity_spot_check": ns_parity}
        else:
            north = {"skipped": f"needs {need / 1e9:.0f} GB free per GPU, {free / 1e9:.0f} GB available"}

    # ---- e2e: the public API with host buffers, the global ensemble in ONE call (rank 0; sharded inside the library)
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(args, rdist, engine, HBVEdu, f, P_all, n_total, world, rank)
    del P_all

    if rank != 0:
        rdist.shutdown()
        return 0

    peak, peak_src = measured_peaks()
    bytes_per_launch = BYTES_PER_MEMBER_STEP * (hi - lo) * T_STEPS
    achieved = bytes_per_launch / (kernel_ms * 1e-3) / 1e9
    traffic = load_profile_json("hbv_traffic.json").get("dram_bytes_per_launch")
    if args.math != "fast":
        kernel_name = "rrb::hbv_precise_kernel<qsim only>"
    elif (hi - lo + 31) // 32 <= 16 * SM_COUNT:   # launch_hbvedu: up to 16 member-warps per SM
        kernel_name = "rrb::hbv_fast2_kernel<1 member/thread, qsim only>, one CTA per SM"
    else:
        kernel_name = "rrb::hbv_fast2_kernel<2 members/thread, qsim only>"
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "kernel": kernel_name, "kernel_ms": kernel_ms,
                "traffic_source": "ncu --set full dram__bytes_read.sum + dram__bytes_write.sum of one launch of this kernel at this "
                                  "size (profiles/hbv_traffic.json; not re-measured in this run)",
                "note": "algorithmic bytes = 8 B x members x timesteps per launch; duration = CUDA-event time of one "
                        "step (forcing pack kernel + ensemble kernel + the idle fallback launch; pack and fallback "
                        "are <0.3% of it, see profiles/ launch list)"}
    issue = None
    if args.math == "fast" and members == MEMBERS_PER_GPU:
        # what actually bounds the kernel: the rate at which a sub-partition can deliver register
        # operands -- cycles = 3 x DFMA + 2 x other fp64 + 1 x every other warp instruction (micro-probe
        # profiles/r02_fp64_probe_v2.txt; consistent with the wet / dry ablation runs, profiles/r02_ablations.txt)
        other = HBV_INSTR["total"] - HBV_INSTR["dfma"] - HBV_INSTR["fp64_other"]
        slots = 3 * HBV_INSTR["dfma"] + 2 * HBV_INSTR["fp64_other"] + other
        mhz = (clocks or {}).get("sm_mhz") or 1965.0
        ceiling = SM_COUNT * SUBPARTITIONS * mhz * 1e6 * 32 / slots
        warps = (hi - lo + 31) // 32
        per_sp = warps / (SM_COUNT * SUBPARTITIONS)
        rate = (hi - lo) * T_STEPS / (kernel_ms * 1e-3)
        issue = {"bound": "register-operand issue: 3 cycles per DFMA with three register operands, 2 per other fp64 instruction, "
                          "1 per other warp instruction",
                 "issue_slots_per_member_step": slots, "ceiling_member_steps_per_s": ceiling, "achieved": rate,
                 "frac": rate / ceiling, "sm_mhz_used": mhz,
                 "load_balance_limit": per_sp / float(int(per_sp) + (per_sp > int(per_sp))),
                 "note": "second roofline, explains the HBM fraction: instruction counts from the committed ncu capture; "
                         "load_balance_limit = mean / max member-warps per SM sub-partition for this ensemble size "
                         "(65 536 members = 3.46 member-warps per sub-partition, 4 on the busiest: one CTA of 14 warps per SM, "
                         "warp w on sub-partition w % 4)"}
    cpu = None
    if world == 1 and not args.no_cpu:
        cpu = cpu_baseline_run(f)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": config_dict(world, members), "roofline": roofline,
            "issue_roofline": issue, "cpu_baseline": cpu, "e2e": e2e,
            "gpu_launches": LAUNCHES_PER_STEP[args.math] * args.steps, "clocks": clocks,
            "parity_spot_check": parity, "parity_spot_check_ranks": world, "extra": {"north_star": north}}
    line["config"]["math"] = args.math
    line["config"]["members_per_gpu"] = members
    print(json.dumps(line), flush=True)
    rdist.shutdown()
    return 0


def live_d2h_probe(world):
    """Ceiling of THIS box for the end-to-end path: aggregate GB/s of concurrent pinned D2H copies on `world` devices
    (csrc/tools/d2h_probe, built by __graft_entry__.build()).  None when the tool is missing."""
    exe = os.path.join(ROOT, "rrmpg_b200", "csrc", "tools", "d2h_probe")
    if not os.path.exists(exe):
        return None
    try:
        out = subprocess.run([exe, str(world), "128", "8", "d2h-only"], capture_output=True, text=True, timeout=120).stdout
    except (OSError, subprocess.TimeoutExpired):
        return None
    best = None
    for line in out.splitlines():
        if " D2H: " in line and f": {world} device(s) x" in line:
            try:
                gbs = float(line.split("stream(s):")[1].split("GB/s")[0])
            except (IndexError, ValueError):
                continue
            best = gbs if best is None else max(best, gbs)
    return best


def run_e2e(args, rdist, engine, HBVEdu, f, P_all, n_total, world, rank):
    """HBVEdu.simulate(numpy) -> numpy [T, n_total]: H2D of forcing + params, D2H of the whole discharge array."""
    engine.DEFAULT_MATH = args.math
    probe = load_profile_json("d2h_probe.json")
    out = None
    rdist.host_group()      # (collective) the ranks that wait below must not occupy their GPUs: gloo, not NCCL
    rdist.host_barrier()
    if rank == 0:
        model = HBVEdu()
        kw = dict(temp=f["temp"], prec=f["prec"], month=f["month"], PE_m=f["PE_m"], T_m=f["T_m"],
                  snow_init=0, soil_init=100, s1_init=3, s2_init=10, params=P_all)
        engine.DEVICES = list(range(world)) if world > 1 else "one"
        for _ in range(2):  # warm-up: pinned output pool, per-device contexts and scratch
            q = model.simulate(**kw)
            del q
        k_e2e = max(2, min(args.steps, 5))
        t0 = time.perf_counter()
        for _ in range(k_e2e):
            q = model.simulate(**kw)
            del q
        dt = time.perf_counter() - t0
        h2d = world * (2 * T_STEPS * 8 + T_STEPS + 2 * 12 * 8) + n_total * 11 * 8
        d2h = T_STEPS * n_total * 8
        d2h_gbs = d2h * k_e2e / dt / 1e9
        ceiling = live_d2h_probe(world)   # after the timed region: the other ranks still wait on the host barrier
        probe_src = "measured live on this box after the timed region (csrc/tools/d2h_probe: concurrent pinned D2H copies)"
        if ceiling is None:
            ceiling = probe.get("d2h_gbs_by_devices", {}).get(str(world))
            probe_src = "profiles/d2h_probe.json (the pool's 8-GPU box; the live probe tool is not built here)"
        out = {"value": n_total * T_STEPS * k_e2e / dt, "unit": UNIT, "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "ms_per_step": dt / k_e2e * 1e3, "steps": k_e2e, "d2h_gbs": d2h_gbs,
               "d2h_probe_gbs": ceiling, "frac_of_probe": (d2h_gbs / ceiling) if ceiling else None,
               "probe_source": probe_src,
               "api": "rrmpg_b200.models.HBVEdu.simulate(numpy) -> numpy [T, N] (pinned), ONE call for the global ensemble; "
                      + (f"the library shards the members over the {world} GPUs (rrb_opts.n_devices, one worker thread per device), "
                         if world > 1 else "") + "D2H pipelined per time slab"}
        # the same call as rrmpg_b200.tools.monte_carlo(..., qobs, return_qsim=False) makes it: the per-member objective is
        # accumulated in the kernel's registers and mse[N] is all that comes back (rrmpg/tools/monte_carlo.py:64-73 loops
        # simulate + calc_mse over the members).  Host buffers in, host result out, wall clock.
        qobs = np.abs(np.random.default_rng(11).normal(2.0, 1.0, T_STEPS))
        def fused_call():
            with engine.fused(qobs, objective="mse", want_qsim=False) as fz:
                model.simulate(**kw)
            return np.asarray(fz.values)
        try:
            for _ in range(2):
                m = fused_call()
            k_f = 10
            t0 = time.perf_counter()
            for _ in range(k_f):
                m = fused_call()
            dtf = time.perf_counter() - t0
            out["fused_objective"] = {"value": n_total * T_STEPS * k_f / dtf, "unit": UNIT, "ms_per_call": dtf / k_f * 1e3,
                                      "h2d_bytes_per_call": h2d + T_STEPS * 8, "d2h_bytes_per_call": n_total * 8,
                                      "finite": bool(np.isfinite(m).all()),
                                      "api": "with engine.fused(qobs, want_qsim=False): HBVEdu.simulate(numpy) -- what "
                                             "tools.monte_carlo(model, num, qobs, return_qsim=False) runs; result = mse[N] on the host"}
        except Exception as exc:   # a side figure must never cost the bench line
            out["fused_objective"] = {"error": repr(exc)}
    rdist.host_barrier()
    return out


if __name__ == "__main__":
    sys.exit(main())
