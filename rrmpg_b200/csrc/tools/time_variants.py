"""Times HBV kernel build variants (RRMPG_B200_LIB=<.so>) on the bench workload.  Development aid.
usage: time_variants.py [--members N] [--env K=V ...] lib1.so lib2.so ..."""
import os, subprocess, sys, json
root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
args = sys.argv[1:]
members, extra_env = None, {}
while args and args[0].startswith("--"):
    if args[0] == "--members":
        members = args[1]; args = args[2:]
    elif args[0] == "--env":
        k, v = args[1].split("=", 1); extra_env[k] = v; args = args[2:]
    else:
        raise SystemExit(f"unknown option {args[0]}")
for lib in args:
    env = dict(os.environ, RRMPG_B200_LIB=os.path.abspath(lib), **extra_env)
    cmd = [sys.executable, os.path.join(root, "bench.py"), "--steps", "10", "--warmup", "3", "--no-e2e", "--no-cpu"]
    if members:
        cmd += ["--members", members]
    out = subprocess.run(cmd, env=env, capture_output=True, text=True).stdout
    try:
        d = json.loads(out.strip().splitlines()[-1])
        print(f"{os.path.basename(lib):24s} {extra_env} members={members or 'default'} {d['ms_per_step']:.3f} ms  {d['value']/1e9:.1f} G/s  "
              f"frac {d['roofline']['frac']:.3f} parity={d['parity_spot_check']}", flush=True)
    except Exception as e:
        print(lib, "failed", out[-300:], flush=True)
