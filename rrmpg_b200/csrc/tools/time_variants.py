"""Times HBV kernel build variants (RRMPG_B200_LIB=<.so>) on the bench workload.  Development aid."""
import os, subprocess, sys, json
root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
for lib in sys.argv[1:]:
    env = dict(os.environ, RRMPG_B200_LIB=os.path.abspath(lib))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--steps", "10", "--warmup", "3", "--no-e2e",
                          "--no-cpu"], env=env, capture_output=True, text=True).stdout
    try:
        d = json.loads(out.strip().splitlines()[-1])
        print(f"{os.path.basename(lib):24s} {d['ms_per_step']:.3f} ms  {d['value']/1e9:.1f} G/s  frac {d['roofline']['frac']:.3f} parity={d['parity_spot_check']}")
    except Exception as e:
        print(lib, "failed", out[-300:])
