"""Executed-instruction mix of one profiled kernel per member-timestep, for bench.py's register-operand roofline.
usage: ncu_instr_mix.py rep members timesteps
Counts warp instructions from the source page of an .ncu-rep (ncu --set full --import-source on): DFMA whose three
source operands are all (non-reused) registers, DFMA with a uniform / immediate / constant / reuse-cached operand, other
fp64-pipe instructions (DADD, DMUL, DSETP, I2F/F2F.F64), everything else."""
import csv, io, re, subprocess, sys
rep, members, steps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
mix = {"dfma_3reg": 0, "dfma_other": 0, "fp64_other": 0, "rest": 0}
for r in rows[2:]:
    n = int(r[ix["Instructions Executed"]] or 0)
    text = re.sub(r"^@!?U?P\d+\s+", "", r[ix["Source"]].strip())
    op = text.split()[0] if text else ""
    base = op.split(".")[0]
    if base == "DFMA":
        srcs = text.split(",")[1:]
        plain = sum(1 for s in srcs if re.search(r"(?<![U\w])R\d+", s) and ".reuse" not in s)
        mix["dfma_3reg" if plain == 3 else "dfma_other"] += n
    elif base in ("DADD", "DMUL", "DSETP") or (base in ("I2F", "F2F") and "F64" in op):
        mix["fp64_other"] += n
    else:
        mix["rest"] += n
per = members / 32 * steps   # warp-steps of 32 members
tot = sum(mix.values())
print(f"warp instructions per member-timestep (x32 members per warp): total {tot / per:.2f}")
for k, v in mix.items():
    print(f"  {k:12s} {v / per:7.2f}")
slots = (3 * mix["dfma_3reg"] + 2 * (mix["dfma_other"] + mix["fp64_other"]) + mix["rest"]) / per
print(f"issue slots per member-timestep (3 / 2 / 2 / 1 cycles): {slots:.2f}")
