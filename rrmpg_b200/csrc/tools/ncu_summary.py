"""Summarise an .ncu-rep: key metrics, stall breakdown, hottest SASS lines.  usage: ncu_summary.py rep [nlines]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; nl = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "launch__registers_per_thread", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_write.sum", "dram__bytes_read.sum", "launch__grid_size", "launch__block_size", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"]
d = dict(zip(hdr, zip(units, vals)))
print("kernel:", d.get("Kernel Name", ("", ""))[1][:110])
for w in want:
    if w in d: print(f"  {w:70s} {d[w][1]:>14s} {d[w][0]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}; data = rows[2:]
tot = sum(int(r[ix["# Samples"]]) for r in data)
agg = {h: sum(int(r[ix[h]] or 0) for r in data) for h in hdr if h.startswith("stall_") and "Not Issued" not in h}
print("  stalls:", ", ".join(f"{k[6:]} {100*v/tot:.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
texec = sum(int(r[ix["Instructions Executed"]]) for r in data)
print(f"  warp instructions executed: {texec:,}   static instructions: {len(data)}")
top = sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:nl]
for r in top:
    print(f"   {100*int(r[ix['# Samples']])/tot:5.2f}%  exec={int(r[ix['Instructions Executed']])/1e6:7.1f}M  {r[ix['Source']].strip()[:90]}")
