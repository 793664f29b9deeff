"""Decode the static stall counts (control bits 105..108 of every 128-bit SASS instruction) from
`cuobjdump -sass` output: sum of stalls = cycles one warp needs to issue a straight-line region alone."""
import re, sys, subprocess
so, pat = sys.argv[1], sys.argv[2]
lo = int(sys.argv[3], 16) if len(sys.argv) > 3 else 0
hi = int(sys.argv[4], 16) if len(sys.argv) > 4 else 1 << 30
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
on = False; rows = []; cur = None
for line in txt.splitlines():
    if "Function :" in line:
        on = pat in line
        continue
    if not on: continue
    m = re.search(r"/\*([0-9a-f]{4})\*/\s+(.*?);\s*/\* (0x[0-9a-f]{16}) \*/", line)
    if m:
        cur = [int(m.group(1), 16), m.group(2).strip(), int(m.group(3), 16), None]; rows.append(cur); continue
    m = re.search(r"/\* (0x[0-9a-f]{16}) \*/", line)
    if m and cur is not None and cur[3] is None:
        cur[3] = int(m.group(1), 16)
tot = 0; n = 0
for addr, text, w0, w1 in rows:
    if w1 is None or addr < lo or addr >= hi: continue
    stall = (w1 >> 41) & 0xF; yld = (w1 >> 45) & 1; wbar = (w1 >> 46) & 7; rbar = (w1 >> 49) & 7; wait = (w1 >> 52) & 0x3F
    tot += stall; n += 1
    if len(sys.argv) > 5: print(f"{addr:04x} st={stall:2d} y={yld} wb={wbar} rb={rbar} wm={wait:02x} {text}")
print(f"{n} instructions, sum of stall counts = {tot} cycles")
