"""Static SASS audit of the ensemble kernels (runs without a GPU: `cuobjdump -sass` on the built library).

For every kernel whose demangled name matches one of the patterns the script finds the steady-state timestep loop
(the backward branch that spans the most instructions and contains the output store) and reports its instruction
mix: fp64 pipe (DFMA/DADD/DMUL/DSETP/...), the other pipes, shared-memory forcing reads, global stores, and the
TMA / mbarrier mnemonics (UBLKCP, SYNCS) that prove the forcing ring is a bulk-copy pipeline.  The counts are static
(per loop trip, fallback paths included).

    python rrmpg_b200/csrc/tools/sass_audit.py rrmpg_b200/librrmpg_b200.so > profiles/rNN_sass_audit.md
"""
import collections
import re
import subprocess
import sys

KERNELS = [
    # (label, regex on the demangled name)
    ("HBVEdu FAST, one member per thread, qsim only, 256-step tiles (bench kernel at 65 536 members)",
     r"hbv_fast2_kernel<1, true, false, 0, 0, 256>"),
    ("HBVEdu FAST, two members per thread, qsim only, 256-step tiles (1M members)", r"hbv_fast2_kernel<2, true, false, 0, 0, 256>"),
    ("HBVEdu FAST, rotating schedule, 8 warps, qsim only (opt-in)", r"hbv_rot_kernel<8, true, false, 0>"),
    ("HBVEdu PRECISE, qsim only", r"hbv_precise_kernel<true, false, false>"),
    ("ABC pair kernel (two members per thread)", r"abc_pair_kernel"),
    ("GR4J FAST x4 <= 2.5 class, qsim only", r"gr4j_kernel<rrb::Gr4jMember<3, 7, 0, false>, true, true>"),
    ("Cemaneige contract step, 5 layers, outflow only", r"cema_kernel<5, rrb::NoGr4j, false, true, true, 0>"),
    ("CemaneigeGR4J FAST, 5 layers, qsim only", r"cema_kernel<5, rrb::Gr4jMember<3, 7, 0, false>, true, true, true, 0>"),
]

FP64 = ("DFMA", "DADD", "DMUL", "DSETP", "DMNMX", "MUFU.RCP64H", "MUFU.RSQ64H", "F2F.F64", "I2F.F64", "F2I.F64", "D2")
TMA = ("UBLKCP", "SYNCS", "UTMA")


def functions(so):
    txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
    names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", txt)), capture_output=True,
                           text=True).stdout.splitlines()
    cur, k = None, -1
    out = {}
    for line in txt.splitlines():
        if "Function :" in line:
            k += 1
            cur = out.setdefault(names[k], [])
            continue
        if cur is None:
            continue
        m = re.search(r"/\*([0-9a-f]{4,6})\*/\s+(.*?);\s*/\* 0x[0-9a-f]{16} \*/", line)
        if m:
            cur.append((int(m.group(1), 16), m.group(2).strip()))
    return out


def opcode(text):
    t = re.sub(r"^@!?U?P\d+\s+", "", text)
    return t.split()[0]


def store_loops(rows):
    """Innermost store-carrying loops: backward branches (target address < branch address) whose body holds a
    global store of the output (STG) and no other such loop.  These are the (unrolled) timestep loops of the kernel
    (steady state, tile tails, fallback variants)."""
    addr_index = {a: i for i, (a, _) in enumerate(rows)}
    loops = []
    for i, (a, text) in enumerate(rows):
        if not re.match(r"(@!?U?P\d+\s+)?BRA", text):
            continue
        m = re.search(r"0x([0-9a-f]+)", text)
        if not m:
            continue
        tgt = int(m.group(1), 16)
        if tgt >= a or tgt not in addr_index:
            continue
        lo = addr_index[tgt]
        if any(opcode(t).startswith("STG") for _, t in rows[lo:i + 1]):
            loops.append((lo, i))
    inner = [(lo, hi) for lo, hi in loops if not any((l2, h2) != (lo, hi) and lo <= l2 and h2 <= hi for l2, h2 in loops)]
    return [rows[lo:hi + 1] for lo, hi in inner]


def classify(body):
    mix = collections.Counter()
    for _, t in body:
        op = opcode(t)
        if op.startswith(TMA):
            mix["tma/mbarrier (UBLKCP, SYNCS)"] += 1
        elif op.startswith(FP64):
            mix["fp64 pipe"] += 1
        elif op.startswith("STG"):
            mix["global store " + ".".join(op.split(".")[:1] + [p for p in op.split(".") if p in ("64", "128")])] += 1
        elif op.startswith("LDG"):
            mix["global load"] += 1
        elif op.startswith(("LDS", "STS")):
            mix["shared memory " + op.split(".")[0]] += 1
        elif op.startswith(("BRA", "BSSY", "BSYNC", "WARPSYNC", "BAR", "VOTE", "EXIT", "CALL", "RET")):
            mix["control / vote / barrier"] += 1
        elif op.startswith(("HMMA", "UTC", "TCGEN", "QGMMA", "IMMA", "DMMA")):
            mix["tensor core"] += 1
        else:
            mix["integer / fp32 / move / select"] += 1
    return mix


def main():
    so = sys.argv[1]
    fns = functions(so)
    print("# Static SASS audit of the ensemble kernels (sm_100a)\n")
    print("Produced by `rrmpg_b200/csrc/tools/sass_audit.py` from `cuobjdump -sass` of the built library; no GPU needed.")
    print("Timestep loops = the innermost backward branches whose body contains the output store.  The counts are STATIC:")
    print("the rare-operand fallbacks (CTA-uniform votes) and the tile refill sit on forward branches inside the loops, so")
    print("they bound from above what a warp issues per trip; the EXECUTED counts are in the ncu summaries next to this")
    print("file (HBVEdu FAST: 61.6 warp instructions per member-step).  What the audit proves without a GPU: the forcing")
    print("ring is a TMA bulk-copy pipeline (UBLKCP + SYNCS, no LDG in the timestep loops), there is no tensor-core")
    print("instruction anywhere, and the width of the output stores.\n")
    for label, pat in KERNELS:
        hits = [n for n in fns if re.search(pat, n)]
        if not hits:
            print(f"## {label}\n\nno kernel matches `{pat}`\n")
            continue
        name = hits[0]
        rows = fns[name]
        whole = collections.Counter(opcode(t).split(".")[0] for _, t in rows)
        print(f"## {label}\n")
        print(f"`{name.split('(')[0]}`: {len(rows)} SASS instructions; {whole.get('UBLKCP', 0)} UBLKCP (TMA bulk copy), "
              f"{whole.get('SYNCS', 0)} SYNCS (mbarrier), {whole.get('LDG', 0)} LDG, "
              f"{sum(v for k, v in whole.items() if k.startswith(('HMMA', 'UTC', 'IMMA', 'DMMA', 'QGMMA')))} tensor-core "
              f"instructions; output stores: {', '.join(sorted(set(o for o in (opcode(t) for _, t in rows) if o.startswith('STG'))))}.\n")
        print("| timestep loop | instructions | output stores (= steps x members per thread) | fp64 pipe | LDS | LDG | control | other (int / fp32 / mov / sel) | static instructions per store |")
        print("|---|---|---|---|---|---|---|---|---|")
        for body in store_loops(rows):
            mix = classify(body)
            nst = sum(v for k, v in mix.items() if k.startswith("global store"))
            other = mix.get("integer / fp32 / move / select", 0)
            print(f"| 0x{body[0][0]:05x}..0x{body[-1][0]:05x} | {len(body)} | {nst} | {mix.get('fp64 pipe', 0)} | "
                  f"{mix.get('shared memory LDS', 0)} | {mix.get('global load', 0)} | {mix.get('control / vote / barrier', 0)} | {other} | {len(body) / nst:.1f} |")
        print()


if __name__ == "__main__":
    main()
