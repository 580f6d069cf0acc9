#!/usr/bin/env python
"""Static instruction mix of a step-kernel instantiation, read from the SASS of the object file.

    python scripts/sass_stats.py [object] [mangled-name substring]

Prints the opcode histogram of the whole kernel and of its hottest loop (the backward branch
that spans the most instructions = the unrolled row march), grouped the way the ncu pipes are
(FP64, shared-memory pipe, integer/move, control).  A CPU-side proxy for the issue / FP64 /
shared-memory floors while no GPU is at hand; timing still decides (profiles/r1_sweep_v2b.log).
"""
import collections
import re
import subprocess
import sys

OBJ = sys.argv[1] if len(sys.argv) > 1 else "py-cubed-sphere_b200/build/fused2b.o"
SUB = sys.argv[2] if len(sys.argv) > 2 else "Li160ELi3ELi1ELi0ELi2ELi14ELi0E"


def kernel_lines(obj, sub):
    txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
    out, on = [], False
    for ln in txt.splitlines():
        if "Function :" in ln:
            on = sub in ln
            continue
        if on:
            m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
            if m:
                out.append((int(m.group(1), 16), m.group(2).strip()))
    return out


def opcode(text):
    t = text.split()
    if t[0].startswith("@"):
        t = t[1:]
    return t[0]


GROUPS = [("fp64", ("DFMA", "DMUL", "DADD", "DSETP", "MUFU.RCP64H", "DMNMX")),
          ("lds", ("LDS",)), ("sts", ("STS",)), ("ldg/stg", ("LDG", "STG", "ST.", "LD.", "RED", "ATOM")),
          ("tma/mbar", ("UBLKCP", "SYNCS", "UTMA", "ELECT")), ("bar", ("BAR", "WARPSYNC", "NANOSLEEP")),
          ("sel/setp", ("FSEL", "SEL", "ISETP", "PLOP3", "LOP3", "FSETP", "P2R", "R2P")),
          ("mov", ("MOV", "IMAD.MOV", "UMOV", "PRMT", "SHFL", "S2R", "CS2R", "R2UR", "S2UR")),
          ("int", ("IMAD", "IADD", "LEA", "SHF", "UIADD", "ULEA", "UIMAD", "USHF", "ULOP", "USEL", "UISETP", "VIADD", "UPLOP")),
          ("branch", ("BRA", "BSSY", "BSYNC", "EXIT", "CALL", "RET", "NOP", "YIELD", "DEPBAR", "MEMBAR", "ERRBAR", "CCTL", "FENCE"))]


def group(op):
    for g, pre in GROUPS:
        if any(op.startswith(p) for p in pre):
            if g == "int" and op.startswith("IMAD.MOV"):
                return "mov"
            return g
    return "other:" + op


def hist(lines):
    h = collections.Counter()
    for _, t in lines:
        h[group(opcode(t))] += 1
    return h


def main():
    k = kernel_lines(OBJ, SUB)
    if not k:
        sys.exit("no kernel matching " + SUB)
    addr = {a: i for i, (a, _) in enumerate(k)}
    loops = []
    for i, (a, t) in enumerate(k):
        m = re.search(r"(0x[0-9a-f]+)\s*$", t)
        if m and opcode(t).startswith("BRA"):
            tgt = int(m.group(1), 16)
            spin = any("SYNCS.PHASECHK" in x for _, x in k[max(0, i - 2):i])   # out-of-line mbarrier retry
            if tgt in addr and addr[tgt] < i and not spin:
                loops.append((i - addr[tgt] + 1, addr[tgt], i))
    # the row march: the tightest loop with at least two rows (four block barriers)
    march = None
    for span, lo, hi in sorted(loops):
        if sum(1 for _, t in k[lo:hi + 1] if opcode(t).startswith("BAR")) >= 4:
            march = (lo, hi)
            break
    print("kernel:", SUB, " instructions:", len(k))
    if march is None:
        sys.exit("no march loop found")
    body = k[march[0]:march[1] + 1]
    rows = max(1, sum(1 for _, t in body if opcode(t).startswith("BAR")) // 2)
    common, elected, on = [], [], False
    for a, t in body:
        op = opcode(t)
        if op.startswith("ELECT"):
            on = True
        (elected if on else common).append((a, t))
        if op.startswith("BSYNC"):
            on = False
    print("march loop 0x%x..0x%x: %d rows per trip" % (body[0][0], body[-1][0], rows))
    for name, seg in (("every warp", common), ("elected lane of warp 0 (TMA issue)", elected)):
        h = hist(seg)
        tot = sum(h.values())
        print("%s: %d instructions = %.1f per row" % (name, tot, tot / rows))
        for g, n in sorted(h.items(), key=lambda x: -x[1]):
            print("   %-16s %6d  %6.1f / row" % (g, n, n / rows))
    ops = collections.Counter(opcode(t) for _, t in common)
    print("opcodes (every warp):", ", ".join("%s %d" % x for x in sorted(ops.items(), key=lambda x: -x[1])[:24]))


if __name__ == "__main__":
    try:
        main()
    except BrokenPipeError:
        pass
