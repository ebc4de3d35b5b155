#!/usr/bin/env python3
"""Weighted static SASS cost per out-of-line device function of one kernel.

For the pairing kernels a launch's time is (4.1 x IMAD.WIDE + 2.19 x FP64 + 2.0 x ALU-pipe instructions) per scheduler,
times ceil(warps per scheduler) / (warps per scheduler) -- `bench.py` checks that model against the ncu instruction mix
of every run (`roofline.pipe_model`).  This script applies the same weights to the SASS of a build, split at the targets of
the kernel's CALL instructions (the out-of-line Fq2 / hexad operations), and lists each piece with the pieces it calls, so
that two builds can be compared before any GPU time is spent:

    cost(line doubling step) = body + 4 x duo_mul + 6 x duo_sqr + 2 x fp_mul_ni + 3 x duo_mul_xi

usage: tools/sass_cost.py <lib.so> <mangled-kernel-name-or-substring>
Straight-line pieces only: loops (the hexad product rounds) are counted once -- use ncu's instruction mix for those.
"""
import bisect
import collections
import re
import subprocess
import sys

# issue interval (scheduler cycles per warp instruction) of the three pipes that add up; everything else (fma-pipe moves and
# carries, shared-memory accesses, control) fits in the gaps
WEIGHT = {"IMAD.WIDE": 4.1, "DFMA": 2.19, "DADD": 2.19, "DMUL": 2.19}
ALU = {"IADD3", "LOP3", "SHF", "SEL", "ISETP", "LEA", "IADD", "VIADD", "PRMT", "SGXT", "PLOP3", "POPC", "FLO", "BREV", "IABS", "IMNMX", "VIMNMX"}


def opcode(text):
    t = text.split()
    op = t[1] if t[0].startswith("@") else t[0]
    if op.startswith("IMAD"):
        return "IMAD.WIDE" if "WIDE" in op else "IMAD"
    return op.split(".")[0]


def main():
    so, kernel = sys.argv[1], sys.argv[2]
    sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout.splitlines()
    ins, on = [], False
    for line in sass:
        if "Function :" in line:
            on = kernel in line
            continue
        if on:
            m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);", line)
            if m:
                ins.append((int(m.group(1), 16), m.group(2)))
    if not ins:
        sys.exit("no kernel matching %r in %s" % (kernel, so))
    call = re.compile(r"CALL\.\w+\.?\w*\s+.*?(0x[0-9a-f]+)")
    starts = sorted({0} | {int(m.group(1), 16) for _, t in ins for m in [call.search(t)] if m})
    pieces = collections.OrderedDict((s, {"n": 0, "ops": collections.Counter(), "calls": collections.Counter()}) for s in starts)
    for addr, text in ins:
        p = pieces[starts[bisect.bisect_right(starts, addr) - 1]]
        p["n"] += 1
        p["ops"][opcode(text)] += 1
        m = call.search(text)
        if m:
            p["calls"][int(m.group(1), 16)] += 1
    print("%-8s %6s %8s  %s" % ("piece", "instr", "cycles", "mix / calls"))
    for s, p in pieces.items():
        cyc = sum(WEIGHT.get(o, 2.0 if o in ALU else 0.0) * c for o, c in p["ops"].items())
        p["cyc"] = cyc
        mix = " ".join("%s %d" % oc for oc in p["ops"].most_common(6))
        calls = " ".join("%dx0x%x" % (c, t) for t, c in p["calls"].items())
        print("0x%-6x %6d %8.0f  %s%s" % (s, p["n"], cyc, mix, ("  | calls " + calls) if calls else ""))
    print("\nwith callees (one level of straight-line callees):")
    for s, p in pieces.items():
        if p["calls"]:
            tot_n = p["n"] + sum(c * pieces[t]["n"] for t, c in p["calls"].items())
            tot_c = p["cyc"] + sum(c * pieces[t]["cyc"] for t, c in p["calls"].items())
            print("0x%-6x %6d instructions, %8.0f weighted cycles per call" % (s, tot_n, tot_c))


if __name__ == "__main__":
    main()
