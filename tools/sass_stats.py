#!/usr/bin/env python3
"""Static SASS statistics per device function of one kernel: instruction mix and the sum of the scheduler's static stall
counts (control bits 105-108 of each 128-bit instruction), i.e. the minimum issue time of ONE warp through that code.

usage: tools/sass_stats.py <lib.so> <kernel-name-substring> [function-substring-to-dump]
Used to compare builds without spending GPU time: sum(stall) ~ 4 x instructions means the code is one dependent
chain; the fma-heavy pipe bound is 4 cycles per IMAD.WIDE per warp (2 warps per scheduler share it).
"""
import os
import re
import subprocess
import sys
import tempfile
from collections import Counter


def disasm(so):
    if so.endswith(".cubin"):
        return subprocess.run(["nvdisasm", "-c", "-hex", so], capture_output=True, text=True).stdout.splitlines()
    d = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=d, capture_output=True)
    cubin = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
    return subprocess.run(["nvdisasm", "-c", "-hex", os.path.join(d, cubin)], capture_output=True, text=True).stdout.splitlines()


def functions(txt):
    """yield (name, [(addr, text, hi)])"""
    rx = re.compile(r"^\s+/\*([0-9a-f]{4,})\*/\s+(.*?);\s+/\* 0x([0-9a-f]{16}) \*/")
    rx2 = re.compile(r"^\s+/\* 0x([0-9a-f]{16}) \*/")
    name, cur = None, []
    i = 0
    while i < len(txt):
        l = txt[i]
        m = re.match(r"^\s+\.type\s+(\S+),@function", l)
        if m:
            if name is not None:
                yield name, cur
            name, cur = m.group(1), []
            i += 1
            continue
        m = rx.match(l)
        if m and name is not None:
            hi = int(rx2.match(txt[i + 1]).group(1), 16)
            cur.append((int(m.group(1), 16), m.group(2).strip(), hi))
            i += 2
            continue
        i += 1
    if name is not None:
        yield name, cur


def short(n):
    m = re.search(r"\$_ZN2bn\d+([a-z0-9_]+?)(?:I|E)", n)
    if m:
        return m.group(1)
    m = re.search(r"\$_Z\d+([a-z0-9_]+?)(?:RK|I|E|P)", n)
    return m.group(1) if m else n


def opcode(t):
    p = t.split()
    return p[1] if p[0].startswith("@") else p[0]


def main():
    so, kname = sys.argv[1], sys.argv[2]
    dump = sys.argv[3] if len(sys.argv) > 3 else None
    txt = disasm(so)
    print("%-22s %6s %6s %6s %6s %6s %6s %6s %7s %6s" % ("function", "instr", "IMAD.W", "IADD3", "SEL", "mem", "IMAD*", "other", "stall", "st/ins"))
    for name, ins in functions(txt):
        base = name.split("$")[1] if name.startswith("$") else name
        if kname not in base or (kname + "_") in base:
            continue
        c = Counter()
        st = 0
        for a, t, h in ins:
            op = opcode(t)
            s = (h >> 41) & 0xF
            st += s
            if op.startswith("IMAD.WIDE"):
                c["imadw"] += 1
            elif op.startswith("IMAD"):
                c["imad"] += 1
            elif op.startswith("IADD") or op.startswith("VIADD"):
                c["iadd"] += 1
            elif op.startswith("SEL"):
                c["sel"] += 1
            elif op.split(".")[0] in ("LD", "ST", "LDS", "STS", "LDG", "STG", "LDL", "STL", "LDC"):
                c["mem"] += 1
            else:
                c["other"] += 1
                c["o:" + op.split(".")[0]] += 1
        sn = short(name) if name.startswith("$") else "<kernel body>"
        others = ", ".join("%s %d" % (k[2:], v) for k, v in c.most_common() if k.startswith("o:"))[:70]
        n = max(len(ins), 1)
        print("%-22s %6d %6d %6d %6d %6d %6d %6d %7d %6.2f  %s" % (sn[:22], len(ins), c["imadw"], c["iadd"], c["sel"], c["mem"], c["imad"],
                                                             c["other"], st, st / n, others))
        if dump and dump in sn:
            for a, t, h in ins:
                print("    %05x s%-2d y%d w%d r%d m%02x  %s" % (a, (h >> 41) & 0xF, (h >> 45) & 1, (h >> 46) & 7, (h >> 49) & 7, (h >> 52) & 0x3F, t))


if __name__ == "__main__":
    main()
