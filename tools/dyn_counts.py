#!/usr/bin/env python3
"""Dynamic instruction estimate per hexad operation from the static SASS: straight-line count + (trip - 1) x loop body for
the single counted loop of hx_mul / hx_sqr / hx_mul_line (cold divergent-warp paths after the RET are ignored), callee
bodies added once per static call site on the hot path.  usage: tools/dyn_counts.py <lib.so|cubin> <kernel>"""
import re, sys
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.abspath(__file__)))
import sass_stats as S

TRIPS = {"hx_mul": 6, "hx_sqr": 4, "hx_mul_line": 3, "hx_mul_fq6": 3}

def main():
    so, kname = sys.argv[1], sys.argv[2]
    funcs = {}
    for name, ins in S.functions(S.disasm(so)):
        base = name.split("$")[1] if name.startswith("$") else name
        if kname not in base or (kname + "_") in base:
            continue
        funcs[S.short(name) if name.startswith("$") else "<body>"] = ins
    res = {}
    def hot(fn):
        if fn in res:
            return res[fn]
        ins = funcs[fn]
        # hot path = up to the first RET (cold paths are placed after it)
        end = next((i for i, (a, t, h) in enumerate(ins) if S.opcode(t).startswith("RET")), len(ins) - 1)
        body = ins[:end + 1]
        n = len(body); w = sum(1 for a, t, h in body if S.opcode(t).startswith("IMAD.WIDE"))
        # one backward branch inside the hot path = the round loop
        addr = {a: i for i, (a, t, h) in enumerate(body)}
        for i, (a, t, h) in enumerate(body):
            m = re.search(r"BRA\S*\s+(?:U?P\d,\s*|!U?P\d,\s*)?`\((\.L_x_\d+)\)", t)
            if m and fn in TRIPS:
                pass
        # loops: find label positions through nvdisasm text is not available here; approximate with the known structure:
        # backward BRA.U UP0 closes the loop; its target is the first instruction after the prologue's last STS / before first LDS of the loop
        loop = None
        for i, (a, t, h) in enumerate(body):
            if S.opcode(t).startswith("BRA.U") and "UP0" in t and "!UP0" not in t:
                loop = i
        if loop is not None and fn in TRIPS:
            # loop start: last UMOV/UIADD3 loop-counter init before the first IMAD.WIDE
            first_w = next(i for i, (a, t, h) in enumerate(body) if S.opcode(t).startswith("IMAD.WIDE"))
            start = max(i for i in range(first_w) if S.opcode(body[i][1]).split(".")[0] in ("UMOV", "CS2R", "UIADD3", "HFMA2", "NOP")) + 1 if first_w else 0
            ln = loop - start + 1
            lw = sum(1 for a, t, h in body[start:loop + 1] if S.opcode(t).startswith("IMAD.WIDE"))
            n += (TRIPS[fn] - 1) * ln; w += (TRIPS[fn] - 1) * lw
        for a, t, h in body:
            m = re.search(r"CALL\S*\s+`\(\$[^$]+\$(\S+)\)", t)
            if m:
                callee = S.short("$x$" + m.group(1))
                if callee in funcs:
                    cn, cw = hot(callee)
                    n += cn; w += cw
        res[fn] = (n, w)
        return res[fn]
    for fn in funcs:
        if fn == "<body>":
            continue
        n, w = hot(fn)
        print("%-22s dyn instr %6d  IMAD.WIDE %5d  other %5d   slots %7.0f" % (fn, n, w, n - w, 4 * w + 1.77 * (n - w)))

main()
