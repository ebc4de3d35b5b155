#!/usr/bin/env python3
"""Summarise one `ncu --set full --import-source on` capture per pairing kernel into profiles/ncu_kernels.json:
DRAM traffic, duration, pipe / stall metrics and the executed-instruction mix (per-opcode counts from the source page).
usage: tools/ncu_kernels.py <tag> <out.json> name=file.ncu-rep ..."""
import collections, csv, io, json, os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

RAW = ["sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
       "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
       "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
       "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
       "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed",
       "smsp__average_warp_latency_per_inst_issued.ratio", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
       "sm__cycles_elapsed.max", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    return {h: (v, u) for h, u, v in zip(hdr, units, vals)}


def opcode_mix(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[1]
    isrc, iex = hdr.index("Source"), hdr.index("Instructions Executed")
    mix = collections.Counter()
    for r in rows[2:]:
        if len(r) <= iex:
            continue
        p = r[isrc].split()
        if not p:
            continue
        op = p[1] if p[0].startswith("@") else p[0]
        key = "IMAD.WIDE" if op.startswith("IMAD.WIDE") else op.split(".")[0]
        if op.startswith("IMAD.") and op.split(".")[1] in ("MOV", "IADD", "SHL", "X"):
            key = "IMAD." + op.split(".")[1]
        mix[key] += int(r[iex])
    return mix


def num(x):
    try:
        return float(x.replace(",", ""))
    except Exception:
        return x


def main():
    tag, outp = sys.argv[1], sys.argv[2]
    from bn_b200 import build as b
    res = {"source": "ncu --set full --clock-control none --import-source on, one launch each inside `python bench.py --steps 1 --warmup 3` "
                     "(2^14 pairings); capture " + tag,
           "source_hash": b.source_hash(),  # bench.py only uses this artefact when it matches the build it is timing
           "kernels": {}}
    for spec in sys.argv[3:]:
        name, rep = spec.split("=")
        d = raw(rep)
        mix = opcode_mix(rep)
        k = {m: num(d[m][0]) for m in RAW if m in d}
        k["units"] = {m: d[m][1] for m in RAW if m in d}
        stalls = {h.split("issue_stalled_")[1].split("_per_issue")[0]: num(v[0]) for h, v in d.items()
                  if "issue_stalled" in h and h.endswith("per_issue_active.ratio")}
        k["stall_cycles_per_issued_inst"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:10])
        tot = sum(mix.values())
        k["inst_total"] = tot
        k["inst_imad_wide"] = mix["IMAD.WIDE"]
        k["inst_fp64"] = mix["DFMA"] + mix["DADD"] + mix["DMUL"]
        # ALU-pipe opcodes (half rate: 16 lanes/clk/scheduler, tools/ubench/pipes2.cu)
        k["inst_alu"] = sum(mix[o] for o in ("IADD3", "LOP3", "SHF", "SEL", "LEA", "ISETP", "VIADD", "PRMT", "IABS", "IMNMX", "FLO", "POPC", "BREV", "MOV", "P2R", "R2P", "PLOP3", "FSEL"))
        k["inst_mix_top"] = dict(mix.most_common(14))
        ur = d["dram__bytes_read.sum"][1]
        scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}[ur]
        k["dram_read_bytes"] = int(num(d["dram__bytes_read.sum"][0]) * scale)
        uw = d["dram__bytes_write.sum"][1]
        k["dram_write_bytes"] = int(num(d["dram__bytes_write.sum"][0]) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}[uw])
        k["pairs"] = 16384
        res["kernels"][name] = k
    with open(outp, "w") as f:
        json.dump(res, f, indent=1)
    for name, k in res["kernels"].items():
        print(name, "ms", k["gpu__time_duration.sum"], k["units"]["gpu__time_duration.sum"], "inst", k["inst_total"], "imadw", k["inst_imad_wide"],
              "(%.1f%%)" % (100.0 * k["inst_imad_wide"] / k["inst_total"]), "dram R/W", k["dram_read_bytes"], k["dram_write_bytes"])
        cyc = (4.0 * k["inst_imad_wide"] + 1.77 * (k["inst_total"] - k["inst_imad_wide"])) / 592
        print("   slot-model cycles/SMSP %.3e  -> %.3f ms at 1.965 GHz; fmaheavy %.1f%% alu %.1f%% fp64 %s%% issue %.1f%%" % (
            cyc, cyc / 1.965e6, k["sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed"],
            k["sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed"],
            k.get("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "-"), k["smsp__issue_active.avg.pct_of_peak_sustained_active"]))


main()
