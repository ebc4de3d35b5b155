#!/usr/bin/env python3
"""bench.py -- pairings/sec of the batched optimal-ate pairing on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one pass of the hot path (line schedule, Miller loop and final-exponentiation kernels) over one batch
of 2^14 synthetic (G1, G2) pairs per GPU (BASELINE config 4; N GPUs => N * 2^14 pairs, config 5's 2^17 at N = 8,
weak scaling, plus an NCCL all-gather of the 384-byte Gt results).  Prints ONE JSON line on rank 0.

  value     whole-job pairings/s, inputs resident in HBM, CUDA events on the launching stream, max over ranks
  e2e       same metric through the host-buffer C ABI call (pinned host inputs, H2D + kernels + D2H each step)
  roofline  dominant kernel (k_fexp) against the integer-multiply (IMAD.WIDE.U32) issue peak measured
            live by the calibration kernel; HBM GB/s reported beside it (the path is not HBM-bound)
  cpu_baseline  the C restatement of the reference algorithm (oracle/bn_ref.c, "port") on this box's host cores

--impl reference times that CPU restatement only (the Rust crate cannot be built in this image: no rustc/cargo).
"""
import argparse
import ctypes
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

PAIRS_PER_GPU = 1 << 14
METRIC = "pairings/sec (batched optimal-ate)"
# algorithmic work (SURVEY.md section 8d / BASELINE.md section 2): 136 IMAD per Fq multiplication
IMAD_PER_M = 136
M_PAIRING = 18995                 # real Fq mults per pairing
M_MILLER = 6690                   # miller_loop (k_miller)
M_FEXP = 227 + 8541               # final exponentiation (k_fexp)
M_MILLER_FEXP = M_MILLER + M_FEXP
M_LINES = 21 + 3516               # to_affine + line precomputation (the k_pair_lines kernel)
BYTES_IN, BYTES_OUT = 288, 384    # per pairing


def splitmix_scalars(seed, n):
    """n pseudo-random scalars mod r as Montgomery Fr images [n,4] (same generator as tests/util.py)."""
    M64 = (1 << 64) - 1
    R_ORDER = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
    out = np.empty((n, 4), dtype=np.uint64)
    for i in range(n):
        state = (seed * 0x9E3779B97F4A7C15 + i * 0xD1B54A32D192ED03) & M64
        v = 0
        for k in range(8):
            state = (state + 0x9E3779B97F4A7C15) & M64
            z = state
            z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M64
            z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M64
            v |= (z ^ (z >> 31)) << (64 * k)
        m = (v % R_ORDER) * (1 << 256) % R_ORDER
        out[i] = np.frombuffer(m.to_bytes(32, "little"), dtype="<u8")
    return out


GEN1 = None


def generators():
    """G1::one(), G2::one() Montgomery images (reference src/groups/mod.rs:356-362, 378-390), derived from (1,2) and
    the standard alt_bn128 G2 generator."""
    Q = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47
    mont = lambda x: np.frombuffer((x * (1 << 256) % Q).to_bytes(32, "little"), dtype="<u8")
    g1 = np.concatenate([mont(1), mont(2), mont(1)])
    g2x = (10857046999023057135944570762232829481370756359578518086990519993285655852781,
           11559732032986387107991004021392285783925812861821192530917403151452391805634)
    g2y = (8495653923123431417604973247489272438418190587263600148770280649306958101930,
           4082367875863433681332203403145435568316851327593401208105741076214120093531)
    g2 = np.concatenate([mont(g2x[0]), mont(g2x[1]), mont(g2y[0]), mont(g2y[1]), mont(1), mont(0)])
    return g1.astype(np.uint64), g2.astype(np.uint64)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index, period_ms=20):
        self.idx = gpu_index
        self.period_ms = period_ms
        self.rows = []      # (host time, fields)
        self.proc = None
        self.t_load = [None, None]   # host-time window in which the GPU is known to be under load

    def mark_load_start(self):
        self.t_load[0] = time.time()

    def mark_load_end(self):
        self.t_load[1] = time.time()

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", str(self.period_ms), "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        return summarize_clock_rows(self.rows, self.t_load)


def summarize_clock_rows(rows, window=(None, None)):
    """rows: (host time, nvidia-smi csv fields).  Median SM clock over the samples taken while the GPU was under load
    (the marked window; every sample if the window caught none).  Never raises; an unsampled run gives sm_mhz None."""
    good = [(t, r) for t, r in rows if len(r) >= 9 and r[1].isdigit()]
    lo, hi = window
    inside = [(t, r) for t, r in good if (lo is None or t >= lo) and (hi is None or t <= hi + 0.05)]
    use = inside or good
    sm = sorted(int(r[1]) for _, r in use)
    mx = [int(r[2]) for _, r in good if r[2].isdigit()]
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    reasons = sorted({names[j] for _, r in use for j in range(4) if r[5 + j].lower().startswith("active")})
    return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
            "reasons": reasons, "samples": len(sm), "samples_total": len(good),
            "window": "warm-up start .. end of the device-side measurement legs (GPU under load throughout)"}


def cpu_reference_rate(sample, threads):
    """pairings/s of the C restatement of the reference algorithm on `threads` host threads (bounded sample)."""
    from oracle import cref
    from tests import util
    g1, g2 = util.synth_pairs(0xB2000001, min(sample, 64), threads)
    reps = (sample + len(g1) - 1) // len(g1)
    g1 = np.tile(g1, (reps, 1))[:sample]
    g2 = np.tile(g2, (reps, 1))[:sample]
    cref.pairing_batch(g1[:threads], g2[:threads], threads)  # warm
    t0 = time.perf_counter()
    cref.pairing_batch(g1, g2, threads)
    dt = time.perf_counter() - t0
    return sample / dt, dt


def make_config(n, world, num_lines=88):
    """The `config` object both arms print (identical for the same N, so the driver's same_config check holds)."""
    return {"workload": "2^14 batched optimal-ate pairings per GPU (BASELINE config 4; N=8 is config 5's 2^17)",
            "pairs_per_gpu": n, "global_pairs": world * n, "parallelism": "dp%d (independent pairs)" % world,
            "l2": "256 MiB flush write between steps (> 126 MB L2)",
            "inputs": "P=G1::one()*a, Q=G2::one()*b, Jacobian z!=1, seeds 0xB2000004+rank"}


def run_reference(args, rank, world):
    """--impl reference: the reference algorithm's CPU implementation (C restatement, all host threads) on the SAME
    pairs-per-step as our arm (2^14); rank 0 only."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    sample = args.pairs
    rates, times = [], []
    for i in range(args.warmup + args.steps):
        r, dt = cpu_reference_rate(sample, cores)
        if i >= args.warmup:
            rates.append(r)
            times.append(dt)
    value = sample * len(times) / sum(times)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "pairings/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64 (4x64-bit Montgomery limbs)",
        "data": "synthetic",
        "config": make_config(args.pairs, max(args.gpus, 1)),
        "cpu_baseline": {"value": value, "unit": "pairings/s", "cores": cores, "kind": "port",
                         "sample": "%d pairs per step (one GPU's shard of the batch) on %d host threads; C restatement of the "
                                   "reference algorithm (oracle/bn_ref.c) -- the Rust crate cannot be built here (no rustc/cargo)"
                                   % (sample, cores)},
        "e2e": {"value": value, "unit": "pairings/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def build_line(m):
    """Assemble the JSON line from the measurements dict `m` (plain Python values; no torch).  The core keys come first and
    cannot fail on missing optional measurements: anything optional (clocks, ncu artefact, cpu baseline) may be None."""
    n, world, steps = m["n"], m["world"], m["steps"]
    clocks = m.get("clocks") or {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"], "samples": 0}
    line = {
        "metric": METRIC, "value": m["value"], "unit": "pairings/s", "n_gpus": world, "steps": steps,
        "warmup": m["warmup"], "ms_per_step": m["ms_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32 limbs (256-bit Montgomery integers)", "data": "synthetic",
        "config": make_config(n, world, m.get("num_lines", 88)),
        "gather": m.get("gather_mode", "none"),
        "clocks": clocks,
        "gpu_launches": int(m.get("launches", 0)),
    }
    if m.get("e2e_value") is not None:
        line["e2e"] = {"value": m["e2e_value"], "unit": "pairings/s", "h2d_bytes_per_step": n * BYTES_IN,
                       "d2h_bytes_per_step": n * BYTES_OUT,
                       "path": m.get("e2e_path", "bn_b200_pairing_batch (host pointers, pinned)")}
    if m.get("parity") is not None:
        line["parity_checked"] = m["parity"]
    try:
        line["roofline"] = build_roofline(m, clocks)
    except Exception as e:  # noqa: BLE001 -- a reporting problem must never lose a finished measurement
        line["roofline_error"] = "%s: %s" % (type(e).__name__, e)
    if m.get("cpu"):
        line["cpu_baseline"] = m["cpu"]
    return line


def build_roofline(m, clocks):
    n = m["n"]
    kern = {k: v for k, v in m["kernels"].items() if v[0]}  # name -> (ms, algorithmic Fq mults per pairing)
    imad_peak = m["imad_peak"]
    dominant = max(kern, key=lambda k: kern[k][0])
    achieved = n / (kern[dominant][0] * 1e-3) * kern[dominant][1] * IMAD_PER_M
    peaks = m.get("peaks") or {}
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    ncu = m.get("ncu") or {}
    ncu_k = ncu.get("kernels", {})
    fresh = bool(ncu) and ncu.get("source_hash") is not None and ncu.get("source_hash") == m.get("source_hash")
    traffic = None
    if dominant in ncu_k and fresh:
        traffic = ncu_k[dominant]["dram_read_bytes"] + ncu_k[dominant]["dram_write_bytes"]
    # Pipe model (tools/ubench/pipes2.cu): scheduler cycles per warp instruction are 4.1 for
    # IMAD.WIDE (fmaheavy), 2.19 for DFMA/DADD (FP64 pipe), 2.0 for ALU-pipe integer instructions.  With independent
    # instruction streams the pipes run concurrently (bound_ms = the busiest pipe); in these kernels (2-6 warps of dependent
    # code per scheduler) the three intervals ADD UP (sum_ms), and a launch lasts as long as its busiest scheduler:
    # model_ms = sum_ms * ceil(warps per scheduler) / (warps per scheduler).  Instruction counts: ncu source page of THIS build.
    sm_mhz = clocks.get("sm_mhz") or clocks.get("sm_max_mhz") or 1965
    slot = {}
    for k, v in ncu_k.items():
        if k in kern and "inst_imad_wide" in v and fresh:
            smsp = m.get("sm_count", 148) * 4
            scale = (n / v.get("pairs", n)) / smsp / (sm_mhz * 1e3)
            pipes = {"fmaheavy_imad_wide": 4.1 * v["inst_imad_wide"] * scale, "fp64": 2.19 * v.get("inst_fp64", 0) * scale,
                     "alu": 2.0 * v.get("inst_alu", 0) * scale}
            bound_ms = max(pipes.values())
            sum_ms = sum(pipes.values())
            # warps of the launch: a lane pair per pairing (line kernel), five 6-lane hexads per warp (the others)
            warps = math.ceil(2 * n / 32) if "lines" in k else math.ceil(n / 5)
            wps = warps / smsp
            model_ms = sum_ms * math.ceil(wps - 1e-9) / wps
            slot[k] = {"bound_ms": bound_ms, "busiest_pipe": max(pipes, key=pipes.get), "pipe_ms": pipes, "measured_ms": kern[k][0],
                       "frac": bound_ms / kern[k][0], "sum_ms": sum_ms, "warps_per_scheduler": wps, "model_ms": model_ms,
                       "model_over_measured": model_ms / kern[k][0], "inst_total": v["inst_total"],
                       "issue_active_pct": v.get("smsp__issue_active.avg.pct_of_peak_sustained_active")}
    total_ms = sum(v[0] for v in kern.values())
    line_bytes = m.get("line_bytes_per_pairing", 0)
    r = {
        "bound": "int-imad", "kernel": dominant,
        "achieved": achieved / 1e12, "peak": imad_peak / 1e12, "unit": "TIMAD/s (IMAD.WIDE.U32 32x32+64)",
        "frac": achieved / imad_peak,
        "traffic": traffic,
        "bound_note": "integer modular arithmetic: neither HBM- nor tensor-bound (SURVEY.md section 8d); the denominator is the "
                      "IMAD.WIDE issue rate measured live by k_imad_peak: 8 lanes/clk/SMSP = 148*4*8*1.965e9 = 9.31e12/s, half of "
                      "SURVEY.md's nominal 18.6e12 assumption (tools/ubench/pipes2.cu times every 32x32 multiply form: none is cheaper per product bit)",
        "frac_of_survey_nominal_18.6T": achieved / 18.6e12,
        "algorithmic": "Fq mults of the reference algorithm x 136 IMAD: lines %d, Miller loop %d, final exponentiation %d (pairing: %d)"
                       % (M_LINES, M_MILLER, M_FEXP, M_PAIRING),
        "kernels": {k: {"ms": v[0], "achieved_timad": n / (v[0] * 1e-3) * v[1] * IMAD_PER_M / 1e12,
                        "frac": n / (v[0] * 1e-3) * v[1] * IMAD_PER_M / imad_peak} for k, v in kern.items()},
        "kernels_note": "per-kernel times: CUDA events inside the library around ONE sequence of full-size kernels (profiling mode, as "
                        "the ncu captures); in the timed steps a call of this size runs as two sub-batches on two streams whose kernels "
                        "overlap (DESIGN.md section 5), so ms_per_step may be below the sum of these times",
        "whole_path_frac": (n / (total_ms * 1e-3)) * M_PAIRING * IMAD_PER_M / imad_peak,
        "whole_path_frac_of": "the sum of the per-kernel times above (one sequence of kernels), not ms_per_step",
        "pipe_model": slot or None,
        "ncu_artefact": {"file": "profiles/ncu_kernels.json", "capture": ncu.get("source"), "matches_this_build": fresh},
        "traffic_detail": {"algorithmic_bytes_per_pairing": BYTES_IN + BYTES_OUT, "scratch_bytes_per_pairing": line_bytes,
                           "ncu": {k: {"dram_read_bytes": v.get("dram_read_bytes"), "dram_write_bytes": v.get("dram_write_bytes")}
                                   for k, v in ncu_k.items()} if fresh else None},
        "hbm": {"achieved_gbs": n * (BYTES_IN + BYTES_OUT) / (m["ms_step"] * 1e-3) / 1e9,
                "peak_gbs": hbm_peak, "of": "measured" if peaks else "fallback"},
    }
    for key in ("fused_pairing_pow", "g1_scalar_mul", "g2_scalar_mul", "fq_mul_chain", "fq_sqr_chain"):
        if m.get(key) is not None:
            r[key] = m[key]
    return r


def source_hash():
    from bn_b200 import build as b
    return b.source_hash()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pairs", type=int, default=PAIRS_PER_GPU, help="pairs per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--gather", default="auto", choices=["auto", "fused", "nccl"],
                    help="N > 1: fused = results stored by the kernel epilogue into every peer's buffer over NVLink; "
                         "nccl = all_gather after the kernels; auto = fused when peer memory is available")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    import bn_b200
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (bn_b200 has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    lib = bn_b200.init(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    n = args.pairs
    # clocks / throttle reasons: sampled from before input generation to the end of the device-side legs; the median is
    # taken over the window in which the GPU is under load (warm-up start .. end of the per-kernel / calibration legs)
    sampler = ClockSampler(local_rank)
    sampler.start()
    # An explicit (non-default) stream: torch events must see the kernels.
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    sp = ctypes.c_void_p(stream.cuda_stream)

    def dptr(t):
        return ctypes.c_void_p(t.data_ptr())

    def chk(rc):
        bn_b200._lib.check(rc)

    # ---- synthetic inputs, generated on the device by the product's own scalar-mul kernels (row f-2):
    #      P_i = G1::one() * a_i, Q_i = G2::one() * b_i, Jacobian with z != 1 (mirrors G::random, groups/mod.rs:220-222)
    g1gen, g2gen = generators()
    seed = 0xB2000004 + 7919 * rank
    ka = torch.from_numpy(splitmix_scalars(seed, n).view(np.int64)).to(dev)
    kb = torch.from_numpy(splitmix_scalars(seed ^ 0x5555, n).view(np.int64)).to(dev)
    base1 = torch.from_numpy(np.tile(g1gen, (n, 1)).view(np.int64)).to(dev)
    base2 = torch.from_numpy(np.tile(g2gen, (n, 1)).view(np.int64)).to(dev)
    d_g1 = torch.empty((n, 12), dtype=torch.int64, device=dev)
    d_g2 = torch.empty((n, 24), dtype=torch.int64, device=dev)
    chk(lib.bn_b200_g1_mul_batch_dev(dptr(base1), dptr(ka), dptr(d_g1), ctypes.c_size_t(n), sp))
    chk(lib.bn_b200_g2_mul_batch_dev(dptr(base2), dptr(kb), dptr(d_g2), ctypes.c_size_t(n), sp))
    torch.cuda.synchronize()
    d_out = torch.empty((n, 48), dtype=torch.int64, device=dev)
    gathered = torch.empty((world * n, 48), dtype=torch.int64, device=dev) if world > 1 else None
    fused, gather_mode = None, "none"
    if world > 1:
        gather_mode = "NCCL all_gather of Gt (384 B/pair)"
        if args.gather in ("auto", "fused"):
            try:
                from bn_b200.dist import FusedGather
                fused = FusedGather(n, dev)
                gather_mode = "fused: k_fexp_gather stores each Gt into every peer's buffer over NVLink (symmetric memory) + barrier"
            except Exception as e:  # noqa: BLE001
                if args.gather == "fused":
                    raise
                if rank == 0:
                    print("bench.py: fused gather unavailable (%s); using NCCL all_gather" % e, file=sys.stderr)
        # all ranks must take the same path
        flag = torch.tensor([1 if fused is not None else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0 and fused is not None:
            fused, gather_mode = None, "NCCL all_gather of Gt (384 B/pair)"
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def step():
        flush.zero_()  # L2 flush between steps
        if fused is not None:
            fused.pairing_batch(lib, d_g1, d_g2, sp)
            fused.barrier()
            return
        chk(lib.bn_b200_pairing_batch_dev(dptr(d_g1), dptr(d_g2), dptr(d_out), ctypes.c_size_t(n), sp))
        if world > 1:
            dist.all_gather_into_tensor(gathered, d_out)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler.mark_load_start()
    for _ in range(args.warmup):
        step()
    barrier()

    # ---- timed region: exactly K steps, CUDA events on the launching stream
    l0 = lib.bn_b200_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches = lib.bn_b200_launch_count() - l0
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = world * n / (ms_step * 1e-3)
    m = {"n": n, "world": world, "steps": args.steps, "warmup": args.warmup, "value": value, "ms_step": ms_step,
         "launches": launches, "gather_mode": gather_mode, "num_lines": lib.bn_b200_num_lines(),
         "sm_count": lib.bn_b200_sm_count(), "kernels": {}, "imad_peak": 9.26e12}

    # Everything below explains / cross-checks the number above.  A failure there is reported in the JSON line
    # ("extras_error") and on stderr, but can no longer lose the measurement or change the exit code of the timed run;
    # PARITY failures are the exception: a wrong result must fail the run.
    parity_failed = None
    try:
        # ---- parity, outside the timed region: (i) fused gather == compute + NCCL all_gather, bit for bit;
        # (ii) every rank compares 256 of its own results with the ORACLE; (iii) rank 0 checks 32 results per peer
        # in the GATHERED buffer against the oracle (the peers' inputs for those indices are gathered too).
        chk(lib.bn_b200_pairing_batch_dev(dptr(d_g1), dptr(d_g2), dptr(d_out), ctypes.c_size_t(n), sp))
        if world > 1:
            dist.all_gather_into_tensor(gathered, d_out)
        torch.cuda.synchronize()
        if fused is not None and not torch.equal(gathered, fused.result()):
            parity_failed = "fused peer-store gather differs from the NCCL all_gather"
        from oracle import cref
        thr = max(1, (os.cpu_count() or 1) // max(world, 1))
        own_idx = torch.arange(0, n, max(1, n // 256), device=dev)[:256]
        g1s = d_g1[own_idx].cpu().numpy().view(np.uint64)
        g2s = d_g2[own_idx].cpu().numpy().view(np.uint64)
        mine = (fused.result()[rank * n:(rank + 1) * n] if fused is not None else d_out)[own_idx].cpu().numpy().view(np.uint64)
        if not np.array_equal(mine, cref.pairing_batch(g1s, g2s, thr)):
            parity_failed = "rank %d: GPU results differ from the oracle" % rank
        m["parity"] = {"own_results_vs_oracle": int(len(own_idx)), "per_rank": True}
        if world > 1:
            pidx = torch.arange(7, n, max(1, n // 32), device=dev)[:32]
            s1 = torch.empty((world * len(pidx), 12), dtype=torch.int64, device=dev)
            s2 = torch.empty((world * len(pidx), 24), dtype=torch.int64, device=dev)
            dist.all_gather_into_tensor(s1, d_g1[pidx].contiguous())
            dist.all_gather_into_tensor(s2, d_g2[pidx].contiguous())
            if rank == 0:
                full = fused.result() if fused is not None else gathered
                gidx = torch.cat([pidx + r * n for r in range(world)])
                want = cref.pairing_batch(s1.cpu().numpy().view(np.uint64), s2.cpu().numpy().view(np.uint64), os.cpu_count() or 1)
                if not np.array_equal(full[gidx].cpu().numpy().view(np.uint64), want):
                    parity_failed = "gathered buffer differs from the oracle"
                m["parity"]["gathered_vs_oracle_per_peer"] = int(len(pidx))
            f2 = torch.tensor([1 if parity_failed else 0], device=dev)
            dist.all_reduce(f2, op=dist.ReduceOp.MAX)
            if int(f2.item()) and not parity_failed:
                parity_failed = "a peer rank reported a parity failure"

        # ---- per-kernel timing (same command, CUDA events inside the library, on the launching stream)
        chk(lib.bn_b200_set_profiling(1))
        k_ms = np.zeros((args.steps, 3), dtype=np.float32)
        for i in range(args.steps):
            flush.zero_()
            chk(lib.bn_b200_pairing_batch_dev(dptr(d_g1), dptr(d_g2), dptr(d_out), ctypes.c_size_t(n), sp))
            chk(lib.bn_b200_last_pairing_kernel_ms3(k_ms[i].ctypes.data_as(ctypes.c_void_p)))
        chk(lib.bn_b200_set_profiling(0))
        names = bn_b200.pairing_kernel_names(lib)
        algo = {"lines": M_LINES, "miller": M_MILLER, "fexp": M_FEXP}
        m["kernels"] = {names[i]: (float(k_ms[:, i].mean()), algo[k]) for i, k in enumerate(("lines", "miller", "fexp"))}
        m["line_bytes_per_pairing"] = int(lib.bn_b200_scratch_bytes_per_pairing())

        def timed(fn, reps=1):
            fn()
            torch.cuda.synchronize()
            best = 1e30
            for _ in range(reps):
                a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a0.record(stream)
                fn()
                a1.record(stream)
                torch.cuda.synchronize()
                best = min(best, a0.elapsed_time(a1))
            return best * 1e-3

        # ---- IMAD issue peak, measured live: pure IMAD.WIDE.U32 kernel, 148*k blocks x 256 threads
        scratch = torch.zeros(16, dtype=torch.int32, device=dev)
        sms = lib.bn_b200_sm_count()
        blocks, iters = sms * 8, 4096
        m["imad_peak"] = blocks * 256 * iters * 32 / timed(lambda: chk(lib.bn_b200_imad_peak_dev(dptr(scratch), blocks, iters, sp)), 5)

        # ---- config 2: Fq Montgomery multiplication chain (2^20 lanes x 1024)
        nf, chain = 1 << 20, 1024
        fa = torch.from_numpy(splitmix_scalars(0xB2000002, 1024).view(np.int64)).to(dev).repeat(nf // 1024, 1)
        fb = torch.from_numpy(splitmix_scalars(0xB2000012, 1024).view(np.int64)).to(dev).repeat(nf // 1024, 1)
        fo = torch.empty_like(fa)
        dt = timed(lambda: chk(lib.bn_b200_fq_mul_chain_dev(dptr(fa), dptr(fb), dptr(fo), ctypes.c_size_t(nf), chain, sp)))
        m["fq_mul_chain"] = {"config": "2^20 lanes x 1024 Montgomery muls (BASELINE config 2)", "fq_mul_per_s": nf * chain / dt,
                             "imad_frac": nf * chain / dt * IMAD_PER_M / m["imad_peak"]}
        del fa, fb, fo

        # ---- config 3: 2^16 batched G1 scalar multiplications (random Fr), device resident
        n3 = 1 << 16
        k3 = torch.from_numpy(splitmix_scalars(0xB2000003, 4096).view(np.int64)).to(dev).repeat(n3 // 4096, 1)
        p3 = d_g1.repeat(n3 // n, 1) if n3 >= n else d_g1[:n3]
        o3 = torch.empty_like(p3)
        dt = timed(lambda: chk(lib.bn_b200_g1_mul_batch_dev(dptr(p3), dptr(k3), dptr(o3), ctypes.c_size_t(n3), sp)))
        m["g1_scalar_mul"] = {"config": "2^16 G1 * Fr, random scalars (BASELINE config 3)", "per_s": n3 / dt,
                              "imad_frac": n3 / dt * 3800 * IMAD_PER_M / m["imad_peak"]}
        del k3, p3, o3

        # ---- row f-1: fused pairing(...).pow(s), same batch
        d_pw = torch.empty_like(d_out)
        dt = timed(lambda: chk(lib.bn_b200_pairing_pow_batch_dev(dptr(d_g1), dptr(d_g2), dptr(ka), dptr(d_pw), ctypes.c_size_t(n), sp)))
        m["fused_pairing_pow"] = {"config": "2^14 x pairing(P,Q).pow(s) in one pass (row f-1)", "per_s": n / dt}
        del d_pw
    except Exception as e:  # noqa: BLE001
        m["extras_error"] = "%s: %s" % (type(e).__name__, e)
        print("bench.py: diagnostic legs failed: %r" % (e,), file=sys.stderr)
    sampler.mark_load_end()
    m["clocks"] = sampler.stop()

    # ---- e2e: the host-buffer C ABI call a bn-crate user would make; pinned host buffers, H2D+D2H in the timed region
    try:
        h_g1 = torch.empty((n, 12), dtype=torch.int64).pin_memory()
        h_g2 = torch.empty((n, 24), dtype=torch.int64).pin_memory()
        h_out = torch.empty((n, 48), dtype=torch.int64).pin_memory()
        h_g1.copy_(d_g1)
        h_g2.copy_(d_g2)

        def e2e_step():
            chk(lib.bn_b200_pairing_batch(dptr(h_g1), dptr(h_g2), dptr(h_out), ctypes.c_size_t(n)))

        for _ in range(2):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()
        barrier()
        e2e_s = time.perf_counter() - t0
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        m["e2e_value"] = world * n * args.steps / float(t.item())
        m["e2e_path"] = ("bn_b200_pairing_batch (host pointers, pinned): H2D + kernels per step, results stored by the last "
                         "kernel's epilogue directly into the pinned output buffer (zero-copy D2H)")
        if not torch.equal(h_out.to(dev), d_out):
            parity_failed = parity_failed or "host-buffer path and device-pointer path disagree"
    except Exception as e:  # noqa: BLE001
        m["extras_error"] = (m.get("extras_error", "") + " | e2e: %s: %s" % (type(e).__name__, e)).strip(" |")
        print("bench.py: e2e leg failed: %r" % (e,), file=sys.stderr)

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            cores = os.cpu_count() or 1
            rate0, _ = cpu_reference_rate(64 * cores, cores)          # calibration
            sample = max(64 * cores, int(rate0 * 12.0) // cores * cores)  # about 12 s of work on all host threads
            rate, dt = cpu_reference_rate(sample, cores)
            rate1, dt1 = cpu_reference_rate(128, 1)
            m["cpu"] = {"value": rate, "unit": "pairings/s", "cores": cores, "kind": "port",
                        "sample": "%d pairs on %d host threads in %.1f s (single thread: %.0f pairings/s on 128 pairs); C restatement "
                                  "of the reference algorithm, oracle/bn_ref.c" % (sample, cores, dt, rate1),
                        "single_thread_value": rate1}
        except Exception as e:  # noqa: BLE001
            print("bench.py: cpu baseline failed: %r" % (e,), file=sys.stderr)

    if rank == 0:
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                m["peaks"] = json.load(f)
        except Exception:
            pass
        try:
            with open(os.path.join(ROOT, "profiles", "ncu_kernels.json")) as f:
                m["ncu"] = json.load(f)
            m["source_hash"] = source_hash()
        except Exception:
            pass
        try:
            line = build_line(m)
        except Exception as e:  # noqa: BLE001 -- last resort: the core measurement alone
            line = {"metric": METRIC, "value": value, "unit": "pairings/s", "n_gpus": world, "steps": args.steps,
                    "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
                    "vs_baseline": None, "data": "synthetic", "config": make_config(n, world),
                    "gpu_launches": int(launches), "report_error": "%s: %s" % (type(e).__name__, e)}
        if m.get("extras_error"):
            line["extras_error"] = m["extras_error"]
        if parity_failed:
            line["parity_failed"] = parity_failed
        print(json.dumps(line), flush=True)
    if world > 1:
        try:
            dist.barrier()
            dist.destroy_process_group()
        except Exception:  # noqa: BLE001
            pass
    if parity_failed:
        raise SystemExit("bench.py: PARITY FAILURE: " + parity_failed)


if __name__ == "__main__":
    main()
