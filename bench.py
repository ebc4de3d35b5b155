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

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = sorted(int(r[1]) for r in self.rows if len(r) >= 9 and r[1].isdigit())
        mx = [int(r[2]) for r in self.rows if len(r) >= 9 and r[2].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[j] for r in self.rows if len(r) >= 9 for j in range(4) if r[5 + j].lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


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


def run_reference(args, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    sample = 128 * cores
    rates, times = [], []
    for i in range(args.warmup + args.steps):
        r, dt = cpu_reference_rate(sample, cores)
        if i >= args.warmup:
            rates.append(r)
            times.append(dt)
    value = sum(rates) / len(rates)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "pairings/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64 (4x64-bit Montgomery limbs)",
        "data": "synthetic",
        "config": {"workload": "optimal-ate pairings, bounded sample of the 2^14-pair batch", "pairs_per_step": sample},
        "cpu_baseline": {"value": value, "unit": "pairings/s", "cores": cores, "kind": "port",
                         "sample": "%d pairs per step on %d host threads; C restatement of the reference algorithm "
                                   "(oracle/bn_ref.c) -- the Rust crate cannot be built here (no rustc/cargo)" % (sample, cores)},
        "e2e": {"value": value, "unit": "pairings/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


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
    # An explicit (non-default) stream: the default stream's handle is NULL, which the C ABI maps to the library's
    # own stream -- torch events would then not see the kernels.
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
    del base1, base2
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
        flush.zero_()  # L2 flush between steps (the 535 MB line buffer streamed per step also exceeds L2)
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

    # clocks / throttle reasons are sampled from the start of the warm-up to the end of the timed region (the GPU is
    # under the same load throughout; the timed region alone can be shorter than nvidia-smi's sampling period)
    sampler = ClockSampler(local_rank)
    sampler.start()
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
    if fused is not None:  # outside the timed region: the fused gather must equal compute + NCCL all_gather, bit for bit
        chk(lib.bn_b200_pairing_batch_dev(dptr(d_g1), dptr(d_g2), dptr(d_out), ctypes.c_size_t(n), sp))
        dist.all_gather_into_tensor(gathered, d_out)
        torch.cuda.synchronize()
        assert torch.equal(gathered, fused.result()), "fused peer-store gather differs from the NCCL all_gather"
    clocks = sampler.stop()
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = world * n / (ms_step * 1e-3)

    # ---- per-kernel timing (same command, CUDA events inside the library, on the launching stream)
    chk(lib.bn_b200_set_profiling(1))
    k_ms = np.zeros((args.steps, 3), dtype=np.float32)
    for i in range(args.steps):
        flush.zero_()
        chk(lib.bn_b200_pairing_batch_dev(dptr(d_g1), dptr(d_g2), dptr(d_out), ctypes.c_size_t(n), sp))
        chk(lib.bn_b200_last_pairing_kernel_ms3(k_ms[i].ctypes.data_as(ctypes.c_void_p)))
    chk(lib.bn_b200_set_profiling(0))
    ms_lines, ms_mil, ms_fexp = float(k_ms[:, 0].mean()), float(k_ms[:, 1].mean()), float(k_ms[:, 2].mean())
    ms_miller = ms_mil + ms_fexp  # Miller loop + final exponentiation (two kernels since run 28)

    # ---- IMAD issue peak, measured live: pure IMAD.WIDE.U32 kernel, 148*k blocks x 256 threads
    scratch = torch.zeros(16, dtype=torch.int32, device=dev)
    sms = lib.bn_b200_sm_count()
    blocks, iters = sms * 8, 4096
    imads = blocks * 256 * iters * 32
    chk(lib.bn_b200_imad_peak_dev(dptr(scratch), blocks, 256, sp))
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(5):
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record(stream)
        chk(lib.bn_b200_imad_peak_dev(dptr(scratch), blocks, iters, sp))
        a1.record(stream)
        torch.cuda.synchronize()
        best = min(best, a0.elapsed_time(a1))
    imad_peak = imads / (best * 1e-3)  # IMAD/s

    # ---- config 2: Fq Montgomery multiplication chain (2^20 lanes x 1024)
    nf, chain = 1 << 20, 1024
    fa = torch.from_numpy(splitmix_scalars(0xB2000002, 1024).view(np.int64)).to(dev).repeat(nf // 1024, 1)
    fb = torch.from_numpy(splitmix_scalars(0xB2000012, 1024).view(np.int64)).to(dev).repeat(nf // 1024, 1)
    fo = torch.empty_like(fa)
    chk(lib.bn_b200_fq_mul_chain_dev(dptr(fa), dptr(fb), dptr(fo), ctypes.c_size_t(nf), 8, sp))
    torch.cuda.synchronize()
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a0.record(stream)
    chk(lib.bn_b200_fq_mul_chain_dev(dptr(fa), dptr(fb), dptr(fo), ctypes.c_size_t(nf), chain, sp))
    a1.record(stream)
    torch.cuda.synchronize()
    fq_mul_rate = nf * chain / (a0.elapsed_time(a1) * 1e-3)
    del fa, fb, fo

    # ---- config 3: 2^16 batched G1 scalar multiplications (random Fr), device resident
    n3 = 1 << 16
    k3 = torch.from_numpy(splitmix_scalars(0xB2000003, 4096).view(np.int64)).to(dev).repeat(n3 // 4096, 1)
    p3 = d_g1.repeat(n3 // n, 1) if n3 >= n else d_g1[:n3]
    o3 = torch.empty_like(p3)
    chk(lib.bn_b200_g1_mul_batch_dev(dptr(p3), dptr(k3), dptr(o3), ctypes.c_size_t(n3), sp))
    torch.cuda.synchronize()
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a0.record(stream)
    chk(lib.bn_b200_g1_mul_batch_dev(dptr(p3), dptr(k3), dptr(o3), ctypes.c_size_t(n3), sp))
    a1.record(stream)
    torch.cuda.synchronize()
    g1_mul_rate = n3 / (a0.elapsed_time(a1) * 1e-3)
    del k3, p3, o3

    # ---- row f-1: fused pairing(...).pow(s), same batch
    d_pw = torch.empty_like(d_out)
    chk(lib.bn_b200_pairing_pow_batch_dev(dptr(d_g1), dptr(d_g2), dptr(ka), dptr(d_pw), ctypes.c_size_t(n), sp))
    torch.cuda.synchronize()
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a0.record(stream)
    chk(lib.bn_b200_pairing_pow_batch_dev(dptr(d_g1), dptr(d_g2), dptr(ka), dptr(d_pw), ctypes.c_size_t(n), sp))
    a1.record(stream)
    torch.cuda.synchronize()
    pairing_pow_rate = n / (a0.elapsed_time(a1) * 1e-3)
    del d_pw

    # ---- e2e: the host-buffer C ABI call a bn-crate user would make; pinned host buffers, H2D+D2H in the timed region
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
    e2e_value = world * n * args.steps / float(t.item())
    assert torch.equal(h_out.to(dev), d_out), "host-buffer path and device-pointer path disagree"

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        rate0, _ = cpu_reference_rate(64 * cores, cores)          # calibration
        sample = max(64 * cores, int(rate0 * 12.0) // cores * cores)  # about 12 s of work on all host threads
        rate, dt = cpu_reference_rate(sample, cores)
        rate1, dt1 = cpu_reference_rate(128, 1)
        cpu = {"value": rate, "unit": "pairings/s", "cores": cores, "kind": "port",
               "sample": "%d pairs on %d host threads in %.1f s (single thread: %.0f pairings/s on 128 pairs); C restatement "
                         "of the reference algorithm, oracle/bn_ref.c" % (sample, cores, dt, rate1),
               "single_thread_value": rate1}

    if rank == 0:
        # algorithmic IMAD/s per kernel (Fq mults of the reference algorithm x 136)
        kern = {"k_pair_lines_duo": (ms_lines, M_LINES), "k_miller": (ms_mil, M_MILLER), "k_fexp": (ms_fexp, M_FEXP)}
        dominant = max(kern, key=lambda k: kern[k][0])
        achieved = n / (kern[dominant][0] * 1e-3) * kern[dominant][1] * IMAD_PER_M
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peaks = json.load(f)
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        # committed ncu capture of this build (profiles/ncu_kernels.json): DRAM traffic and executed-instruction mix per launch
        ncu = {}
        try:
            with open(os.path.join(ROOT, "profiles", "ncu_kernels.json")) as f:
                ncu = json.load(f)
        except Exception:
            pass
        traffic = None
        if dominant in ncu.get("kernels", {}):
            traffic = ncu["kernels"][dominant]["dram_read_bytes"] + ncu["kernels"][dominant]["dram_write_bytes"]
        # Issue-slot model (profiles/r01_run25_ubench_pipes.txt): on B200 the integer ALU and the
        # fma-heavy pipe do not overlap for this instruction mix; a warp instruction costs ~4 issue cycles (IMAD.WIDE)
        # or ~1.77 (everything else) per scheduler.  bound_ms = that sum over the executed instructions of the launch.
        slot = {}
        for k, v in ncu.get("kernels", {}).items():
            if k in kern and "inst_imad_wide" in v:
                cyc = (4.0 * v["inst_imad_wide"] + 1.77 * (v["inst_total"] - v["inst_imad_wide"])) / (lib.bn_b200_sm_count() * 4)
                bound_ms = cyc / (clocks.get("sm_mhz", 1965) * 1e3) * (n / v.get("pairs", n))
                slot[k] = {"bound_ms": bound_ms, "measured_ms": kern[k][0], "frac": bound_ms / kern[k][0]}
        line_bytes = lib.bn_b200_num_lines() * 320
        line = {
            "metric": METRIC, "value": value, "unit": "pairings/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u32 limbs (256-bit Montgomery integers)", "data": "synthetic",
            "config": {"workload": "2^14 batched optimal-ate pairings per GPU (BASELINE config 4; N=8 is config 5's 2^17)",
                       "pairs_per_gpu": n, "global_pairs": world * n, "parallelism": "dp%d (independent pairs)" % world,
                       "gather": gather_mode,
                       "l2": "256 MiB flush write between steps + %d MB line buffer streamed per step (> 126 MB L2)" % (n * lib.bn_b200_num_lines() * 320 // 1000000),
                       "inputs": "P=G1::one()*a, Q=G2::one()*b, Jacobian z!=1, seeds 0xB2000004+rank"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "pairings/s", "h2d_bytes_per_step": n * BYTES_IN,
                    "d2h_bytes_per_step": n * BYTES_OUT,
                    "path": "bn_b200_pairing_batch (host pointers, pinned): H2D + 3 kernels per step, results stored by the last "
                            "kernel's epilogue directly into the pinned output buffer (zero-copy D2H)"},
            "gpu_launches": int(launches),
            "roofline": {
                "bound": "int-imad", "kernel": dominant,
                "achieved": achieved / 1e12, "peak": imad_peak / 1e12, "unit": "TIMAD/s (IMAD.WIDE.U32 32x32+64)",
                "frac": achieved / imad_peak,
                "traffic": traffic,
                "bound_note": "integer modular arithmetic: neither HBM- nor tensor-bound (SURVEY.md section 8d); the denominator is the "
                              "IMAD.WIDE issue rate measured live by k_imad_peak (%d blocks x 256 threads): 8 lanes/clk/SMSP = "
                              "148*4*8*1.965e9 = 9.31e12/s, half of SURVEY.md's nominal 18.6e12 assumption" % blocks,
                "frac_of_survey_nominal_18.6T": achieved / 18.6e12,
                "algorithmic": "Fq mults of the reference algorithm x 136 IMAD: lines %d, Miller loop %d, final exponentiation %d (pairing: %d)"
                               % (M_LINES, M_MILLER, M_FEXP, M_PAIRING),
                "kernels": {k: {"ms": v[0], "achieved_timad": n / (v[0] * 1e-3) * v[1] * IMAD_PER_M / 1e12,
                                "frac": n / (v[0] * 1e-3) * v[1] * IMAD_PER_M / imad_peak} for k, v in kern.items()},
                "kernel_ms": {"k_pair_lines": ms_lines, "k_miller_fexp": ms_miller, "k_miller": ms_mil, "k_fexp": ms_fexp},
                "whole_path_frac": (n / ((ms_lines + ms_miller) * 1e-3)) * M_PAIRING * IMAD_PER_M / imad_peak,
                "issue_slot_model": slot or None,
                "traffic_detail": {"algorithmic_bytes_per_launch": {"k_pair_lines_duo": n * (BYTES_IN + line_bytes),
                                                                    "k_miller": n * (line_bytes + BYTES_OUT), "k_fexp": n * 2 * BYTES_OUT},
                                   "ncu": {k: {"dram_read_bytes": v.get("dram_read_bytes"), "dram_write_bytes": v.get("dram_write_bytes")}
                                           for k, v in ncu.get("kernels", {}).items()}, "source": ncu.get("source")},
                "hbm": {"achieved_gbs": n * (BYTES_IN + BYTES_OUT) / (ms_step * 1e-3) / 1e9,
                        "with_line_buffer_gbs": n * (BYTES_IN + BYTES_OUT + 2 * line_bytes) / (ms_step * 1e-3) / 1e9,
                        "peak_gbs": hbm_peak, "of": "measured" if peaks else "fallback"},
                "fused_pairing_pow": {"config": "2^14 x pairing(P,Q).pow(s) in one pass (row f-1)", "per_s": pairing_pow_rate},
                "g1_scalar_mul": {"config": "2^16 G1 * Fr, random scalars (BASELINE config 3)", "per_s": g1_mul_rate},
                "fq_mul_chain": {"config": "2^20 lanes x 1024 Montgomery muls (BASELINE config 2)",
                                 "fq_mul_per_s": fq_mul_rate, "imad_frac": fq_mul_rate * IMAD_PER_M / imad_peak},
            },
        }
        if cpu:
            line["cpu_baseline"] = cpu
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
