"""bench.py's reporting block on CPU: the JSON line must survive every optional measurement being absent
(the 8-GPU run died on `None * 1e3` when the nvidia-smi sampler caught no sample)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402


def measurements(**over):
    m = {"n": 16384, "world": 8, "steps": 20, "warmup": 5, "value": 1.6e7, "ms_step": 8.0, "launches": 60,
         "gather_mode": "fused", "num_lines": 88, "sm_count": 148, "imad_peak": 9.26e12,
         "kernels": {"k_pair_lines_duo": (1.57, bench.M_LINES), "k_miller": (2.18, bench.M_MILLER), "k_fexp": (3.77, bench.M_FEXP)},
         "e2e_value": 1.5e7,
         "ncu": {"source": "x", "source_hash": "abc", "kernels": {"k_fexp": {"inst_total": 10, "inst_imad_wide": 3, "inst_fp64": 4, "inst_alu": 3, "pairs": 16384,
                                                                                 "dram_read_bytes": 1, "dram_write_bytes": 2}}},
         "source_hash": "abc"}
    m.update(over)
    return m


def test_report_with_unsampled_clocks():
    for clocks in (None, {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0},
                   {"sm_mhz": None, "sm_max_mhz": 1965, "reasons": []}, {"sm_mhz": 1965, "sm_max_mhz": 1965, "reasons": []}):
        line = bench.build_line(measurements(clocks=clocks))
        json.dumps(line)
        assert line["value"] == 1.6e7 and line["n_gpus"] == 8 and "roofline" in line and "roofline_error" not in line
        assert line["roofline"]["pipe_model"]["k_fexp"]["bound_ms"] > 0
        assert line["config"]["global_pairs"] == 8 * 16384


def test_report_without_optional_legs():
    m = measurements(clocks=None)
    for k in ("ncu", "source_hash", "e2e_value"):
        m.pop(k)
    line = bench.build_line(m)
    assert "e2e" not in line and line["roofline"]["traffic"] is None and line["roofline"]["pipe_model"] is None
    # a stale ncu artefact (other build) is not used
    line = bench.build_line(measurements(source_hash="different"))
    assert line["roofline"]["traffic"] is None and line["roofline"]["ncu_artefact"]["matches_this_build"] is False
    # broken kernel timings: the core line survives
    line = bench.build_line(measurements(kernels={}))
    assert line["value"] == 1.6e7 and "roofline_error" in line


def test_clock_summary_handles_empty_and_window():
    assert bench.summarize_clock_rows([])["sm_mhz"] is None
    row = lambda mhz: ["0", str(mhz), "1965", "300", "0x0", "Not Active", "Not Active", "Not Active", "Not Active"]
    rows = [(1.0, row(600)), (2.0, row(1965)), (2.5, row(1950)), (3.0, row(1965)), (9.0, row(500)), (9.5, ["garbage"])]
    s = bench.summarize_clock_rows(rows, (1.5, 3.5))
    assert s["sm_mhz"] == 1965 and s["samples"] == 3 and s["samples_total"] == 5 and s["reasons"] == []
    assert bench.summarize_clock_rows(rows, (100.0, 101.0))["samples"] == 5   # empty window: every sample
    rows.append((2.2, ["0", "1200", "1965", "900", "0x4", "Not Active", "Not Active", "Not Active", "Active"]))
    assert bench.summarize_clock_rows(rows, (1.5, 3.5))["reasons"] == ["sw_power_cap"]


def test_both_arms_print_the_same_config():
    assert bench.make_config(1 << 14, 8) == bench.make_config(1 << 14, 8)
    out = subprocess.check_output([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                                   "--warmup", "1", "--pairs", "64", "--gpus", "2"], text=True, env=dict(os.environ, RANK="0"))
    line = json.loads(out.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["config"] == bench.make_config(64, 2) and line["n_gpus"] == 2
    assert line["cpu_baseline"]["kind"] == "port" and line["e2e"]["h2d_bytes_per_step"] == 0
    # the other ranks of a torchrun launch print nothing and exit 0
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, env=dict(os.environ, RANK="1", WORLD_SIZE="2"))
    assert out.returncode == 0 and out.stdout.strip() == ""
