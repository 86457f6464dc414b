"""The bench.py output contract, checked on the committed B200 lines (profiles/) so that a CPU-only run notices a
key that went missing: one JSON object per run with the driver's keys, the roofline / cpu_baseline / e2e blocks of
the task contract, and the multi-GPU lines of the same metric."""
import json
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "roofline", "e2e", "gpu_launches", "clocks"}


def _line(name):
    txt = (ROOT / "profiles" / name).read_text().strip().splitlines()
    return json.loads([l for l in txt if l.startswith("{")][-1])


def test_single_gpu_line_has_the_contract_keys():
    d = _line("r1_bench_v3.json")
    assert BASE_KEYS <= set(d), BASE_KEYS - set(d)
    assert d["metric"] == "qeq_solve_time_n20k_sto_fp64" and d["unit"] == "s" and d["higher_is_better"] is False
    assert d["n_gpus"] == 1 and d["dtype"] == "f64" and d["data"] == "synthetic" and d["warmup"] >= 3
    assert "workload" in d["config"] and "S20k" in d["config"]["workload"] and "model" not in d["config"]
    r = d["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r) and r["bound"] == "tensor"
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and 0.5 < r["frac"] < 1.0
    c = d["cpu_baseline"]
    assert {"value", "unit", "cores", "kind", "sample"} <= set(c) and c["kind"] == "port" and c["cores"] >= 1
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["unit"] == "s"
    assert d["gpu_launches"] > 0
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert d["value"] < 1.0                        # north_star: well under a second on one B200
    assert c["value"] / d["e2e"]["value"] > 50     # reported beside the CPU path, not the target


def test_reference_arm_line():
    d = _line("r1_bench_v3_reference_arm.json")
    assert d["impl"] == "reference" and d["metric"] == "qeq_solve_time_n20k_sto_fp64"
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["cpu_baseline"]["value"] == d["value"] and d["gpu_launches"] == 0


@pytest.mark.parametrize("name, gpus", [("r1_bench_dist2_v4_final.json", 2), ("r1_bench_dist4_v3_dma_flat.json", 4),
                                        ("r1_bench_dist8_v4_final.json", 8)])
def test_multi_gpu_lines_are_strong_scaling_of_the_same_metric(name, gpus):
    d = _line(name)
    one = _line("r1_bench_v3.json")
    assert BASE_KEYS <= set(d)
    assert d["n_gpus"] == gpus and d["scaling"] == "strong" and d["metric"] == one["metric"]
    assert d["config"]["workload"] == one["config"]["workload"]          # the SAME system, solved together
    assert d["value"] < one["value"]                                     # and faster than on one GPU


# ---- round 2 lines ----------------------------------------------------------------------------------------------
def test_round2_single_gpu_line_carries_parity_and_a_measured_cpu_arm():
    d = _line("r2_bench_v2_f32R_trusted.json")
    assert BASE_KEYS <= set(d), BASE_KEYS - set(d)
    assert d["metric"] == "qeq_solve_time_n20k_sto_fp64" and d["n_gpus"] == 1 and d["scaling"] == "strong"
    assert d["value"] < 0.45                        # <= 0.45 s on one B200
    p = d["parity"]
    assert p["ok"] and p["iterations"] == p["iterations_ref"] == 18
    assert p["dq"] <= 1e-8 and p["dmu"] <= 1e-8 and p["max_delta_trace_diff"] <= 1e-9
    c = d["cpu_baseline"]
    assert c["kind"] == "port" and c["cores"] >= 1 and "ONE complete CPU solve" in c["sample"]
    assert d["e2e"]["h2d_bytes_per_step"] == 20001 * 49      # xyz + chi + hardness + radius + n
    m = d["roofline_matvec"]
    assert m["bound"] == "hbm" and m["peak"] and 0 < m["frac"] < 1
    assert "1" in d["config"]["scf_modes"].split()[0]        # stale-factor iterations really ran
    one_r1 = _line("r1_bench_v3.json")
    assert d["value"] < 0.5 * one_r1["value"]


@pytest.mark.parametrize("name, gpus", [("r2_bench_dist2_v1.json", 2), ("r2_bench_dist8_v1.json", 8)])
def test_round2_multi_gpu_lines_check_themselves_against_a_single_gpu_solve(name, gpus):
    d = _line(name)
    assert d["n_gpus"] == gpus and d["scaling"] == "strong"
    chk = d["dist_check"]
    assert chk["dq"] < 1e-10 and chk["dmu"] < 1e-10 and chk["iterations"][0] == chk["iterations"][1]
    assert d["value"] < _line("r2_bench_v2_f32R_trusted.json")["value"]


def test_round2_s100k_line_is_a_bench_py_line():
    d = _line("r2_bench_s100k_8gpu_v1.json")
    assert d["metric"] == "qeq_solve_time_n100k_sto_fp64" and d["n_gpus"] == 8 and "S100k" in d["config"]["workload"]
    assert d["config"]["factor_tflops_overall"] / d["roofline"]["peak"] >= 0.5   # north_star: >= 50 % of FP64 peak


# ---- round 2, tcgen05 slice-product updates ------------------------------------------------------------------------
def test_tcgen05_line_reports_the_new_dominant_kernel_against_the_int8_peak():
    d = _line("r2_bench_v5_final.json")
    assert BASE_KEYS <= set(d), BASE_KEYS - set(d)
    assert d["metric"] == "qeq_solve_time_n20k_sto_fp64" and d["n_gpus"] == 1 and d["warmup"] >= 3
    assert d["value"] < 0.2 and d["e2e"]["value"] < 0.2          # was 0.299 s with FP64 DMMA only, 0.767 s in round 1
    p = d["parity"]
    assert p["ok"] and p["iterations"] == p["iterations_ref"] == 18 and p["dq"] <= 1e-8 and p["dmu"] <= 1e-8
    assert p["max_delta_trace_diff"] <= 1e-9
    r = d["roofline"]
    assert "tcgen05" in r["kernel"] and r["bound"] == "tensor" and r["unit"] == "TFLOP/s"
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and 0.3 < r["frac"] < 1.0
    assert r["peak"] >= 2 * 1500                                  # int8 dense tensor peak: at least 2 x the bf16 figure
    assert r["fp64_equivalent_tflops"] > 2 * 35.4                 # more than twice what any FP64 GEMM can do on this part
    assert r["traffic"] and r["largest_launch"]["tensor_pipe_active"] > 0.8
    o = d["roofline_other_gemm"]
    assert "DMMA" in o["kernel"] and 0 < o["frac"] < 1
    assert d["e2e"]["clocks"]["reasons"] == [] and d["clocks"]["reasons"] == []


@pytest.mark.parametrize("name, gpus", [("r2_bench_dist2_v2_tcgen05.json", 2), ("r2_bench_dist8_v3_tcgen05.json", 8)])
def test_tcgen05_multi_gpu_lines(name, gpus):
    d = _line(name)
    assert d["n_gpus"] == gpus and d["scaling"] == "strong"
    chk = d["dist_check"]
    assert chk["dq"] < 1e-10 and chk["dmu"] < 1e-10 and chk["iterations"][0] == chk["iterations"][1]
    assert d["value"] < _line("r2_bench_v5_final.json")["value"]              # faster than one GPU ...
    assert d["value"] < _line({2: "r2_bench_dist2_v1.json", 8: "r2_bench_dist8_v1.json"}[gpus])["value"]   # ... and than DMMA
