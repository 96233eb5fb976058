"""The bench lines committed under profiles/ (produced by bench.py on a B200) carry every key the bench contract names;
a CPU-only guard against the JSON line drifting away from the contract."""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load(name):
    return json.load(open(os.path.join(ROOT, "profiles", name)))


def test_headline_line_has_the_contract_keys():
    d = _load("r1_bench_e.json")
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
              "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["n_gpus"] == 1 and d["scaling"] == "weak" and d["higher_is_better"] is True and d["vs_baseline"] is None and d["dtype"] == "f32"
    assert "workload" in d["config"] and "model" not in d["config"]
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(d["e2e"])
    assert d["e2e"]["h2d_bytes_per_step"] == 16 * (1 << 20) and d["e2e"]["d2h_bytes_per_step"] == 53 * (1 << 20)
    r = d["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r) and r["bound"] == "hbm" and r["unit"] == "GB/s"
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    # achieved = algorithmic bytes per env-step x envs per launch / kernel time
    assert abs(r["achieved"] - 405 * (1 << 20) / (r["kernel_ms"] * 1e-3) / 1e9) < 1e-3 * r["achieved"]
    assert d["gpu_launches"] == d["steps"] and d["gpu_launches"] > 0
    c = d["cpu_baseline"]
    assert {"value", "unit", "cores", "kind", "sample"} <= set(c) and c["kind"] in ("port", "reference")
    assert abs(d["value"] - (1 << 20) / (d["ms_per_step"] * 1e-3)) < 1e-6 * d["value"]


def test_reference_arm_line():
    d = _load("r1_bench_reference_arm.json")
    assert d["impl"] == "reference" and d["metric"] == "env_steps_per_sec" and d["gpu_launches"] == 0
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1


def test_multi_gpu_lines_are_whole_job_aggregates():
    one = _load("r1_bench_e.json")["value"]
    for name, n in (("r1_bench_2gpu_final.json", 2), ("r1_bench_8gpu_pdl.json", 8)):
        d = _load(name)
        assert d["n_gpus"] == n and d["scaling"] == "weak"
        assert 0.9 * n * one < d["value"] < 1.1 * n * one  # no collective on the step path: linear within run-to-run noise


def test_round2_line_is_self_contained_and_same_config():
    """Round 2 (VERDICT r1 'Next round' 1-2): the one JSON line carries the reference's own precision, BASELINE configs 2 and 3
    and the UNMODIFIED reference timed on the GPU box; the reference arm runs the GPU arm's own config for >= 2 s and agrees
    with the same box's cpu_baseline.port."""
    d, ref = _load("r2_bench_c_20steps.json"), _load("r2_bench_b_reference_arm.json")
    r = d["roofline"]
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and abs(r["dram_frac"] - r["traffic"] / (r["kernel_ms"] * 1e-3) / 1e9 / r["peak"]) < 1e-9
    assert "NOT measured in this run" in r["traffic_source"] and 0.7 < r["dram_frac"] < r["frac"] <= 1.1
    f = d["fp64"]
    assert f["dtype"] == "f64" and f["roofline"]["algorithmic_bytes_per_env_step"] == 793 and f["roofline"]["frac"] >= 0.90
    assert abs(f["value"] - (1 << 20) / (f["ms_per_step"] * 1e-3)) < 1e-6 * f["value"]
    c = d["configs"]
    assert c["config3"]["env"] == "docking-v0" and c["config3"]["envs_per_gpu"] == 1 << 20 and c["config3"]["roofline"]["frac"] > 0.9
    assert c["config2"]["envs"] == 4096 and c["config2"]["f32"]["per_launch_us_per_step"] <= 6.0 and c["config2"]["f32"]["step_n_us_per_step"] < 2.0
    assert "PyTorch extension" in c["config2"]["binding"]
    b = d["cpu_baseline"]
    assert b["kind"] == "reference" and b["cores"] >= 1 and 100 < b["per_core"] < 5000 and abs(b["value"] - b["per_core"] * b["cores"]) < 1e-6 * b["value"]
    assert b["port"]["kind"] == "port" and "1048576 envs" in b["port"]["sample"] and b["config1"]["steps_per_s"] > 100
    assert abs(b["config1"]["final_pos"][0] + 0.2) < 1e-6  # run_sim_PID.py's scenario really converged to pos_des
    # the reference arm: same config dict as the GPU arm (the driver's same_config), >= 2 s timed, within 15 % of the port leg
    assert ref["impl"] == "reference" and ref["config"] == dict(d["config"], parallelism=ref["config"]["parallelism"])
    assert ref["config"]["envs_per_gpu"] == 1 << 20 and ref["timed_seconds"] >= 2.0 and ref["timed_steps"] >= ref["steps"]
    assert abs(ref["value"] - b["port"]["value"]) < 0.15 * b["port"]["value"]
    k = d["callers"]
    assert k["update_epoch_ms"] <= 15.0 and k["fused_step_kernel"]["gpu_launches_per_rollout"] == 18 and k["gpu_launches_per_rollout"] == 67


def test_round2_final_line_callers():
    """The round's last driver-style run (after the forward kernel's fix): the callers' numbers quoted in DESIGN.md section 9."""
    d = _load("r2_bench_g_20steps.json")
    k = d["callers"]
    assert k["rollout_env_steps_per_s"] >= 7.9e9 and k["update_epoch_ms"] <= 13.0 and k["gpu_launches_per_rollout"] == 67
    assert d["roofline"]["frac"] >= 0.99 and d["fp64"]["roofline"]["frac"] >= 0.95 and d["clocks"]["reasons"] == []
    ref = _load("r2_bench_g_reference_arm.json")
    assert ref["impl"] == "reference" and ref["timed_seconds"] >= 2.0 and abs(ref["value"] - d["cpu_baseline"]["port"]["value"]) < 0.15 * ref["value"]
