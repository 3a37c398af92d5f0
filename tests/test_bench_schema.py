"""The committed bench lines carry every key of the driver contract (checked on the CPU box, no GPU needed)."""
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _line(name):
    path = os.path.join(ROOT, "profiles", name)
    if not os.path.exists(path):
        pytest.skip(f"{name} not recorded yet")
    return json.loads(open(path).read().strip().splitlines()[-1])


def test_b200_arm_line():
    d = _line("r01_bench_n1_final.json")
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "roofline", "e2e", "gpu_launches", "clocks"):
        assert key in d, key
    assert d["vs_baseline"] is None and d["data"] == "synthetic" and d["scaling"] == "weak" and d["dtype"] == "bf16"
    assert "workload" in d["config"] and "model" not in d["config"] and d["config"]["l2"]
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert r["traffic"] is None or r["traffic"] > 0
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] < d["value"]
    assert d["gpu_launches"] >= d["steps"]  # at least one of our kernels per step
    assert d["clocks"]["sm_max_mhz"] and not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown",
                                                                            "sw_thermal_slowdown"}
    assert abs(d["value"] - d["n_gpus"] * 8 / (d["ms_per_step"] * 1e-3)) / d["value"] < 1e-6


def test_cpu_baseline_and_reference_arm_lines():
    d = _line("r01_bench_n1.json")
    c = d["cpu_baseline"]
    assert c["kind"] == "reference" and c["cores"] >= 1 and c["value"] > 0 and c["sample"]
    r = _line("r01_bench_reference_arm.json")
    assert r["impl"] == "reference" and r["metric"] == d["metric"] and r["unit"] == d["unit"]
    assert r["e2e"]["h2d_bytes_per_step"] == 0 and r["e2e"]["d2h_bytes_per_step"] == 0
    assert r["cpu_baseline"]["value"] == r["value"]


def test_round2_lines():
    """Round 2: the N=1 line also carries the reference's function on the same GPU, the config-3 training block and the
    copy ceiling of the e2e leg; the N=2 / N=8 lines (torchrun) hold the same training block with the NCCL all-reduce."""
    d = _line("r02_bench_n1.json")
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks",
                "gpu_reference", "train"):
        assert key in d, key
    assert d["n_gpus"] == 1 and d["roofline"]["kernel"] == "msda_bwd_mma_kernel"
    assert abs(d["value"] - 8 / (d["ms_per_step"] * 1e-3)) / d["value"] < 1e-6
    g = d["gpu_reference"]
    assert g["fp32_ms_per_step"] > d["ms_per_step"] and g["autocast_bf16_ms_per_step"] > d["ms_per_step"]
    c = d["e2e"]["copy_ceiling"]
    assert 0 < c["fraction_reached"] <= 1.1 and c["ms_per_step"] > 0
    t1 = d["train"]
    assert t1["n_gpus"] == 1 and t1["allreduce_bytes"] == 0 and t1["stock_hf"]["images_per_s"] > 0
    r = _line("r02_bench_reference_arm.json")
    assert r["impl"] == "reference" and r["metric"] == d["metric"] and r["config"]["workload"] == d["config"]["workload"]
    for n in (2, 4, 8):
        m = _line(f"r02_bench_n{n}.json")
        t = m["train"]
        assert m["n_gpus"] == n and t["n_gpus"] == n and t["backend"] == "nccl"
        assert t["allreduce_bytes"] == 4 * t["params"]  # one fp32 gradient all-reduce per optimiser step
        # weak scaling of the training step: >= 7/8 of linear against the N=1 line (north star: >= 7x at N=8)
        assert t["images_per_s"] >= 0.875 * n * t1["images_per_s"]
