"""CPU: the bench contract.  The reference arm runs here (it is the CPU leg), so its JSON line is checked live;
the GPU arm's line is checked on the committed evidence of the round's last GPU job (profiles/)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config"}


def test_reference_arm_prints_one_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert BASE_KEYS <= set(d) and d["impl"] == "reference" and d["unit"] == "GB/s" and d["higher_is_better"] is True
    assert d["metric"] == json.load(open(os.path.join(ROOT, "BASELINE.json")))["metric"].split(" (")[0]
    assert d["vs_baseline"] is None and d["dtype"] == "f32" and d["data"] == "synthetic" and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert 0.05 < d["value"] < 50                                           # one host core, per-element iterators


def test_reference_arm_is_silent_on_other_ranks():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "2"],
                         capture_output=True, text=True, timeout=120, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


@__import__("pytest").mark.parametrize("rnd", ["r01", "r02"])
def test_committed_gpu_line_meets_the_contract(rnd):
    d = json.load(open(os.path.join(ROOT, "profiles", f"{rnd}_bench_final.json")))
    ref = json.load(open(os.path.join(ROOT, "profiles", f"{rnd}_bench_reference_arm.json")))
    assert BASE_KEYS <= set(d) and "impl" not in d or d.get("impl") == "ours"
    assert d["metric"] == ref["metric"] and d["unit"] == ref["unit"] and d["config"]["workload"] == ref["config"]["workload"]
    assert d["n_gpus"] == 1 and d["warmup"] >= 3 and d["gpu_launches"] == 2 * d["steps"] and d["scaling"] == "weak"
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-3 and r["traffic"] > 0
    assert r["algorithmic_bytes_per_launch"] == 3 * 8192 * 8192 * 4
    assert abs(r["achieved"] - r["algorithmic_bytes_per_launch"] / (r["avg_launch_ms"] * 1e-3) / 1e9) / r["achieved"] < 0.01
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == 1 and cb["unit"] == "GB/s" and cb["sample"]
    e = d["e2e"]
    assert e["unit"] == "GB/s" and e["h2d_bytes_per_step"] == 2 * 8192 * 8192 * 4 + 8192 * 4 and e["d2h_bytes_per_step"] == 8192 * 8192 * 4
    assert e["value"] < d["value"]                                          # host link inside the timed region
    c = d["clocks"]
    assert c["sm_mhz"] > 0.9 * c["sm_max_mhz"] and not set(c["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert d["value"] * d["ms_per_step"] * 1e-3 * 1e9 / d["config"]["bytes_per_step_per_gpu"] == __import__("pytest").approx(1.0, rel=1e-3)
    if rnd == "r02":                                                         # round 2: the reference arm times the whole workload
        assert ref["impl"] == "reference" and ref["ms_per_step"] > 500 and ref["config"]["sample"] == "the full workload per step"
        x = d["extras"]
        assert x["heat3d_2048_f32"]["steps"] == 100 and x["heat3d_2048_f32"]["subcube_vs_oracle"] and len(x["heat3d_2048_f32"]["field_hash"]) == 16
        assert x["reduce_sum_1e9_f32"]["result_ok"] and x["reduce_argmax_1e9_f32"]["result_ok"]
        e2 = d["e2e"]
        assert e2["naive"]["same_result"] and 0.85 < e2["frac_of_host_link_ceiling"] < 1.1


def test_committed_multi_gpu_lines_agree_with_each_other():
    """The 2 / 4 / 8-GPU lines of round 2 (both transports): the stencil's field hash is the single GPU's, every
    self-check holds, the in-bench agreement checks passed on every rank."""
    one = json.load(open(os.path.join(ROOT, "profiles", "r02_bench_final.json")))["extras"]["heat3d_2048_f32"]["field_hash"]
    for name in ("n2_p2p", "n2_nccl", "n4_p2p", "n8_p2p", "n8_nccl"):
        d = json.load(open(os.path.join(ROOT, "profiles", f"r02_bench_{name}.json")))
        n = int(name[1])
        x = d["extras"]
        assert d["n_gpus"] == n and x["heat3d_2048_f32"]["field_hash"] == one and x["heat3d_2048_f32"]["subcube_vs_oracle"]
        assert x["reduce_sum_1e9_f32"]["result_ok"] and x["reduce_argmax_1e9_f32"]["result_ok"] and x["sharded_permute_16384_f64"]["result_ok"]
        if "multi_gpu_parity" in x:
            assert x["multi_gpu_parity"]["ok"] and all(x["multi_gpu_parity"]["ranks_ok"]) and x["multi_gpu_parity"]["checks"] >= 100
