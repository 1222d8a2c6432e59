"""The profile tooling parses the committed ncu captures (bench.py reports roofline.traffic from its output)."""
import json
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_dram_per_step_tool_reproduces_the_committed_json(tmp_path):
    out = tmp_path / "dram.json"
    subprocess.run([sys.executable, os.path.join(REPO, "tools", "dram_per_step.py"),
                    os.path.join(REPO, "profiles", "r1_tc_dram_launches.csv"), str(out), "--r1"], check=True, capture_output=True)
    got = json.load(open(out))
    ref = json.load(open(os.path.join(REPO, "profiles", "r1_tc_dram_per_step.json")))
    assert got["launches"] == ref["launches"] == 42           # gather + parity + wgrad of E2-E8 / D8-D2, three passes each
    assert abs(got["dram_bytes_per_launch"] - ref["dram_bytes_per_launch"]) <= 1e-6 * ref["dram_bytes_per_launch"]
    assert 2e9 < got["dram_bytes_per_step"] < 4e9
    thin = [e for e in got["per_launch"] if "gemm_tn" in e["kernel"]]
    assert len(got["per_launch"]) == 48 and len(thin) == 2


def test_dram_per_step_tool_on_the_round2_capture(tmp_path):
    """Round-2 capture: several eager steps of every tc_ launch, current template order; the tool keeps the last complete
    step (STFT GEMM + 42 convolution launches; the thin layers run in their own kernels now)."""
    out = tmp_path / "dram2.json"
    subprocess.run([sys.executable, os.path.join(REPO, "tools", "dram_per_step.py"),
                    os.path.join(REPO, "profiles", "r2_tc_dram_launches.csv"), str(out)], check=True, capture_output=True)
    got = json.load(open(out))
    ref = json.load(open(os.path.join(REPO, "profiles", "r2_tc_dram_per_step.json")))
    assert got["launches"] == ref["launches"] == 42 and len(got["per_launch"]) == 43
    assert abs(got["dram_bytes_per_launch"] - ref["dram_bytes_per_launch"]) <= 1e-6 * ref["dram_bytes_per_launch"]
    assert 2e9 < got["dram_bytes_per_step"] < 4e9          # ~3.09 GB algorithmic: no re-reads from HBM
    assert sum("tc_wgrad" in e["kernel"] for e in got["per_launch"]) == 14


def test_launch_summary_tool_finds_one_step():
    res = subprocess.run([sys.executable, os.path.join(REPO, "tools", "launch_summary.py"),
                          os.path.join(REPO, "profiles", "r1q_launches_bench_b64.csv"), "--md", "--tc"], check=True,
                         capture_output=True, text=True).stdout
    assert "launches per step: 140" in res and "tc_wgrad_kernel" in res and "clip_adamw_kernel" in res


def test_bench_lines_of_the_round_are_well_formed():
    need = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
            "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline"}
    for n in (1, 2, 8):
        d = json.load(open(os.path.join(REPO, "profiles", "r1_bench_%dgpu.json" % n)))
        assert need <= set(d), (n, need - set(d))
        assert d["n_gpus"] == n and d["unit"] == "samples/s" and d["scaling"] == "weak" and d["dtype"] == "bf16"
        assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(d["e2e"]) and d["e2e"]["h2d_bytes_per_step"] > 0
        assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(d["roofline"])
        assert "workload" in d["config"] and d["gpu_launches"] > 0
    assert "cpu_baseline" in json.load(open(os.path.join(REPO, "profiles", "r1_bench_1gpu.json")))
    # the round-2 line adds the secondary (HBM-bound) rooflines, the large / deep split and the reference-kind CPU leg
    d = json.loads(open(os.path.join(REPO, "profiles", "r2i_bench_1gpu.json")).read().strip().splitlines()[-1])
    assert need <= set(d) and d["n_gpus"] == 1 and d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1
    r = d["roofline"]
    assert r["bound"] == "tensor" and 0.3 < r["frac"] < 1.0 and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert {"large_layers", "deep_layers", "secondary", "families", "traffic"} <= set(r) and r["traffic"] > 0
    assert len(r["secondary"]) == 5 and all(s["bound"] == "hbm" and 0 < s["frac"] < 1.05 for s in r["secondary"])
    assert d["gpu_launches"] == 99 * d["steps"]
