"""The bench JSON contract (one line per run): keys the driver reads, checked on the committed lines
of the last GPU runs (profiles/r2_bench_n*.json: the default command at 1 / 2 / 4 / 8 GPUs) and on a
live `--impl reference` run (CPU only)."""
import json
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
BASE = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
        "vs_baseline", "dtype", "data", "config", "e2e"}
LINES = sorted((ROOT / "profiles").glob("r2_bench_n*.json"))


def _check_record(d, workload, full=True):
    if full:
        assert BASE <= d.keys(), BASE - d.keys()
        assert d["metric"] == "numerov_grid_steps_x_trial_energies_per_s" and d["unit"] == "steps/s"
        assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["dtype"] == "f64" and d["data"] == "synthetic"
        assert d["scaling"] == ("strong" if workload in ("c4", "c5") else "weak") and d["warmup"] >= 3
        assert d["config"]["workload"].startswith(workload) and "model" not in d["config"]
        c = d["clocks"]
        assert {"sm_mhz", "sm_max_mhz", "reasons"} <= c.keys()
        assert not ({"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(c["reasons"]))
    assert d["value"] > 0 and d["ms_per_step"] > 0
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= d["e2e"].keys()
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and 0 < d["e2e"]["value"] <= d["value"] * 1.001
    assert d["gpu_launches"] > 0
    if "roofline" in d:
        r = d["roofline"]
        assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= r.keys()
        assert r["unit"] == "TFLOP/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and 0.2 < r["frac"] < 0.8
        assert r["flop_per_step"] in (6, 8) and r["fp64_instr_per_step"] in (4, 5)
        if r["traffic"] is not None:  # a committed capture on the current kernel sources
            assert r["traffic"] >= 0.9 * r["algorithmic_bytes_per_launch"] or workload == "c5"
    if "cpu_baseline" in d:
        assert {"value", "unit", "cores", "kind", "sample"} <= d["cpu_baseline"].keys()
        assert d["cpu_baseline"]["kind"] == "port"
        for k in ("levels_bit_identical_to_gpu", "nodes_bit_identical_to_gpu"):
            if k in d["cpu_baseline"]:
                assert d["cpu_baseline"][k] is True
    if "nodes_bit_identical_to_oracle_full_size_sample" in d:
        assert d["nodes_bit_identical_to_oracle_full_size_sample"] is True


@pytest.mark.parametrize("path", LINES, ids=lambda p: p.name)
def test_committed_bench_lines(path):
    text = path.read_text().strip()
    assert text.count("\n") == 0, "exactly one JSON line"
    d = json.loads(text)
    _check_record(d, "c5")  # headline: the north star's largest energy sweep, energy-range sharded
    assert "nodes_bit_identical_to_oracle_full_size_sample" in d
    assert d["measured_peaks"]["fp64_tflops"] > 30
    subs = d["sub_records"]
    assert set(subs) == {"c2", "c3", "c4"}  # every other BASELINE config rides in the same line
    for name, rec in subs.items():
        _check_record(rec, name)
        _check_record(rec["accurate_mode"], name, full=False)
    _check_record(d["accurate_mode"], "c5", full=False)
    for name in ("c2", "c4"):
        assert subs[name]["time_to_all_levels_ms"] == subs[name]["ms_per_step"]
        assert subs[name]["cooley_mode"]["time_to_all_levels_ms"] > 0
    assert subs["c2"]["cooley_mode"]["max_rel_diff_vs_ksection"] < 1e-12
    assert subs["c2"]["levels_found"] == 17


def test_there_are_committed_lines():
    assert any(p.name == "r2_bench_n1.json" for p in LINES)


def test_reference_arm_line_live():
    """`bench.py --impl reference` runs without a GPU: same metric / config as the CUDA arm, one line."""
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, cwd=ROOT, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [x for x in r.stdout.splitlines() if x.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and BASE <= d.keys()
    assert d["e2e"]["h2d_bytes_per_step"] == 0 == d["e2e"]["d2h_bytes_per_step"] and d["e2e"]["value"] == d["value"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    sys.path.insert(0, str(ROOT))
    import bench

    # the headline of both arms is the north star's largest energy sweep (BASELINE.json configs[4])
    assert d["config"] == bench.workload_config("c5") and d["metric"] == bench.METRIC and d["scaling"] == "strong"
    committed = ROOT / "profiles" / "r2_bench_n1.json"
    if committed.exists():
        assert json.loads(committed.read_text())["config"] == d["config"]


def test_traffic_figures_belong_to_the_committed_kernel_sources():
    """profiles/traffic.json (ncu DRAM bytes per workload) carries the digest of the csrc it was captured
    on; bench.py reports a figure from other sources as stale (traffic: null).  The committed file must
    match the committed sources, and the committed headline line must carry its c5 figure."""
    sys.path.insert(0, str(ROOT))
    import bench

    t = json.loads((ROOT / "profiles" / "traffic.json").read_text())
    assert {"c2", "c3", "c4", "c5", "c2_dform", "c3_dform", "c4_dform", "c5_dform"} <= t.keys()
    assert {v["csrc_digest"] for v in t.values()} == {bench.csrc_digest()}
    line = json.loads((ROOT / "profiles" / "r2_bench_n1.json").read_text())
    assert line["roofline"]["traffic"] == t["c5"]["dram_bytes_per_launch"]
    # the constant-bank sweep's carried state stays in L2: HBM traffic within 10x of the algorithmic bytes
    assert line["roofline"]["traffic"] < 10 * line["roofline"]["algorithmic_bytes_per_launch"]


def test_cross_run_checksums_agree_across_gpu_counts():
    """Sum of all 2^24 node counts of c5 and xor of the level bits of c4: the same at 1, 2, 4 and 8 GPUs."""
    sums, xors = set(), set()
    for p in LINES:
        d = json.loads(p.read_text())
        sums.add(d["checksum"]["sum_of_node_counts_all_energies"])
        xors.add(d["sub_records"]["c4"]["checksum"]["xor_of_level_bits"])
    assert len(LINES) >= 4 and len(sums) == 1 and len(xors) == 1
