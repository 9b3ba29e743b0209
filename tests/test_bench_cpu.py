"""The CPU arm of bench.py (`--impl reference`) runs without a GPU and prints the
contract's JSON line."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "cora",
                          "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["value"] > 0 and line["steps"] == 2
    # the reference's own code when oracle/_ref is staged (or /root/reference is mounted), else the port
    sys.path.insert(0, ROOT)
    from oracle import ref_shim
    want_kind = "reference" if ref_shim.reference_root() is not None else "port"
    assert line["cpu_baseline"]["kind"] == want_kind and line["cpu_baseline"]["cores"] == 1
    split = line["cpu_baseline"]["split_seconds"]
    assert split["total"] > 0 and ("laplacian" in split or "laplacian_and_rescale" in split) and "recurrence" in split
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["value"] == line["value"]
    # same configuration object as the GPU arm prints: the full named shape, not a sample
    from efficient_gnn_b200 import synth
    cfg = line["config"]
    assert cfg["workload"] == "cora-shape" and cfg["n"] == synth.SHAPES["cora"].n and cfg["k"] == 3 and cfg["f"] == 1
    assert cfg["nnz"] == synth.SHAPES["cora"].nnz + cfg["n"] and line["scaling"] == "strong"


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "cora",
                          "--gpus", "2", "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=300,
                         cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


import pytest  # noqa: E402


@pytest.mark.gpu
def test_gpu_arm_prints_the_contract_line():
    """The GPU arm on a small shape: one JSON line with the contract's keys, the roofline and the
    end-to-end objects, and a config object identical to the reference arm's."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--workload", "pubmed", "--steps", "6",
                          "--warmup", "3"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert key in line, key
    assert line["n_gpus"] == 1 and line["steps"] == 6 and line["warmup"] >= 3 and line["value"] > 0
    assert line["scaling"] == "strong" and line["vs_baseline"] is None and line["dtype"] == "f32"
    r = line["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and r["peak"] > 0 and 0 < r["frac"] < 1.2
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    e = line["e2e"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] < line["value"]
    assert line["gpu_launches"] > 0
    ref = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "pubmed",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert ref.returncode == 0, ref.stderr[-2000:]
    assert json.loads(ref.stdout.strip().splitlines()[-1])["config"] == line["config"]
