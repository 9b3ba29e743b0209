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
