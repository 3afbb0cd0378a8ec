"""bench.py's reference arm end to end on the CPU (`--impl reference`: the unmodified reference's CLI from oracle/_ref
where present, else the oracle port) on a small sample, and the shape of the JSON line it prints -- the part of the
measurement contract that does not need a GPU."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _reference_line(extra, env=None):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--sample-reads", "1000"] + extra, capture_output=True, text=True,
                         timeout=600, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.split("\n") if ln.startswith("{")]
    assert len(lines) == 1                    # ONE JSON line
    return json.loads(lines[0])


def test_reference_arm_prints_the_contract_line():
    line = _reference_line(["--workload", "c1"])
    assert line["impl"] == "reference"
    assert line["metric"] == "k-mers/sec counted+filtered+graph-built" and line["unit"] == "k-mers/s"
    assert line["higher_is_better"] is True and line["vs_baseline"] is None and line["data"] == "synthetic"
    assert line["n_gpus"] == 1 and line["steps"] == 1 and line["warmup"] == 0
    assert line["value"] > 0 and line["ms_per_step"] > 0
    cfg = line["config"]
    assert cfg["workload"].startswith("C1") and cfg["k"] == 28 and cfg["filter"] == 3 and cfg["paired"] is True
    assert cfg["occurrences"] == 19000 * 2 * (100 - 28 + 2)
    base = line["cpu_baseline"]
    have_ref = os.path.exists(os.path.join(ROOT, "oracle", "_ref", "assemble.py"))
    assert base["kind"] == ("reference" if have_ref else "port") and base["cores"] == 1
    assert base["value"] == line["value"] and "1000 reads" in base["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "k-mers/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_under_torchrun_only_rank_zero_works():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=120, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""       # other ranks exit 0 without work
    line = _reference_line(["--workload", "c2", "--gpus", "2"], env=dict(os.environ, RANK="0", WORLD_SIZE="2", LOCAL_RANK="0"))
    assert line["n_gpus"] == 2 and line["config"]["sharding"] != "none" and line["config"]["paired"] is False


def test_both_arms_print_the_same_config():
    """The driver compares the two arms on `config`, `metric`, `unit`, `higher_is_better`: one function makes the
    config of both (bench.workload_config)."""
    sys.path.insert(0, ROOT)
    try:
        import bench
    finally:
        sys.path.remove(ROOT)
    import inspect
    src_ref, src_gpu = inspect.getsource(bench.run_reference_arm), inspect.getsource(bench.run_gpu_arm)
    assert "workload_config(args" in src_ref and "workload_config(args" in src_gpu
    assert '"metric": "k-mers/sec counted+filtered+graph-built"' in src_ref
    assert '"metric": "k-mers/sec counted+filtered+graph-built"' in src_gpu
