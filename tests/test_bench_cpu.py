"""CPU-only checks of bench.py's reference arm (the leg the driver runs beside our arm): it runs without a GPU,
prints ONE JSON line with the contract's keys, times the reference's own binary when it is staged (else the oracle
port), and its same-workload record steps the arm's own pile shape with the oracle's grid prefilter."""
import json
import os
import subprocess
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*extra):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1",
                          "--steps", "1", "--warmup", "1", *extra], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, out.stdout[-2000:]
    return json.loads(lines[0])


def test_reference_arm_line_has_the_contract_keys():
    d = _run("--no-subrecords")
    assert d["impl"] == "reference" and d["metric"] == "body-steps/s" and d["unit"] == "body-steps/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 1
    assert d["value"] > 0 and d["ms_per_step"] > 0
    assert d["config"]["workload"] == "cube_pile_1M_100x100x100" and d["config"]["same_config"] is False
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] == 1 and cb["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "body-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "same_workload_grid_port" not in d        # --no-subrecords


def test_other_ranks_of_a_torchrun_launch_print_nothing():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                          "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=120, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_same_workload_record_small_pile():
    sys.path.insert(0, ROOT)
    import bench
    args = types.SimpleNamespace(bodies=1000, side=10, settle=12)
    r = bench.cpu_same_workload(args, timed_steps=2)
    assert r["bodies"] == 1000 and r["shape"] == "10x10x10" and r["steps"] == 2 and r["settle_steps"] == 12
    assert r["value"] > 0 and r["cores"] >= 1 and r["unit"] == "body-steps/s"
