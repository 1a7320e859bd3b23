"""bench.py's output contract on the arm that runs without a GPU (`--impl reference`: the oracle port on the host cores):
exactly ONE JSON line on stdout (library chatter goes to stderr), the keys the driver reads, and silent non-zero ranks."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env=None, *flags):
    env = dict(os.environ)
    env.pop("RANK", None)
    env.update(extra_env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", *flags],
                          capture_output=True, text=True, env=env, timeout=600)


def test_reference_arm_prints_one_json_line():
    r = _run()
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "gnss_iq_synth_msamples_per_s" and d["unit"] == "Msamples/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0 and d["value"] > 0
    assert "workload" in d["config"] and "e1c_8prn_600s_cn34_orbital.yaml" in d["config"]["workload"] and d["scaling"] == "strong"
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["acq"]["metric"] == "pcps_acq_cells_per_s" and d["acq"]["value"] > 0 and d["gpu_launches"] == 0


def test_reference_arm_other_ranks_are_silent():
    r = _run({"RANK": "1", "WORLD_SIZE": "2"}, "--gpus", "2")
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_reference_arm_takes_another_workload():
    r = _run(None, "--workload", "e1c_8prn_60s_cn34_orbital.yaml")
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads(r.stdout.strip())
    assert "e1c_8prn_60s_cn34_orbital.yaml" in d["config"]["workload"]
