"""bench.py's reference arm on the CPU (no GPU needed): the record must carry the driver's keys and be honest about
the size it ran (the arm printed `config.n = 1000000` while timing a 100 k sample)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_record_says_what_it_ran():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--cpu-sample", "6000", "--levels", "2"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    rec = json.loads(out.stdout.strip().splitlines()[-1])
    assert rec["impl"] == "reference" and rec["higher_is_better"] is True and rec["unit"] == "Gaussians/s"
    assert rec["value"] > 0 and rec["steps"] == 1 and rec["n_gpus"] == 1 and rec["vs_baseline"] is None
    cfg = rec["config"]
    assert cfg["n"] == 6000 and cfg["sample_of_n"] == 3_000_000 and cfg["same_config"] is False      # the size it ran, and the size it stands for
    assert "6000-Gaussian" in cfg["workload"] and "configs[2]" in cfg["workload"]
    cb = rec["cpu_baseline"]
    assert cb["n"] == 6000 and cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == rec["value"]
    assert rec["e2e"] == {"value": rec["value"], "unit": "Gaussians/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert rec["gpu_launches"] == 0


def test_reference_arm_other_ranks_print_nothing():
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                          "--warmup", "0", "--cpu-sample", "2000"], capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""
