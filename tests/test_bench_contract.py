"""bench.py's reference arm runs on the CPU (it times the oracle port, the only thing besides tests/ and smoke() that may
execute oracle/): check the JSON line contract on a tiny workload.  The CUDA arm needs a B200 and is covered by the
driver's own run."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra, env=None):
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--files", "16", "--data-len", "20000",
           "--steps", "2", "--warmup", "1", "--ref-files-per-thread", "1"] + extra
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    return out.stdout


def test_reference_arm_json_line():
    lines = [l for l in _run([]).splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "pcm_decode_mix_gsamples_per_s" and d["unit"] == "Gsamples/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 2 and d["value"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": "Gsamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "aiff::parse" in cb["sample"]
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["gpu_launches"] == 0


def test_reference_arm_only_rank0_prints_under_torchrun():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    assert _run(["--gpus", "2"], env=env).strip() == ""
