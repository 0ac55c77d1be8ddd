"""bench.py --impl reference (the CPU oracle port) prints one JSON line with the contract's keys; runs on CPU."""
import json
import pathlib
import subprocess
import sys

ROOT = pathlib.Path(__file__).resolve().parent.parent


def test_reference_arm_json_line():
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--nspins", "40", "--nsamples", "1e5", "--cpu-rows", "256", "--sweeps", "5"],
                         capture_output=True, text=True, timeout=300, check=True).stdout.strip().splitlines()
    line = json.loads(out[-1])
    assert line["impl"] == "reference" and line["higher_is_better"] is True
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "dtype", "data", "config",
                "cpu_baseline", "e2e", "gpu_launches"):
        assert key in line
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["value"] == line["value"] > 0
    assert "workload" in line["config"]
