"""bench.py --impl reference: the JSON line of the reference arm (a real Gibbs iteration of the CPU restatement per step) carries the
keys the measurement contract names, on the same workload dict as the GPU arm, for both synthetic masks.  Tiny size: a few seconds."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("mask", ["band", "galplane"])
def test_reference_arm_json_contract(mask):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--nside", "16", "--lmax", "32",
                          "--steps", "2", "--warmup", "1", "--mask", mask], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "gibbs_iters_per_s" and line["unit"] == "it/s"
    assert line["n_gpus"] == 1 and line["steps"] == 2 and line["warmup"] == 1
    assert line["higher_is_better"] is True and line["scaling"] == "weak" and line["vs_baseline"] is None
    assert line["dtype"] == "f64" and line["data"] == "synthetic"
    cfg = line["config"]
    assert "workload" in cfg and "PNCP" in cfg["workload"] and cfg["nside"] == 16 and cfg["lmax"] == 32
    assert cfg["mask"].startswith(mask + ":")
    assert line["value"] > 0 and abs(line["value"] * line["ms_per_step"] - 1e3) < 1e-6 * 1e3
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and "Gibbs iteration" in cb["sample"]
    e2e = line["e2e"]
    assert e2e["value"] == line["value"] and e2e["unit"] == "it/s" and e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0
    assert len(line["pcg_iterations_per_step"]) == 2 and all(n > 0 for n in line["pcg_iterations_per_step"])
