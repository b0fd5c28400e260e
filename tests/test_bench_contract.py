"""not gpu: the reference arm of bench.py runs on the host cores and prints the contract's JSON line.

`python bench.py --impl reference` times the unmodified reference (baseline/_ref when it was built in
this container, else the oracle port) on the C2 workload; no GPU is involved, so the line can be
checked here: same metric / unit / direction as the product arm, a `cpu_baseline` describing the run
and an `e2e` that repeats the value with zero transfer bytes."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, out.stdout[-2000:]
    d = json.loads(lines[0])
    assert d["impl"] == "reference"
    assert d["metric"].startswith("InterRDF frames/s") and d["unit"] == "frames/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["value"] > 0.0 and d["ms_per_step"] > 0.0
    assert d["config"]["workload"].startswith("C2") and d["config"]["atoms"] == 33334 and d["config"]["nbins"] == 150
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_product_arm_refuses_to_run_without_a_gpu():
    """No CPU fallback: without CUDA the product arm must fail loudly, not time something else."""
    import torch
    if torch.cuda.is_available():
        return
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0",
                          "--no-secondary", "--no-cpu"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode != 0
    assert not [ln for ln in out.stdout.splitlines() if ln.startswith("{") and '"value"' in ln]
