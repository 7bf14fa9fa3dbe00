"""CPU checks of the driver-facing `bench.py` contract that need no GPU: the reference arm (`--impl reference`, the oracle
port timed on the host cores) prints ONE JSON line with the agreed keys, and under torchrun only rank 0 works and prints
while the other ranks exit 0.  (The product arm needs a CUDA device and must fail loudly without one.)"""
import json
import os
import subprocess
import sys

from helpers import ROOT

BENCH = os.path.join(ROOT, "bench.py")
SMALL = ["--steps", "1", "--warmup", "0", "--no-measured-configs", "--num-gaussians", "20000"]


def _json_lines(out):
    return [json.loads(l) for l in out.splitlines() if l.startswith("{")]


def _check_reference_line(d, n_gpus):
    assert d["impl"] == "reference" and d["n_gpus"] == n_gpus
    assert d["metric"].startswith("fwd+bwd Gaussians/sec") and d["unit"] == "Gaussians/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["steps"] == 1
    assert d["estimated"] is True and 0 < d["sampled_fraction"]["gaussians"] <= 1.0
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_prints_one_contract_line():
    r = subprocess.run([sys.executable, BENCH, "--impl", "reference", *SMALL], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = _json_lines(r.stdout)
    assert len(lines) == 1
    _check_reference_line(lines[0], 1)


def test_reference_arm_under_torchrun_only_rank0_prints():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29537", BENCH, "--impl", "reference", "--gpus", "2", *SMALL]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = _json_lines(r.stdout)
    assert len(lines) == 1, "exactly one rank prints the reference line"
    _check_reference_line(lines[0], 2)


def test_product_arm_refuses_to_run_without_cuda():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a CUDA device is present: the product arm would run")
    r = subprocess.run([sys.executable, BENCH, "--steps", "1", "--warmup", "3", "--num-gaussians", "1000"], capture_output=True,
                       text=True, timeout=600)
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)
    assert not _json_lines(r.stdout), "no bench line may be printed by a run that did no GPU work"
