"""CPU tests (no GPU) of bench.py's contract: the reference arm's JSON line, the rank rule under torchrun, and that the GPU arm
fails loudly — never falls back to the CPU — when there is no device."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(*flags, env=None, timeout=300):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *flags], capture_output=True, text=True, timeout=timeout, env=e, cwd=ROOT)


def test_reference_arm_prints_one_contract_line(model_dir):
    r = run_bench("--impl", "reference", "--model", "micro", "--steps", "1", "--warmup", "0", "--ref-budget-s", "20",
                  env={"NOBS_BENCH_MODEL_DIR": model_dir})
    assert r.returncode == 0, r.stderr[-500:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "audio-seconds/sec" and d["unit"] == "audio-s/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["n_gpus"] == 1 and d["gpu_launches"] == 0 and d["vs_baseline"] is None
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_runs_on_rank_0_only(model_dir):
    """Under torchrun every rank is started with --impl reference: rank 0 alone times the CPU path, the others exit 0 silently."""
    r = run_bench("--impl", "reference", "--model", "micro", "--steps", "1", "--warmup", "0", "--gpus", "2",
                  env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1", "NOBS_BENCH_MODEL_DIR": model_dir})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_gpu_arm_fails_loudly_without_a_device(model_dir):
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a GPU is present")
    r = run_bench("--model", "micro", "--windows", "1", "--steps", "1", "--warmup", "0", "--no-cpu-baseline", "--latency-clips", "0", "--beam-clips", "0",
                  env={"NOBS_BENCH_MODEL_DIR": model_dir})
    assert r.returncode != 0
    assert not [ln for ln in r.stdout.splitlines() if ln.startswith("{")]   # no bench line, and in particular no CPU-computed one
