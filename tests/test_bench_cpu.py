"""The parts of bench.py's contract that need no GPU: the reference arm (`--impl reference`, the CPU port of the reference's
algorithm on all host threads) prints ONE JSON line with the keys the driver reads, on the same `config` object as our arm of
the same command line, for every BASELINE.json preset; under torchrun only rank 0 runs it, with every host thread although
torchrun exports OMP_NUM_THREADS=1."""
import json
import os
import subprocess
import sys
import types

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CONTRACT = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
            "dtype", "data", "config", "e2e", "cpu_baseline", "impl")


def run(args, env=None, timeout=240):
    e = dict(os.environ)
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "OMP_NUM_THREADS"):
        e.pop(k, None)
    e.update(env or {})
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, env=e, timeout=timeout, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    return [l for l in p.stdout.splitlines() if l.startswith("{")]


@pytest.mark.parametrize("preset,batch", [(1, None), (3, 8), (4, 4)])
def test_reference_arm_prints_the_contract_line(preset, batch):
    args = ["--impl", "reference", "--config", str(preset), "--steps", "2", "--warmup", "1"] + (["--batch", str(batch)] if batch else [])
    lines = run(args)
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert all(k in d for k in CONTRACT), [k for k in CONTRACT if k not in d]
    assert d["impl"] == "reference" and d["unit"] == "images/s" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["steps"] == 2 and d["n_gpus"] == 1 and d["value"] > 0 and d["gpu_launches"] == 0
    assert abs(d["value"] - d["config"]["batch_per_gpu"] / (d["ms_per_step"] * 1e-3)) <= 1e-6 * d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] == d["value"] and cb["cores"] >= 1 and "steps" in cb["sample"]
    # the same `config` object as our arm of this command line (the driver compares the two)
    sys.path.insert(0, ROOT)
    import bench
    a = types.SimpleNamespace(config=preset, batch=batch, priors=None, gmax=None)
    bench.resolve(a, 1)
    assert d["config"] == bench.config_dict(a, d["config"]["num_priors"])
    assert d["config"]["workload"].startswith("configs[%d]" % preset) and d["scaling"] == bench.PRESETS[preset]["scaling"]


def test_reference_arm_under_torchrun_runs_on_rank_0_with_all_threads():
    env = {"WORLD_SIZE": "2", "OMP_NUM_THREADS": "1", "MASTER_ADDR": "127.0.0.1", "MASTER_PORT": "29511"}
    args = ["--impl", "reference", "--gpus", "2", "--steps", "2", "--warmup", "1", "--batch", "4"]
    assert run(args, dict(env, RANK="1", LOCAL_RANK="1")) == []                      # the other ranks exit 0 without work
    lines = run(args, dict(env, RANK="0", LOCAL_RANK="0"))
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["n_gpus"] == 2 and d["config"]["parallelism"] == "dp2" and d["config"]["global_batch"] == 8
    assert d["details"]["omp_threads"] == d["cpu_baseline"]["cores"] >= min(2, os.cpu_count() or 1)   # not torchrun's single thread
