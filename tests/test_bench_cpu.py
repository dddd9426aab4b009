"""bench.py contract checks that need no GPU: the reference arm's JSON line, and the product arm refusing to run on CPU."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, cwd=ROOT,
                          env=dict(os.environ, **(env or {})), timeout=600)


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    r = _run(["--impl", "reference", "--steps", "1", "--warmup", "1"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "cd_cloud_pairs_per_s" and d["unit"] == "cloud-pairs/s"
    assert d["higher_is_better"] is True and d["steps"] == 1 and d["warmup"] == 1 and d["value"] > 0
    assert d["config"]["workload"] == "allpairs_cd_1000x1000_clouds_2048pts"
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_quietly():
    r = _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"], env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_product_arm_refuses_to_run_without_cuda():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("needs a CPU-only box")
    r = _run(["--steps", "1", "--warmup", "0"])
    assert r.returncode != 0
    assert "no CPU fallback" in (r.stderr + r.stdout)
