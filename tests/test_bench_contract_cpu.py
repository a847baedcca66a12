"""CPU: the reference arm of bench.py (the oracle port timed on the host cores) prints exactly one JSON line with the
contract's keys; FLOP accounting matches SURVEY.md §8d."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup",
                        "0", "--cpu-sample-tokens", "128"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "impl", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == "denoising_steps_per_sec" and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["value"] > 0


def test_flop_accounting_matches_survey():
    sys.path.insert(0, ROOT)
    import bench
    assert abs(bench.fwd_flops(75600) / 1e15 - 6.523) < 0.01      # 6.523 PF per 720P forward
    assert abs(bench.fwd_flops(131040) / 1e15 - 17.257) < 0.01    # 17.257 PF per 1080P forward
