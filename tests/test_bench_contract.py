"""The reference arm of bench.py (`--impl reference`) on the CPU: one JSON line with the keys the driver reads, and the SAME
`config` object the GPU arm prints for the same arguments (the driver compares the two arms on it)."""
import importlib.util
import json
import os
import subprocess
import sys
import types

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench_module():
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "liboracle.so")), reason="oracle not built")
def test_reference_arm_line_and_config():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--cpu-sample-mib", "8"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert key in d, key
    assert d["impl"] == "reference" and d["gpu_launches"] == 0 and d["vs_baseline"] is None and d["value"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    # the config is a function of the arguments alone: what the GPU arm prints for the same command line
    args = types.SimpleNamespace(size_mib=1024, chunk=65536, corpus="mixed")
    assert d["config"] == _bench_module().workload_config(args, 1)
    assert "workload" in d["config"] and "model" not in d["config"]
