"""Cheap guards: every Python entry point the driver runs must at least compile."""
import os
import py_compile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("rel", ["bench.py", "__graft_entry__.py", "tools/profile_shia.py", "tools/bench_path_a.py",
                                 "tools/bench_ensemble.py", "tools/e2e_probe.py", "tools/ncu_lines.py", "tools/ncu_raw.py",
                                 "tools/ncu_launches.py"])
def test_script_compiles(rel):
    py_compile.compile(os.path.join(ROOT, rel), doraise=True)


def test_bench_reference_arm_runs_small(tmp_path):
    """--impl reference on the small workload prints one JSON line with the contract's keys (CPU only)."""
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c2-small",
                          "--steps", "1", "--warmup", "0", "--ref-sample-steps", "4"], capture_output=True, text=True,
                         timeout=600, env={**os.environ, "CUDA_VISIBLE_DEVICES": ""})
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "config",
              "cpu_baseline", "e2e"):
        assert k in line, k
    from oracle import ref
    assert line["impl"] == "reference" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == ("reference" if ref.available() else "port")
