"""bench.py's JSON line carries every key of the bench contract (assembled by a GPU-free function)."""
import argparse
import json
import os
import subprocess
import sys

import numpy as np
import scipy.sparse as spa

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_report_has_contract_keys():
    import bench
    args = argparse.Namespace(steps=5, warmup=3)
    tm = dict(tile_nodes=4, threads=416, tiles=148, smem_bytes=183168, launches=33, stream_bytes=4 * 10**12, h2d_bytes=1, d2h_bytes=2)
    inst0 = (spa.identity(500, format="csc"), None, spa.random(1050, 500, density=0.7, format="csc", random_state=0), None, None, None)
    out = bench.build_report(args, 1, "w", 800, 800.0, 1.0e6, 1000000, 2.7, 2.7, 2.8, 2.9, tm, tm, np.array([25, 50]),
                             np.array([1, 1]), 1063.0, inst0, {"sm_mhz": 1965.0, "sm_max_mhz": 1965.0, "reasons": []}, 20.0, 1)
    json.dumps(out)
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline"):
        assert k in out, k
    assert out["dtype"] == "f64" and out["vs_baseline"] is None and "workload" in out["config"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in out["roofline"], k
    for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert k in out["e2e"], k
    assert abs(out["roofline"]["frac"] - out["roofline"]["achieved"] / out["roofline"]["peak"]) < 1e-12
    assert out["gpu_launches"] == 5 * 33


def test_reference_arm_runs_on_cpu():
    """--impl reference needs no GPU: it times the CPU oracle and prints the same JSON shape."""
    env = dict(os.environ, BENCH_REFERENCE_MAX_INSTANCES="1")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    line = json.loads(p.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["cpu_baseline"]["kind"] == "port" and line["value"] > 0
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["metric"] == "QP-relaxations/sec"
