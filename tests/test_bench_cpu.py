"""bench.py contract checks that need no GPU: the reference arm prints one JSON line with the agreed keys,
and the synthetic population generator produces boards with the reference recipe's shape."""
import json
import os
import subprocess
import sys

import numpy as np

from conftest import ROOT


def test_reference_arm_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "1", "--generations", "0", "--workload", "c2"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["unit"] == "evals/s" and line["value"] > 0
    have_ref = os.path.isfile("/root/reference/GA.py") or os.path.isfile(os.path.join(ROOT, "oracle", "_ref", "GA.py"))
    assert line["cpu_baseline"]["kind"] == ("reference" if have_ref else "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["cpu_baseline_c_openmp"]["kind"] == "port"
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["value"] == line["value"]
    assert "workload" in line["config"]


def test_reference_arm_other_ranks_stay_silent():
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, timeout=120, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_synthetic_population_follows_the_reference_recipe():
    sys.path.insert(0, ROOT)
    import bench
    boards, rowlen = bench.synth_population(6, 60, 1024, 7)
    assert boards.shape == (1024, 6, 64) and rowlen.shape == (1024, 6)
    n_clone = round(1024 * bench.CLONE_RATE)
    assert (rowlen[:n_clone] == 60).all() and (boards[:n_clone] == boards[0]).all()       # exact copies first
    assert (rowlen[n_clone:] == 63).all()                                               # max(1, round(60 * .05)) = 3 gaps
    for p in (n_clone, 700, 1023):
        for r in range(6):
            row = boards[p, r, :63]
            assert int((row == 5).sum()) - int((boards[0, r, :60] == 5).sum()) == 3      # three gaps more than the clone
            assert row[row != 5].tolist() == boards[0, r, :60][boards[0, r, :60] != 5].tolist()
    assert (boards[:, :, 63:] == 5).all() and set(np.unique(boards)) <= {1, 2, 3, 4, 5}


def test_both_arms_report_the_same_config():
    """`config` is built by one function for both arms (the driver compares them)."""
    sys.path.insert(0, ROOT)
    import bench
    a = bench.workload_config("c3")
    assert a == bench.workload_config("c3") and a["population_per_gpu"] == 65536 and a["agent_window"] == [6, 30]
