"""bench.py contract pieces that need no GPU: the algorithmic-byte formula of SURVEY 8d and the reference arm's JSON line."""
import importlib.util
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
bench = importlib.util.module_from_spec(spec)
spec.loader.exec_module(bench)


def test_algorithmic_bytes_match_survey_8d():
    """C3: 140 B per extended cell + 48 B per CPML-slab cell and axis = 6.755 GB per time step; per kernel
    60 N + 24 N_slab (velocity) and 80 N + 24 N_slab (stress).  2-D acoustic 48 B, 3-D acoustic 64 B, 2-D elastic 80 B per cell."""
    n = [338, 338, 338]
    bv, bs = bench.algorithmic_bytes_per_step(3, True, n, [2, 2, 2])
    N, Nslab = 338.0 ** 3, 3 * 82 * 338.0 ** 2
    assert bv == 60 * N + 24 * Nslab and bs == 80 * N + 24 * Nslab
    assert abs((bv + bs) / 1e9 - 6.755) < 2e-3
    for nd, el, per_cell in ((2, False, 48), (3, False, 64), (2, True, 80), (3, True, 140)):
        shape = [100, 100, 100][:nd] if nd == 3 else [100, 100]
        bv, bs = bench.algorithmic_bytes_per_step(nd, el, shape, [0] * nd)
        assert (bv + bs) / float(np.prod(shape)) == per_cell


def test_measured_peak_falls_back(tmp_path):
    assert bench.measured_peak() in ((bench.FALLBACK_HBM_GBS, "fallback"),) or bench.measured_peak()[1] == "measured"


def test_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU restatement with all host threads on a bounded sample): one JSON line with the base
    contract's keys, `impl: reference`, a `cpu_baseline` describing the run and a zero-copy `e2e`."""
    # torchrun exports OMP_NUM_THREADS=1 to every rank: the CPU arm must use the host's cores all the same (round-1 defect)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c3small", "--steps", "1",
                        "--warmup", "0", "--cpu-steps", "2"], capture_output=True, text=True, timeout=600,
                       env=dict(os.environ, OMP_NUM_THREADS="1"))
    assert r.returncode == 0, r.stderr[-2000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "Gcell-updates/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["cpu_baseline"]["cores"] == bench.host_threads()
    # the same `config` keys as our arm prints (the driver compares the two arms' configs)
    wl = {"label": "x"}
    assert set(d["config"]) == set(bench.bench_config(wl, 2, 1, [1, 1, 1], 1, 1, 1))
    assert d["config"]["time_steps_per_step"] == 200 and d["cpu_baseline"]["sample_time_steps"] == 2
