"""The FLOP accounting bench.py reports against: visible (row, col) pairs under the kernel's mask rules."""
import importlib.util
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_visible_pairs_match_survey_figures():
    b = _bench()
    assert b.visible_pairs(4608, 4608, False, -1) == 4608 * 4608                       # config 2: 21 233 664
    N, W = 32768, 4096
    assert b.visible_pairs(N, N, True, W) == (W + 1) * (W + 2) // 2 + (N - W - 1) * (W + 1) == 125_859_840   # config 4
    w = b.WORKLOADS["flux"]
    assert abs(b.fwd_flops(w) - 260.919263232e9) < 1.0                                 # 260.92 GFLOP per launch
    import sys
    sys.path.insert(0, os.path.join(ROOT, "universal-metal-flash-attention_b200"))
    from umfa import ring
    assert ring.visible_pairs_causal(131072) == 8_590_000_128                          # config 5


def test_visible_pairs_brute_force():
    b = _bench()
    rng = np.random.default_rng(0)
    for _ in range(20):
        Sq, Skv = int(rng.integers(1, 60)), int(rng.integers(1, 60))
        causal = bool(rng.integers(0, 2))
        window = int(rng.integers(-1, 20))
        r, c = np.arange(Sq)[:, None], np.arange(Skv)[None, :]
        vis = np.ones((Sq, Skv), bool)
        if causal:
            vis &= c <= r
        if window >= 0:
            vis &= r <= c + window
        assert b.visible_pairs(Sq, Skv, causal, window) == int(vis.sum()), (Sq, Skv, causal, window)


def test_reference_arm_prints_the_contract_line():
    """bench.py --impl reference needs no GPU: it times the CPU oracle and prints one JSON line with the contract keys"""
    import json
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "small",
                        "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "TFLOP/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["gpu_launches"] == 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
