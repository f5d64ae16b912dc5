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
