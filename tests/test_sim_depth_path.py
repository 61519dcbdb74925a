"""The CUDA sources of the depth path, executed by the fiber emulator (tests/hostsim/cusim.h) through the
same C ABI as the real library, against the oracle.  Catches logic errors without a GPU; the `-m gpu`
tests repeat these cases (and larger ones) on the real library."""
import os
import subprocess
import sys

import numpy as np
import pytest

from common import make_case, run_depth_case

CASES = [
    # L chosen so that L mod 20 covers the three GC tail branches (0, 1, >=2)
    dict(L=400_000, seed=1),
    dict(L=400_001, seed=2),
    dict(L=400_019, seed=3),
    dict(L=600_007, seed=4, kw=dict(gcadjust=False)),
    dict(L=600_007, seed=5, kw=dict(cap=-1.0)),
    dict(L=600_007, seed=6, kw=dict(trans="MED")),
    dict(L=800_003, seed=7, kw=dict(m=51)),
    dict(L=1_500_003, seed=8, kw=dict(m=501)),
    dict(L=700_003, seed=9, stress=True),
    dict(L=500_003, seed=10, kw=dict(merge=False), stress=True),
    dict(L=500_003, seed=39, stress=True), dict(L=500_003, seed=51, stress=True),
    dict(L=600_007, seed=6, kw=dict(trans="MED", threshold=5.0)),   # Lmax = (4*5)^2 = 400 > 256: the large-footprint scan instantiation
    # the undocumented knobs (rsi.cpp:2020-2026): -e (RSI factor), -reflen (neighbourhood length), -maxchkbp (sub-sampling of long neighbourhoods)
    dict(L=700_003, seed=9, kw=dict(epsilon=0.5), stress=True), dict(L=500_003, seed=39, kw=dict(chklen=1.5), stress=True),
    dict(L=700_003, seed=9, kw=dict(maxchkbp=500), stress=True), dict(L=600_007, seed=5, kw=dict(epsilon=3.0, chklen=4.0, maxchkbp=2000)),
    dict(L=600_007, seed=6, kw=dict(trans="ALL")), dict(L=700_003, seed=9, kw=dict(trans="ALL"), stress=True),   # a final test fails: the speculative per-call results are redone in order
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: f"L{c['L']}-s{c['seed']}-{'-'.join(f'{k}{v}' for k, v in c.get('kw', {}).items()) or 'default'}")
def test_depth_path_matches_oracle(case, sim_lib, oracle):
    fa, d, _ = make_case(case["L"], case["seed"], stress=case.get("stress", False))
    calls, launches = run_depth_case(sim_lib, oracle, fa, d, level0_modes=(2, 1, 0), **case.get("kw", {}))
    assert launches > 0


@pytest.mark.parametrize("order", ["1", "2"])
def test_schedule_order_independent(order, sim_lib):
    """a missing __syncthreads shows up as a difference between fiber resume orders"""
    code = (
        "import sys; sys.path.insert(0, 'tests'); sys.path.insert(0, '.')\n"
        "from bind import Lib; from common import make_case, run_depth_case\n"
        "fa, d, _ = make_case(500_003, 12, stress=True)\n"
        f"run_depth_case({sim_lib!r}, Lib('oracle'), fa, d, level0_modes=(2, 1, 0))\n"
        "print('ok')\n")
    env = dict(os.environ, CUSIM_ORDER=order)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", code], cwd=root, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout + r.stderr


def test_low_depth_gives_no_calls(sim_lib, oracle):
    fa, d, _ = make_case(300_000, 13, mean=3.0)
    calls, _ = run_depth_case(sim_lib, oracle, fa, d, check_bins=False)
    assert calls == []


def test_high_depth_histogram_windows(sim_lib, oracle):
    """200x depth: the value-histogram windows of passes B and C no longer start at 0 (zeros of N stretches, values
    beyond the window and the capped range all take their side paths)"""
    fa, d, _ = make_case(400_003, 17, mean=200.0)
    run_depth_case(sim_lib, oracle, fa, d)
