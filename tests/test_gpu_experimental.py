"""A/B parity of the opt-in kernel switches that have NOT been promoted to defaults ("Experiment switches").

Skipped unless B2M_TEST_EXPERIMENTAL=1: these paths may be untested on hardware (they are off by default for that
reason), so they must not be able to turn the regular GPU suite red.  Each switch is read once per process, hence the
child processes.  Run on the GPU box:  B2M_TEST_EXPERIMENTAL=1 python -m pytest tests/test_gpu_experimental.py -m gpu
"""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("B2M_TEST_EXPERIMENTAL") != "1", reason="opt-in: B2M_TEST_EXPERIMENTAL=1")]

SWITCHES = [
    {"B2M_ATOMCONV": "1"},                                   # first-generation atom-conv kernels (per-lane gathers)
    {"B2M_ATOMCONV": "4"},                                   # third-generation forward + first-generation backward
    {"B2M_AC3_L1PF": "0"},                                   # no L1 prefetch of the next tile's C rows
    {"B2M_GEMM_PIPE": "0"},                                  # first row-GEMM kernel
    {"B2M_L2_PREFETCH": "2"},
    {"B2M_L2_PREFETCH": "0"},
    {"B2M_ATOMCONV": "1", "B2M_FWD_PREFETCH": "0"},          # generation-1 only switches
    {"B2M_ATOMCONV": "1", "B2M_FWD_THREADS": "512", "B2M_BWD_THREADS": "512"},
]


def _run(tmp_path, name, extra_env):
    out = tmp_path / f"{name}.npz"
    env = dict(os.environ, **extra_env)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "_run_case.py"), str(out)], env=env, cwd=ROOT,
                       capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr[-2000:]
    return np.load(out)


@pytest.fixture(scope="module")
def baseline(tmp_path_factory):
    return _run(tmp_path_factory.mktemp("exp"), "default", {})


@pytest.mark.parametrize("sw", SWITCHES, ids=lambda d: ",".join(f"{k}={v}" for k, v in d.items()))
def test_switch_matches_default(tmp_path, baseline, sw):
    got = _run(tmp_path, "case", sw)
    n = baseline["F"].shape[0]
    assert abs(float(got["E"]) - float(baseline["E"])) / n < 1e-7
    assert np.abs(got["F"] - baseline["F"]).max() < 1e-6
    assert np.abs(got["S"] - baseline["S"]).max() < 1e-6
