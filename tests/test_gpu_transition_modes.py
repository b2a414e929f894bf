"""-m gpu: the persistent TMA-pipelined form of the split transition (transition_stream_kernel, selected with CBS_TR_MODE=1: the
switch is read once per process, hence the subprocess) replays reference traces and the K-step comparison like the default kernel."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_stream_transition_kernel_replays_reference_traces():
    env = dict(os.environ, CBS_TR_MODE="1")
    res = subprocess.run([sys.executable, "-m", "pytest", "tests/test_gpu_golden.py", "tests/test_gpu_ksteps.py", "-m", "gpu", "-x", "-q",
                          "-k", "default-g20_control or default-g32_control or default-p8_control_win or default-s16 or ksteps"],
                         cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-2000:]
    assert "passed" in res.stdout
