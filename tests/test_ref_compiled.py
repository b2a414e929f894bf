"""oracle/_ref (bytecode of the reference's step-path modules, compiled by oracle/build_ref.py from /root/reference) imports
without the source tree and steps the reference's own RandomSwitchEnv — what `bench.py --impl reference` times on the GPU box."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_PROBE = r"""
import sys
sys.path.insert(0, %r)
import numpy as np
import ccbs_b200 as cb
from ccbs_b200.gae import GaeWeights
from oracle import ref_bridge as rb
assert rb.reference_kind() == "compiled", rb.reference_kind()
g = cb.synthetic_input_graph(5, 8, pool=cb.synthetic_vuln_pool(1234, 60))
env = rb.make_unpatched_env(rb.reference_model_from_input_graph(g, seed=5), GaeWeights.random(0), cb.EnvConfig(), seed=3)
import cyberbattle._env.cyberbattle_env_switch as sw
assert sw.__file__.endswith(".pyc") and "/oracle/_ref/cyberbattle_ref.zip/" in sw.__file__, sw.__file__
obs, _ = env.reset()
assert obs["graph_embeddings"].shape == (192,)
n = 0
for _ in range(30):
    obs, r, done, trunc, info = env.step(env.action_space.sample())
    n += 1
    if done or trunc:
        env.reset()
print("OK", n)
"""


@pytest.mark.reference
@pytest.mark.skipif(not os.path.isdir("/root/reference/cyberbattle"), reason="reference tree not mounted")
def test_compiled_reference_steps_without_the_source_tree():
    from oracle import build_ref
    import zipfile
    archive = build_ref.build(verbose=False)
    names = set(zipfile.ZipFile(archive).namelist())
    assert names == {m[:-3] + ".pyc" for m in build_ref.MODULES} | {"cyberbattle/utils/__init__.pyc"}, \
        "bytecode only: no reference source may be staged"
    env = dict(os.environ, CBS_REFERENCE_ROOT=archive, PYTHONDONTWRITEBYTECODE="1")
    res = subprocess.run([sys.executable, "-c", _PROBE % ROOT], capture_output=True, text=True, env=env, timeout=600)
    assert res.returncode == 0 and "OK 30" in res.stdout, res.stderr[-2000:]
