"""B200-native batched simulator for C-CyberBattleSim's continuous-env step path.

Importable as ``ccbs_b200`` (see ``ccbs_b200.py`` at the repo root; the directory name
``c-cyberbattlesim_b200`` is not a valid Python identifier)."""
from . import constants, config  # noqa: F401
from .config import EnvConfig  # noqa: F401
from .scenario import (ScenarioSpec, NodeSpec, VulnSpec, ResultSpec, ServiceSpec, ScenarioTables,  # noqa: F401
                       compile_scenarios, spec_from_model, synthetic_spec, synthetic_input_graph,
                       spec_from_input_graph, synthetic_vuln_pool, load_scenario_folder)

from .gae import GaeWeights, fold_gae  # noqa: F401,E402


def __getattr__(name):
    # torch-dependent modules are imported lazily so that `import ccbs_b200` stays cheap
    if name in ("BatchedCyberBattleEnv", "CbsError"):
        from . import batched_env
        return getattr(batched_env, name)
    if name == "ShardedHostEnv":
        from . import host_pipeline
        return host_pipeline.ShardedHostEnv
    if name in ("CyberBattleVecEnv", "RandomSwitchEnvB200"):
        from . import vec_env
        return getattr(vec_env, name)
    raise AttributeError(name)


__version__ = "0.1.0"
