"""The C-ABI library loads without a GPU, exports every symbol include/cbsim.h declares, agrees with the ctypes
struct layouts, and fails loudly (no CPU fallback) when no CUDA device is present."""
import ctypes as ct
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from ccbs_b200 import lib as L
    L.build_library()
    return L.load_library()


def _declared_functions():
    text = open(os.path.join(ROOT, "include", "cbsim.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cbs_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported(lib):
    from ccbs_b200 import lib as L
    declared = _declared_functions()
    assert len(declared) >= 20
    assert sorted(L.SYMBOLS) == declared, "lib.SYMBOLS and include/cbsim.h are out of sync"
    for name in declared:
        assert hasattr(lib, name), f"libcbsim.so does not export {name}"


def test_struct_layouts_match(lib):
    from ccbs_b200 import lib as L
    out = (ct.c_int32 * 3)()
    assert lib.cbs_struct_sizes(out) == 0
    assert list(out) == [ct.sizeof(L.CbsConfig), ct.sizeof(L.CbsScenarioTables), ct.sizeof(L.CbsGaeTables)]
    assert lib.cbs_abi_version() == L.ABI_VERSION


def test_header_constants_match_python():
    import ccbs_b200.constants as C
    text = open(os.path.join(ROOT, "include", "cbsim.h")).read()
    defs = {k: int(v) for k, v in re.findall(r"#define\s+(CBS_[A-Z_]+)\s+(\d+)", text)}
    assert defs["CBS_ACTION_DIM"] == C.ACTION_DIM and defs["CBS_OBS_DIM"] == C.OBS_DIM + 2
    assert defs["CBS_NUM_REWARDS"] == len(C.REWARD_KEYS) and defs["CBS_NUM_PENALTIES"] == len(C.PENALTY_KEYS)
    assert defs["CBS_MAX_NODES"] == C.MAX_NODES and defs["CBS_VULN_EMB_DIM"] == C.VULN_EMB_DIM


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import ccbs_b200 as cb
    from ccbs_b200 import lib as L
    cfg = L.make_config(cb.EnvConfig(), num_envs=4)
    h = ct.c_void_p()
    rc = lib.cbs_create(ct.byref(cfg), ct.byref(h))
    assert rc == -5 and not h.value
    assert b"no CUDA device" in lib.cbs_last_error(None)
    with pytest.raises(cb.CbsError):
        cb.BatchedCyberBattleEnv([cb.synthetic_spec(0, 8)], cb.GaeWeights.random(0))


def test_rejects_bad_arguments(lib):
    import ccbs_b200 as cb
    from ccbs_b200 import lib as L
    cfg = L.make_config(cb.EnvConfig(), num_envs=4)
    cfg.abi_version = 99
    h = ct.c_void_p()
    assert lib.cbs_create(ct.byref(cfg), ct.byref(h)) == -1
    assert b"ABI" in lib.cbs_last_error(None)
    assert lib.cbs_create(None, ct.byref(h)) == -1
    with pytest.raises(ValueError):
        cb.EnvConfig(goal="conquer")
    # compressed:578-579: an unknown decode metric is a ValueError in the reference
    with pytest.raises(ValueError, match="Unsupported metric"):
        cb.EnvConfig(distance_metric="l3")
    cfg = L.make_config(cb.EnvConfig(distance_metric="inf"), num_envs=4)
    assert cfg.distance_metric == 3
    cfg.distance_metric = 7
    assert lib.cbs_create(ct.byref(cfg), ct.byref(h)) == -1
    assert b"metric" in lib.cbs_last_error(None)
