"""constants.py (host) and csrc/cbs_types.h (device) must carry the same numbers."""
import os
import re

import ccbs_b200.constants as C

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HDR = open(os.path.join(ROOT, "c-cyberbattlesim_b200", "csrc", "cbs_types.h")).read()


def _enum(name):
    body = re.search(r"enum\s+" + name + r"\s*:\s*int\s*\{(.*?)\};", HDR, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    body = re.sub(r"//[^\n]*", "", body)
    out, nxt = {}, 0
    for item in body.split(","):
        item = item.strip()
        if not item:
            continue
        if "=" in item:
            k, v = [x.strip() for x in item.split("=")]
            nxt = int(v)
        else:
            k = item
        out[k] = nxt
        nxt += 1
    return out


def test_kinds_and_codes():
    k = _enum("Kind")
    for name in ("K_DOS", "K_DISCOVERY", "K_COLLECTION", "K_EXFILTRATION", "K_RECON", "K_EVASION", "K_PERSISTENCE",
                 "K_PRIVESC", "K_CREDACCESS", "K_LATERAL", "K_EXECUTION", "N_KINDS"):
        assert k[name] == getattr(C, name)
    c = _enum("Code")
    for name, v in c.items():
        assert getattr(C, name) == v


def test_masks_rewards_penalties():
    m = _enum("Mask")
    assert m["N_MASKS"] == C.N_MASKS == len(C.MASK_NAMES)
    for name, v in m.items():
        assert getattr(C, name) == v
    assert _enum("Reward")["N_REWARDS"] == len(C.REWARD_KEYS)
    assert _enum("Penalty")["N_PENALTIES"] == len(C.PENALTY_KEYS)
    assert _enum("Penalty")["P_DISTANCE"] == C.PENALTY_KEYS.index("distance_penalty")
    assert _enum("Penalty")["P_INVALID_ACTION"] == C.PENALTY_KEYS.index("invalid_action")
    assert _enum("Reward")["R_DOS"] == C.REWARD_KEYS.index("dos_coefficient")


def test_scalars_match_lib():
    from ccbs_b200 import lib as L
    s = _enum("Scalar")
    assert s["N_SCALARS"] == L.NUM_SCALARS
    names = [n for n in s if n.startswith("S_")]
    assert len(names) == L.NUM_SCALARS
    for name in names:
        assert s[name] == getattr(L, name), name
    # sector grouping the kernels rely on (transition.cuh): hot words 0-7, list lengths 8-15, episode constants 16-23
    assert s["S_EP_RETURN"] % 2 == 0 and s["S_EP_RETURN_HI"] == s["S_EP_RETURN"] + 1 < 8
    assert all(s[n] < 8 for n in ("S_FLAGS", "S_STEPCOUNT", "S_NUM_ITER", "S_TOTAL_STEPS", "S_OUTCOME", "S_SCST"))
    assert all(8 <= s[n] < 16 for n in ("S_N_DISC", "S_N_OWNED", "S_DISC_AMOUNT", "S_N_OWNED_RAW", "S_N_REIMAGED"))
    assert _enum("Accum")["N_ACCUM"] == L.NUM_ACCUM == len(L.ACCUM_NAMES)


def test_goals():
    g = _enum("Goal")
    for name, v in g.items():
        assert getattr(C, name) == v
    assert sorted(C.GOALS.values()) == list(range(6))
    assert C.obs_dim_for_goal("control") == 194 and C.obs_dim_for_goal("discovery_node") == 258


def test_dimensions():
    assert C.ACTION_DIM == 905 and C.NODE_FEAT_DIM == 1576 and C.OBS_DIM == 192
    assert re.search(r"ACTION_DIM = (\d+)", HDR).group(1) == "905"
    assert [C.onehot_index(1, k) for k in (C.K_CREDACCESS, C.K_LATERAL, C.K_PRIVESC)] == [7, 8, None]
    assert [C.onehot_index(0, k) for k in (C.K_PRIVESC, C.K_LATERAL, C.K_EXECUTION)] == [7, None, None]
