"""Host restatement of csrc/philox.cuh (Philox4x32-10, Salmon et al. SC'11), vectorised with numpy.

Key = the handle's seed; counter = (global env index lo, hi, per-env counter, stream id).  Used by the tests to predict the
device's draws and by the documented sub-sampling rule of the action table (:func:`subset_keep`), which stands in for the
reference's ``np.random.choice`` in ``__balance_action_space_by_outcome`` (cyberbattle_env_compressed.py:553-567)."""
from __future__ import annotations

import numpy as np

_M0, _M1, _W0, _W1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), 0x9E3779B9, 0xBB67AE85
_LO = np.uint64(0xFFFFFFFF)
_S32 = np.uint64(32)

SUBSET_STREAM = 0x80000000      # stream ids with the top bit set: one per action-table row identity (below)


def philox4x32_10(seed: int, env: int, step, stream) -> np.ndarray:
    """The four 32-bit output words, shape [..., 4] (uint32), for broadcastable ``step`` / ``stream`` arrays."""
    step, stream = np.broadcast_arrays(np.asarray(step, dtype=np.uint64) & _LO, np.asarray(stream, dtype=np.uint64) & _LO)
    k0, k1 = int(seed) & 0xFFFFFFFF, (int(seed) >> 32) & 0xFFFFFFFF
    c0 = np.full(step.shape, int(env) & 0xFFFFFFFF, dtype=np.uint64)
    c1 = np.full(step.shape, (int(env) >> 32) & 0xFFFFFFFF, dtype=np.uint64)
    c2, c3 = step.copy(), stream.copy()
    for _ in range(10):
        p0, p1 = _M0 * c0, _M1 * c2
        c0, c1, c2, c3 = ((p1 >> _S32) ^ c1 ^ np.uint64(k0)) & _LO, p1 & _LO, ((p0 >> _S32) ^ c3 ^ np.uint64(k1)) & _LO, p0 & _LO
        k0, k1 = (k0 + _W0) & 0xFFFFFFFF, (k1 + _W1) & 0xFFFFFFFF
    return np.stack([c0, c1, c2, c3], axis=-1).astype(np.uint32)


def row_identity(s, t, kind, vuln_local) -> np.ndarray:
    """Stream id of an action-table row: source node (7 bits), target node (7), outcome kind (4), scenario-local vulnerability
    index (12).  Rows that share all four (two results of one vulnerability with the same outcome class) share the key and keep
    their table order."""
    s, t, kind, v = (np.asarray(x, dtype=np.int64) for x in (s, t, kind, vuln_local))
    if np.any(s >= 128) or np.any(t >= 128) or np.any(kind >= 16) or np.any(v >= 4096):
        raise ValueError("row identity out of range (nodes < 128, kinds < 16, vulnerabilities per scenario < 4096)")
    return (SUBSET_STREAM | (s << 23) | (t << 16) | (kind << 12) | v).astype(np.uint64)


def subset_keep(seed: int, env: int, call: int, identities, k: int) -> np.ndarray:
    """Positions (ascending) of the ``k`` rows kept out of ``len(identities)`` rows of one outcome class.

    Rule: row i gets the 32-bit key ``philox(seed, env, call, identity_i).x``; the k smallest (key, i) pairs stay, in their
    table order.  ``call`` = how many times this env has balanced its action table so far (lifetime counter), so every
    balance draws a fresh uniform subset — the distribution of ``np.random.choice(n, k, replace=False)`` as a SET; the
    reference's dict then holds the kept rows in the order choice returned them, here they keep their insertion order."""
    ids = np.asarray(identities, dtype=np.uint64)
    keys = philox4x32_10(seed, env, call, ids)[..., 0].astype(np.uint64)
    order = np.lexsort((np.arange(len(ids)), keys))
    return np.sort(order[:int(k)])
