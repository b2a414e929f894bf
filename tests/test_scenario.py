"""Scenario compiler: table invariants, reach counts against the reference Model's own shortest-path tables
(recorded in the golden fixtures), (de)serialisation round trip."""
import glob
import os

import numpy as np
import pytest

import ccbs_b200 as cb
import ccbs_b200.constants as C
from ccbs_b200 import scenario as sc


def test_compiled_tables_are_consistent():
    specs = [cb.synthetic_spec(40 + k, n) for k, n in enumerate((5, 17, 33, 64))]
    t = cb.compile_scenarios(specs)
    assert t.words == 2 and t.max_nodes == 64
    assert t.sc_node_off[-1] == sum(s.num_nodes for s in specs) == len(t.nd_value)
    assert len(t.nd_row_off) == 2 * t.sc_node_off[-1] + 1 and np.all(np.diff(t.nd_row_off) >= 0)
    assert t.nd_row_off[-1] == len(t.row_packed) == len(t.row_inst)
    # every candidate row points at an instance of the right vulnerability and carries a legal one-hot slot
    for r in range(0, len(t.row_packed), 7):
        packed, inst = int(t.row_packed[r]), int(t.row_inst[r])
        kind, oh, gv = (packed >> 20) & 15, (packed >> 24) & 15, packed & 0xFFFFF
        assert oh < C.OUTCOME_DIM and kind != C.K_EXECUTION and gv < t.vemb32.shape[0]
        assert (int(t.vi_kinds_any[inst]) >> kind) & 1
    # inst_of is the inverse of (node, vi_ulocal)
    for s, spec in enumerate(specs):
        U = int(t.sc_num_uvuln[s])
        block = t.inst_of[t.sc_instof_off[s]:t.sc_instof_off[s + 1]].reshape(spec.num_nodes, max(U, 1))
        for j, nd in enumerate(spec.nodes):
            assert (block[j] >= 0).sum() == len(nd.vulns)
            for inst in block[j][block[j] >= 0]:
                assert t.uvuln_global[t.sc_uvuln_off[s] + t.vi_ulocal[inst]] < t.vemb32.shape[0]
    assert np.allclose(t.vnorm2, (t.vemb64 ** 2).sum(1))
    # discoverable_amount = N + sum(2*has_data + !visible)   (cyberbattle_env.py:279-288)
    for s, spec in enumerate(specs):
        want = spec.num_nodes + sum(2 * nd.has_data + (not nd.visible) for nd in spec.nodes)
        assert t.sc_discoverable_amount[s] == want


def test_global_vulnerability_table_is_deduplicated():
    pool = cb.synthetic_vuln_pool(5, 50)
    specs = [cb.synthetic_spec(60 + k, 12, pool=pool) for k in range(6)]
    t = cb.compile_scenarios(specs)
    assert t.vemb32.shape[0] <= 50 and t.sc_uvuln_off[-1] > t.vemb32.shape[0]


def test_reach_counts_match_reference_models(golden_dir):
    """The golden fixtures carry ownable / discoverable / disruptable counts taken from the reference Model's
    all-pairs shortest-path tables; compile_scenarios raises if its own graph restatement disagrees."""
    from oracle import gen_golden as gg
    for path in sorted(glob.glob(os.path.join(golden_dir, "*.npz"))):
        spec = gg.load_case(path)["spec"]
        assert spec.ref_counts is not None
        cb.compile_scenarios([spec], check_ref_counts=True)


def test_spec_roundtrip():
    pool = cb.synthetic_vuln_pool(9, 30)
    g = cb.synthetic_input_graph(3, 9, pool=pool)
    spec = cb.spec_from_input_graph(g, 11)
    emb, fv = sc.embeddings_of_input_graph(g)
    back = sc.spec_from_dict(sc.spec_to_dict(spec), emb, fv)
    a, b = cb.compile_scenarios([spec]), cb.compile_scenarios([back])
    for name in ("row_packed", "vi_flags", "vi_success", "nd_ownable", "recon_nodes", "outblock", "vemb64"):
        assert np.array_equal(getattr(a, name), getattr(b, name)), name


def test_limits():
    with pytest.raises(ValueError):
        cb.compile_scenarios([cb.synthetic_spec(1, 129, services_range=(1, 1), vulns_per_service_range=(1, 2))])
