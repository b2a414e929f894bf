import numpy as np
import torch
from torch_geometric.data import Data


def from_networkx(G):
    """Restates torch_geometric.utils.convert.from_networkx (2.5.3) for a DiGraph:
    node index = position in G.nodes; edge order = G.edges order; every node/edge attribute is
    collected per key and stacked.  A graph without edges produces no edge-attribute keys."""
    mapping = dict(zip(G.nodes(), range(G.number_of_nodes())))
    edge_index = torch.empty((2, G.number_of_edges()), dtype=torch.long)
    for i, (src, dst) in enumerate(G.edges()):
        edge_index[0, i] = mapping[src]
        edge_index[1, i] = mapping[dst]
    store = {}
    node_keys = None
    for _, feat in G.nodes(data=True):
        if node_keys is None:
            node_keys = set(feat.keys())
        elif set(feat.keys()) != node_keys:
            raise ValueError("Not all nodes contain the same attributes")
        for k, v in feat.items():
            store.setdefault(str(k), []).append(v)
    edge_keys = None
    for _, _, feat in G.edges(data=True):
        if edge_keys is None:
            edge_keys = set(feat.keys())
        elif set(feat.keys()) != edge_keys:
            raise ValueError("Not all edges contain the same attributes")
        for k, v in feat.items():
            key = f"edge_{k}" if node_keys and k in node_keys else str(k)
            store.setdefault(key, []).append(v)
    out = {}
    for k, vals in store.items():
        try:
            out[k] = torch.as_tensor(np.array(vals))
        except Exception:
            out[k] = vals
    d = Data(edge_index=edge_index.view(2, -1), **out)
    d._num_nodes = G.number_of_nodes()
    return d
