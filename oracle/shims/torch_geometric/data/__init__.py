import torch


class Data:
    """Attribute bag with the subset of torch_geometric.data.Data behaviour that
    cyberbattle_env_compressed.py:257-262 relies on (`in`, attribute access, `.to`)."""

    def __init__(self, **kw):
        for k, v in kw.items():
            setattr(self, k, v)

    def __contains__(self, key):
        return key in self.__dict__

    def keys(self):
        return list(self.__dict__.keys())

    def __getitem__(self, k):
        return self.__dict__[k]

    def __setitem__(self, k, v):
        self.__dict__[k] = v

    def to(self, device):
        for k, v in list(self.__dict__.items()):
            if torch.is_tensor(v):
                self.__dict__[k] = v.to(device)
        return self

    @property
    def num_edges(self):
        return int(self.edge_index.shape[1])

    @property
    def num_nodes(self):
        return int(self.x.shape[0]) if "x" in self else int(self.__dict__.get("_num_nodes", 0))


class DataLoader:  # only imported (never used) on the hot path
    def __init__(self, *a, **k):
        raise NotImplementedError("shim: DataLoader is training-only")
