from .convert import from_networkx  # noqa: F401
