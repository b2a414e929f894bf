"""Test-only restatement of the few torch_geometric 2.5.3 symbols the reference's hot path
touches (gae/model.py:14,17 ; _env/cyberbattle_env_compressed.py:30).  Plain torch; see
oracle/shims/README.md.  NOT used by the product."""
from . import nn, utils, data  # noqa: F401
