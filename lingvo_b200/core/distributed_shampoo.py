"""Distributed Shampoo (ref `lingvo/core/distributed_shampoo.py`); see
`optimizer.DistributedShampoo` and `matrix_functions.py`."""
from lingvo_b200.core.optimizer import DistributedShampoo  # noqa: F401
