"""AdaGraft (ref `lingvo/core/adagraft.py`); the implementation lives with the
other optimizers."""
from lingvo_b200.core.optimizer import AdaGraft  # noqa: F401
