"""v1 API names (ref `lingvo/core/tpu_embedding_layers_v1.py`)."""
from lingvo_b200.core.tpu_embedding_layers import *  # noqa: F401,F403
