"""LM model registrations (ref `lingvo/tasks/lm/params/params.py`)."""

from lingvo_b200.models.lm.params import one_billion_wds  # noqa: F401
from lingvo_b200.models.lm.params import synthetic_packed_input  # noqa: F401
from lingvo_b200.models.lm.params import wiki_bert  # noqa: F401
