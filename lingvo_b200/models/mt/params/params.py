"""MT model registrations (ref `lingvo/tasks/mt/params/params.py`)."""

from lingvo_b200.models.mt.params import wmt14_en_de  # noqa: F401
from lingvo_b200.models.mt.params import wmtm16_en_de  # noqa: F401
from lingvo_b200.models.mt.params.xendec import wmt14_en_de as xendec_wmt14_en_de  # noqa: F401
