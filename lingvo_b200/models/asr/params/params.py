"""ASR model registrations (ref `lingvo/tasks/asr/params/params.py`): importing this
module registers every ASR experiment."""

from lingvo_b200.models.asr.params import librispeech  # noqa: F401
