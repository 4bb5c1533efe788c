"""Milan model registrations (ref `lingvo/tasks/milan/params/params.py`)."""

from lingvo_b200.models.milan.params import cxc  # noqa: F401
