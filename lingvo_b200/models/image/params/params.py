"""Image model registrations (ref `lingvo/tasks/image/params/params.py`)."""

from lingvo_b200.models.image.params import mnist  # noqa: F401
