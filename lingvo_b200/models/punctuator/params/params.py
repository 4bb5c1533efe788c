"""Punctuator model registrations (ref `lingvo/tasks/punctuator/params/params.py`)."""

from lingvo_b200.models.punctuator.params import codelab  # noqa: F401
