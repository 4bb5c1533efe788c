"""Constants (ref `lingvo/core/constants.py`)."""
REFERENCE_ANNOTATION = 'arxiv.org/abs/1902.08295'
