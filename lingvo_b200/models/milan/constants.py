"""Shared constants (ref `lingvo/tasks/milan/constants.py`)."""


class Modality:
  IMAGE = 'image'
  TEXT = 'text'
  AUDIO = 'audio'


class Split:
  TRAIN = 'Train'
  DEV = 'Dev'
  TEST = 'Test'
