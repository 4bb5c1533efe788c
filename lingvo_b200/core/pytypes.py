"""Typing aliases (ref `lingvo/core/pytypes.py`)."""
from typing import List, Mapping, Tuple, TypeVar, Union

import numpy as np
import torch

from lingvo_b200.core import hyperparams
from lingvo_b200.core import nested_map

NpTensor = np.ndarray
NestedMap = nested_map.NestedMap
Params = hyperparams.Params
InstantiableParams = hyperparams.InstantiableParams
T = TypeVar('T')
Nested = Union[T, Tuple[T, ...], List[T], Mapping[str, T], nested_map.NestedMap]
NestedTensor = Nested[torch.Tensor]
NestedBool = Nested[bool]
NestedInt = Nested[int]
