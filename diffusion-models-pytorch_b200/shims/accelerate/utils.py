"""accelerate.utils subset: set_seed, DistributedDataParallelKwargs, DistributedType."""
import os
import random

import numpy as np
import torch


class DistributedType:
    NO = 'NO'
    MULTI_GPU = 'MULTI_GPU'


class DistributedDataParallelKwargs:
    def __init__(self, **kwargs):
        self.kwargs = kwargs

    def to_kwargs(self):
        return dict(self.kwargs)


def set_seed(seed: int, device_specific: bool = False):
    """accelerate semantics: with device_specific every process uses seed + its process index."""
    if device_specific:
        seed += int(os.environ.get('RANK', '0'))
    random.seed(seed)
    np.random.seed(seed % (2 ** 32))
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)
    return seed
