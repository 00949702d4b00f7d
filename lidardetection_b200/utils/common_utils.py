"""The one helper of pcdet/utils/common_utils.py the hot path uses (common_utils.py:46-49)."""
import numpy as np
import torch


def check_numpy_to_torch(x):
    if isinstance(x, np.ndarray):
        return torch.from_numpy(x).float(), True
    return x, False
