import numpy as np
import torch


def concat_fun(inputs, axis=-1):
    if len(inputs) == 1:
        return inputs[0]
    return torch.cat(inputs, dim=axis)


def slice_arrays(arrays, start=None, stop=None):
    if arrays is None:
        return [None]
    if isinstance(arrays, np.ndarray):
        arrays = [arrays]
    if isinstance(start, list) and stop is not None:
        raise ValueError('The stop argument has to be None if the value of start is a list.')
    elif isinstance(arrays, list):
        if hasattr(start, '__len__'):
            if hasattr(start, 'shape'):
                start = start.tolist()
            return [None if x is None else x[start] for x in arrays]
        return [None if x is None else x[start:stop] for x in arrays]
    if hasattr(start, '__len__'):
        if hasattr(start, 'shape'):
            start = start.tolist()
        return arrays[start]
    if hasattr(start, '__getitem__'):
        return arrays[start:stop]
    return [None]
