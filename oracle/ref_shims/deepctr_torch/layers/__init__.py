"""Names from deepctr-torch `layers` that the reference's hot-path files import."""
import torch
import torch.nn as nn

from .utils import concat_fun, slice_arrays  # noqa: F401


def activation_layer(act_name, hidden_size=None, dice_dim=2):
    if isinstance(act_name, str):
        n = act_name.lower()
        if n == 'relu':
            return nn.ReLU(inplace=True)
        if n == 'sigmoid':
            return nn.Sigmoid()
        if n == 'linear':
            return nn.Identity()
        if n == 'prelu':
            return nn.PReLU()
        raise NotImplementedError(act_name)
    if issubclass(act_name, nn.Module):
        return act_name()
    raise NotImplementedError


class PredictionLayer(nn.Module):
    def __init__(self, task='binary', use_bias=True, **kwargs):
        if task not in ["binary", "multiclass", "regression"]:
            raise ValueError("task must be binary,multiclass or regression")
        super(PredictionLayer, self).__init__()
        self.use_bias = use_bias
        self.task = task
        if self.use_bias:
            self.bias = nn.Parameter(torch.zeros((1,)))

    def forward(self, X):
        output = X
        if self.use_bias:
            output += self.bias
        if self.task == "binary":
            output = torch.sigmoid(output)
        return output


class DNN(nn.Module):  # imported by name only on the SATrans path
    def __init__(self, *a, **k):
        raise NotImplementedError("shim")


class InteractingLayer(nn.Module):  # imported by name only on the SATrans path
    def __init__(self, *a, **k):
        raise NotImplementedError("shim")
