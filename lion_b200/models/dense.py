"""Initialiser of the AdaGN style Linear (reference: models/dense.py:49-68).

`dense(in, out, init_scale)` = nn.Linear with variance-scaling uniform weights and zero bias.
Quirk kept from the reference: its 'fan_avg' mode falls through to fan_out
(models/dense.py:25-26), so bound = sqrt(3 * scale / fan_out).
"""
import math

import torch
import torch.nn as nn


def variance_scaling_init_(tensor, scale):
    fan_out = tensor.shape[0] * (tensor[0][0].numel() if tensor.dim() > 2 else 1)
    gain = 1e-10 if scale == 0 else scale
    bound = math.sqrt(3.0 * gain / max(1.0, fan_out))
    with torch.no_grad():
        return tensor.uniform_(-bound, bound)


def dense(in_channels, out_channels, init_scale=1.0):
    lin = nn.Linear(in_channels, out_channels)
    variance_scaling_init_(lin.weight, init_scale)
    nn.init.zeros_(lin.bias)
    return lin
