""""zero-init" initialisers of the reference (sde_sampler/models/utils.py:7-49): last layers start at scale
1e-6 so that the learned control is ~0 at initialisation.  Host-side only."""
import math

import torch

init_weight_scale = 1e-6


def kaiming_uniform_zeros_(m):
    # kaiming-uniform gain sqrt(2 / (1 + a^2)) with a chosen such that the bound is init_weight_scale * sqrt(3 / fan_in) / sqrt(3)
    return torch.nn.init.kaiming_uniform_(m, a=math.sqrt((6.0 / init_weight_scale ** 2) - 1))


def kaiming_normal_zeros_(m):
    return torch.nn.init.kaiming_normal_(m, a=math.sqrt((6.0 / init_weight_scale ** 2) - 1))


def _fan_in(weight):
    return torch.nn.init._calculate_fan_in_and_fan_out(weight)[0]


def init_bias_uniform_zeros(m, weight):
    fan_in = _fan_in(weight)
    if fan_in > 0:
        bound = init_weight_scale / math.sqrt(fan_in)
        torch.nn.init.uniform_(m, -bound, bound)
    else:
        torch.nn.init.zeros_(m)


def init_bias_uniform_constant(m, weight, val=1.0):
    fan_in = _fan_in(weight)
    if fan_in > 0:
        bound = init_weight_scale / math.sqrt(fan_in)
        torch.nn.init.uniform_(m, val - bound, val + bound)
    else:
        torch.nn.init.constant_(m, val)
