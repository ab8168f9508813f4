"""Deterministic "random-init weights" for the oracle and the tests.  Test infrastructure only.

The reference applies no custom initialisation (nothing in fastscnn.py / contextnet.py),
so "random-init weights" means ``torch.manual_seed(s); fastscnn(3, 19)``: every
``nn.Conv2d`` draws ``kaiming_uniform_(a=sqrt(5))`` = U(+-1/sqrt(fan_in)) for its weight
(and U(+-1/sqrt(fan_in)) for a bias) in construction order; ``nn.BatchNorm2d`` draws
nothing (weight 1, bias 0, running_mean 0, running_var 1, num_batches_tracked 0).
``tests/golden/*_spec.json`` (written by ``oracle/make_golden.py`` from the reference's
own ``state_dict``) lists the keys and shapes in that order; drawing the same uniforms
in the same order reproduces the reference's weights bit for bit
(checked in ``tests/test_oracle.py``).
"""
import json
import math
import os

import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                          'tests', 'golden')


def load_spec(arch):
    with open(os.path.join(GOLDEN_DIR, '%s_spec.json' % arch)) as f:
        return json.load(f)


def init_state(arch, seed=0, in_channels=3, out_channels=19):
    """state_dict (fp32 CPU tensors) equal to ``torch.manual_seed(seed); <arch>(3, 19)``."""
    spec = load_spec(arch)
    assert in_channels == 3 and out_channels == 19, 'spec fixtures are for (3, 19)'
    torch.manual_seed(seed)
    sd = {}
    fan_in = None
    for key, shape in spec:
        if key.endswith('running_mean'):
            sd[key] = torch.zeros(shape)
        elif key.endswith('running_var'):
            sd[key] = torch.ones(shape)
        elif key.endswith('num_batches_tracked'):
            sd[key] = torch.zeros((), dtype=torch.int64)
        elif len(shape) == 4:                       # conv weight
            fan_in = shape[1] * shape[2] * shape[3]
            # same float arithmetic as torch.nn.init.kaiming_uniform_(a=sqrt(5))
            gain = math.sqrt(2.0 / (1 + math.sqrt(5) ** 2))
            bound = math.sqrt(3.0) * (gain / math.sqrt(fan_in))
            sd[key] = torch.empty(shape).uniform_(-bound, bound)
        elif key.endswith('bias') and key[:-4] + 'weight' in sd and sd[key[:-4] + 'weight'].dim() == 4:
            bound = 1.0 / math.sqrt(fan_in)          # conv bias, same fan_in as its weight
            sd[key] = torch.empty(shape).uniform_(-bound, bound)
        elif key.endswith('weight'):                 # BN weight
            sd[key] = torch.ones(shape)
        else:                                        # BN bias
            sd[key] = torch.zeros(shape)
    return sd
