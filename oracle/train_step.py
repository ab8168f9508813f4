"""Oracle for one optimisation step (``engine.py:24-39`` minus apex amp) and for the
CPU arm of ``bench.py``.  Test infrastructure only.
"""
import copy
import math

import torch
import torch.nn.functional as F

from . import fastscnn as o_fastscnn
from . import contextnet as o_contextnet

_BUFFER_TAGS = ('running_mean', 'running_var', 'num_batches_tracked')


def split_state(sd):
    """fp32 CPU clone of a ``state_dict``; parameters become grad-requiring leaves."""
    out = {}
    for k, v in sd.items():
        v = v.detach().to('cpu')
        if k.endswith(_BUFFER_TAGS):
            out[k] = v.clone()
        else:
            out[k] = v.float().clone().requires_grad_(True)
    return out


def param_names(sd):
    return [k for k in sd if not k.endswith(_BUFFER_TAGS)]


def model_forward(arch, sd, x, training, dropout_mask=None):
    if arch == 'fastscnn':
        return o_fastscnn.forward(sd, x, training, dropout_mask)
    if arch.startswith('contextnet'):
        scale = {'contextnet12': 2, 'contextnet14': 4, 'contextnet18': 8}[arch]
        return o_contextnet.forward(sd, x, scale, training, dropout_mask)
    raise ValueError(arch)


def loss_and_grads(arch, sd, x, y, dropout_mask=None, ignore_index=255):
    """forward (train mode) + CE + backward; returns (loss, logits, {name: grad})."""
    names = param_names(sd)
    for k in names:
        sd[k].grad = None
    logits = model_forward(arch, sd, x, True, dropout_mask)
    loss = F.cross_entropy(logits, y, ignore_index=ignore_index)
    loss.backward()
    return loss.detach(), logits.detach(), {k: sd[k].grad for k in names}


class AdamW:
    """``torch.optim.AdamW`` defaults as the scripts use them
    (scripts/train_fastscnn.py:125-129: lr, weight_decay=1e-5; betas (0.9, 0.999),
    eps 1e-8), restated explicitly."""

    def __init__(self, sd, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-5):
        self.sd, self.lr, self.betas, self.eps, self.wd = sd, lr, betas, eps, weight_decay
        self.names = param_names(sd)
        self.m = {k: torch.zeros_like(sd[k]) for k in self.names}
        self.v = {k: torch.zeros_like(sd[k]) for k in self.names}
        self.t = 0

    @torch.no_grad()
    def step(self):
        self.t += 1
        b1, b2 = self.betas
        bc1 = 1 - b1 ** self.t
        bc2 = 1 - b2 ** self.t
        for k in self.names:
            p, g = self.sd[k], self.sd[k].grad
            if g is None:
                continue
            p.mul_(1 - self.lr * self.wd)
            self.m[k].mul_(b1).add_(g, alpha=1 - b1)
            self.v[k].mul_(b2).addcmul_(g, g, value=1 - b2)
            denom = (self.v[k].sqrt() / math.sqrt(bc2)).add_(self.eps)
            p.addcdiv_(self.m[k], denom, value=-self.lr / bc1)


def train_step(arch, sd, opt, x, y, dropout_mask=None):
    """``update_fn`` engine.py:24-39: zero_grad, forward, loss, backward, step, loss.item()."""
    loss, _, _ = loss_and_grads(arch, sd, x, y, dropout_mask)
    opt.step()
    return float(loss)
