"""Minimal stand-in for gpflow.Parameter(transform=positive()) (softplus), on torch.

The reference's host stack (GPflow/TensorFlow) is not installable in this image; the host-side
parameter container therefore lives on torch autograd, which also differentiates the tiny d x d SDE
construction.  Everything heavy (per-time-step work) runs in the CUDA library.
"""
import math

import torch


def _softplus_inv(x):
    x = float(x)
    if x <= 0:
        raise ValueError("positive parameter must be > 0")
    return x + math.log(-math.expm1(-x))


class Parameter:
    def __init__(self, value, dtype=torch.float64, trainable=True, name=None):
        self.unconstrained_variable = torch.tensor(_softplus_inv(value), dtype=dtype, requires_grad=bool(trainable))
        self.trainable = bool(trainable)
        self.prior = None  # optional callable log-density on the constrained value
        self.name = name

    @property
    def value(self):
        return torch.nn.functional.softplus(self.unconstrained_variable)

    def assign(self, value):
        with torch.no_grad():
            self.unconstrained_variable.fill_(_softplus_inv(value))

    def numpy(self):
        return self.value.detach().cpu().numpy()

    def log_prior_density(self):
        """gpflow semantics: prior on the constrained value plus log|d constrained / d unconstrained|."""
        if self.prior is None:
            return torch.zeros((), dtype=self.unconstrained_variable.dtype)
        x = self.value
        log_jac = torch.nn.functional.logsigmoid(self.unconstrained_variable)
        return self.prior(x) + log_jac


def set_trainable(param, flag):
    param.trainable = bool(flag)
    param.unconstrained_variable.requires_grad_(bool(flag))
