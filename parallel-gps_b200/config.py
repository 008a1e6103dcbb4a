"""Mirror of pssgp/config.py plus the default float (gpflow.config.default_float in the reference)."""
import torch

NUMBER_OF_BALANCING_STEPS = 10  # pssgp/config.py:6

_DEFAULT_FLOAT = torch.float64

# Matern52.get_sde through its closed form with an analytic gradient (kernels/matern.py) instead of the generic
# balance + Lyapunov-solve path under torch autograd: same values to rounding, ~1 ms less host time per training step.
FAST_MATERN_SDE = True


# Training path of StateSpaceGP: (F, Pinf, H) and their hyper-parameter Jacobians from the native builder
# (kernels/native.py, C ABI pssgp_sde_batch_jac) instead of torch autograd through the Python get_sde, for every kernel
# inside the native grammar (sums of products of Matern / RBF / Periodic).  Same values and gradients to rounding.
NATIVE_SDE = True


def set_number_balancing_steps(n_balancing_steps):
    """pssgp/config.py:9-16."""
    global NUMBER_OF_BALANCING_STEPS
    NUMBER_OF_BALANCING_STEPS = n_balancing_steps


def default_float():
    return _DEFAULT_FLOAT


def set_default_float(dtype):
    """float64 (reference default) or float32 (opt-in, experiments' --dtype flag)."""
    global _DEFAULT_FLOAT
    import numpy as np
    if dtype in (torch.float32, np.float32, "float32"):
        _DEFAULT_FLOAT = torch.float32
    elif dtype in (torch.float64, np.float64, "float64"):
        _DEFAULT_FLOAT = torch.float64
    else:
        raise TypeError(f"unsupported default float {dtype}")
