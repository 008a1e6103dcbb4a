"""Stand-in for ``gpflow.optimizers.Scipy`` as the reference's MAP experiments use it
(pssgp/experiments/sunspot/map.py:74-82, co2/map.py, toy_models/map.py):

    opt = Scipy()
    eval_func = opt.eval_func(model.training_loss, model.trainable_variables)   # x -> (loss, gradient)
    x0 = opt.initial_parameters(model.trainable_variables)
    result = scipy.optimize.minimize(eval_func, x0, jac=True, options=dict(maxiter=100))
    # or in one call:  opt.minimize(model.training_loss, model.trainable_variables, options=dict(maxiter=100))

The variables are the model's unconstrained torch leaf tensors (pssgp_b200.params.Parameter); every evaluation of the
closure is one fused filter + adjoint step on the GPU (StateSpaceGP.maximum_log_likelihood_objective + autograd).
"""
import numpy as np
import torch


class Scipy:
    @staticmethod
    def initial_parameters(variables):
        """Concatenation of the flattened variables (gpflow.optimizers.Scipy.initial_parameters)."""
        return np.concatenate([v.detach().reshape(-1).cpu().numpy().astype(np.float64) for v in variables])

    @staticmethod
    def assign(variables, x):
        """Writes the flat vector x back into the variables."""
        x = np.asarray(x, dtype=np.float64)
        off = 0
        with torch.no_grad():
            for v in variables:
                k = v.numel()
                v.copy_(torch.as_tensor(x[off:off + k], dtype=v.dtype).reshape(v.shape))
                off += k
        if off != x.size:
            raise ValueError(f"expected {off} values, got {x.size}")

    @classmethod
    def eval_func(cls, closure, variables, compile=False):
        """x (numpy, float64) -> (loss, gradient) with the variables set to x.  ``compile`` is accepted for signature
        compatibility (the reference traces the closure with tf.function; there is nothing to trace here)."""
        variables = list(variables)

        def _eval(x):
            cls.assign(variables, x)
            loss = closure()
            grads = torch.autograd.grad(loss, variables, allow_unused=True)
            g = np.concatenate([(torch.zeros_like(v) if gi is None else gi).detach().reshape(-1).cpu().numpy().astype(np.float64)
                                for v, gi in zip(variables, grads)])
            return float(loss.detach()), g

        return _eval

    def minimize(self, closure, variables, method="L-BFGS-B", **scipy_kwargs):
        """gpflow.optimizers.Scipy.minimize: scipy.optimize.minimize on the flattened variables; the variables hold
        the optimum afterwards.  Returns scipy's OptimizeResult."""
        import scipy.optimize
        variables = list(variables)
        func = self.eval_func(closure, variables)
        res = scipy.optimize.minimize(func, self.initial_parameters(variables), jac=True, method=method, **scipy_kwargs)
        self.assign(variables, res.x)
        return res
