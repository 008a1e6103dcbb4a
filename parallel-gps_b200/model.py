"""Mirror of the reference's pssgp/model.py: ``StateSpaceGP(data, kernel, noise_variance, parallel,
max_parallel)`` with ``maximum_log_likelihood_objective()`` (:113-117) and ``predict_f(Xnew)`` (:92-111).

The log-likelihood is a torch autograd node whose forward is  discretise -> pkf  and whose backward is
the hand-written adjoint scan + discretisation adjoint (C ABI ``pssgp_pkf_backward`` /
``pssgp_discretise_backward``), i.e. the role tf.custom_gradient plays in a TF binding
(INTEGRATION.md).  Gradients then flow through the d x d SDE construction to the kernel's
unconstrained variables exactly as in the reference's test (tests/test_gp_vs_kfs.py:53-67).
"""
import numpy as np
import torch

from . import _arrays as A
from . import _lib
from . import config as pssgp_config
from . import ops
from .kalman.base import LGSSM
from .kernels.base import time_steps
from .params import Parameter


class _LogLikelihood(torch.autograd.Function):
    """ll(F, Pinf, H, R; dts, y) with P0 = Pinf (kernels/base.py:47)."""

    @staticmethod
    def forward(ctx, F, Pinf, H, R, dts, y, sequential=False):
        device, dtype = dts.device, dts.dtype
        # the d x d SDE goes to the device as ONE pinned, asynchronous copy (four pageable copies would each
        # synchronise the host with the stream)
        d = F.shape[0]
        packed = torch.cat([F.detach().reshape(-1), Pinf.detach().reshape(-1), H.detach().reshape(-1),
                            R.detach().reshape(-1)])
        pd = A.to_device(packed, dtype, device, "ll_sde") if not packed.is_cuda else packed.to(dtype)
        Fd, Pd = pd[:d * d].view(d, d), pd[d * d:2 * d * d].view(d, d)
        Hd, Rd = pd[2 * d * d:2 * d * d + d], pd[2 * d * d + d:2 * d * d + d + 1]
        Fs, Qs = ops.discretise(Fd, Pd, dts)
        ctx.host = (F.device, F.dtype, tuple(H.shape), tuple(R.shape))
        if sequential:
            # parallel=False (model.py:74-77): the sequential Kalman filter (C ABI pssgp_kf); the gradient comes from
            # the adjoint scan on its filtered moments (the gradient is a function of (Fs, Qs, y, fms, fPs) only)
            fms, fPs, ll, _, _ = ops.kf(Pd, Fs, Qs, Hd, Rd, y)
            if any(ctx.needs_input_grad[:4]):
                one = torch.ones(1, dtype=dtype, device=device)
                dP0, dFs, dQs, dH, dR = ops.pkf_backward(Pd, Fs, Qs, Hd, Rd, y, fms, fPs, one)
                dF, dPinf = ops.discretise_backward(Fd, Pd, dts, Fs, dFs, dQs)
                ctx.save_for_backward(dF, dPinf + dP0, dH, dR)
        elif any(ctx.needs_input_grad[:4]):
            # training step: the fused filter + adjoint call (C ABI pssgp_pkfs_grad without smoother) with unit upstream
            # gradient, then the discretisation adjoint; backward() only scales (the gradient is linear in g).  One pass
            # over (Fs, Qs, y) fewer than pkf followed by pkf_backward, and nothing of size N is kept for backward.
            one = torch.ones(1, dtype=dtype, device=device)
            (fms, fPs, ll), _, (dP0, dFs, dQs, dH, dR) = ops.pkfs_grad(Pd, Fs, Qs, Hd, Rd, y, one, want_smoother=False)
            dF, dPinf = ops.discretise_backward(Fd, Pd, dts, Fs, dFs, dQs)
            # ll and the four d x d sized gradients come back in ONE device -> host copy (one synchronisation per
            # training step; backward() is host arithmetic only)
            packed = torch.cat([ll.reshape(-1), dF.reshape(-1), (dPinf + dP0).reshape(-1), dH.reshape(-1),
                                dR.reshape(-1)]).to(device=F.device, dtype=F.dtype)
            ctx.save_for_backward(packed[1:])
            return packed[0].clone()
        else:
            fms, fPs, ll, _ = ops.pkf(Pd, Fs, Qs, Hd, Rd, y)
        return ll[0].to(device=F.device, dtype=F.dtype)

    @staticmethod
    def backward(ctx, g):
        hdev, hdt, hshape, rshape = ctx.host
        if len(ctx.saved_tensors) == 1:
            packed = ctx.saved_tensors[0]          # already on the host (forward's single read-back)
            d = hshape[-1]
        else:
            dF, dPinf, dH, dR = ctx.saved_tensors  # sequential mode: still on the device
            d = dF.shape[0]
            packed = torch.cat([dF.reshape(-1), dPinf.reshape(-1), dH.reshape(-1), dR.reshape(-1)]).to(device=hdev, dtype=hdt)
        packed = packed * g.detach().to(device=hdev, dtype=hdt)
        gF, gP = packed[:d * d].reshape(d, d), packed[d * d:2 * d * d].reshape(d, d)
        gH, gR = packed[2 * d * d:2 * d * d + d].reshape(hshape), packed[2 * d * d + d:].reshape(rshape)
        return gF, gP, gH, gR, None, None, None


def _merge_sorted_idx(a, b, *args):
    """model.py:15-55: merge two sorted 1-d device tensors (and companion data) without a sort.  Also returns the
    positions of the elements of `a` and of `b` in the merged array.  No boolean mask is materialised and nothing
    synchronises with the host: the positions of the longer array follow from a second searchsorted
    (its element i comes after the i elements before it and after every element of the shorter array that is
    <= it, ties putting the shorter array first exactly like the reference's scatter)."""
    swapped = a.shape[0] < b.shape[0]
    if swapped:
        a, b = b, a
        args = tuple((j, i) for i, j in args)
    na, nb = a.shape[0], b.shape[0]
    b_idx = torch.arange(nb, device=a.device) + torch.searchsorted(a, b)
    a_idx = torch.arange(na, device=a.device) + torch.searchsorted(b, a, right=True)

    def inner(u, v):
        c = torch.empty((na + nb,) + tuple(u.shape[1:]), dtype=u.dtype, device=u.device)
        c[b_idx] = v
        c[a_idx] = u
        return c

    merged = (inner(a, b),) + tuple(inner(i, j) for i, j in args)
    return merged, ((b_idx, a_idx) if swapped else (a_idx, b_idx))


def _merge_sorted(a, b, *args):
    """model.py:15-55 (same returns as the reference)."""
    return _merge_sorted_idx(a, b, *args)[0]


class StateSpaceGP:
    """model.py:58-117.  ``parallel=True`` runs the temporally-parallel scan kernels (pkf / pkfs, model.py:78-84);
    ``parallel=False`` — the reference's default — runs the sequential Kalman filter and RTS smoother (kf / kfs,
    model.py:73-77; C ABI ``pssgp_kf`` / ``pssgp_ks``): one GPU thread or warp walks the series, which is the
    comparator, not the fast path (≈ 0.3 µs per step).  ``max_parallel`` (depth cap of TFP's recursion) is accepted
    and unused."""

    def __init__(self, data, kernel, noise_variance=1.0, parallel=False, max_parallel=10000, device=None):
        A.require_cuda()
        self.noise_variance = Parameter(noise_variance, name="noise_variance")
        self.kernel = kernel
        self.parallel = parallel
        self.max_parallel = max_parallel
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.data = data

    @property
    def data(self):
        return self._data

    @data.setter
    def data(self, data):
        ts, ys = data
        dtype = pssgp_config.default_float()
        tsd = A.to_device(ts, dtype, self.device, "data_ts").reshape(-1, 1)
        ysd = A.to_device(ys, dtype, self.device, "data_ys").reshape(-1, 1)
        if ysd.shape[0] != tsd.shape[0]:
            raise ValueError("ts and ys must have the same length")
        self._data = (tsd, ysd)

    # --- gpflow.Module stand-ins -------------------------------------------------------------
    @property
    def parameters(self):
        return self.kernel.parameters + [self.noise_variance]

    @property
    def trainable_variables(self):
        return [p.unconstrained_variable for p in self.parameters if p.trainable]

    def log_prior_density(self):
        out = torch.zeros((), dtype=torch.float64)
        for p in self.parameters:
            if p.trainable:
                out = out + p.log_prior_density()
        return out

    def log_posterior_density(self):
        return self.maximum_log_likelihood_objective() + self.log_prior_density()

    def training_loss(self):
        return -self.log_posterior_density()

    # --- the path ----------------------------------------------------------------------------
    def _make_model(self, ts):
        """model.py:86-90."""
        R = self.noise_variance.value.reshape(1, 1)
        if not torch.is_grad_enabled():
            # prediction path: the d x d SDE (balancing + Lyapunov solve on the host) only depends on the
            # hyper-parameters; keep the one built for the current values
            from .kernels.base import _get_ssm
            return _get_ssm(self._cached_sde(), ts, R, 0.)
        return self.kernel.get_ssm(ts, R)

    def _cached_sde(self):
        key = tuple(tuple(p.unconstrained_variable.detach().reshape(-1).tolist()) for p in self.kernel.parameters)
        hit = getattr(self, "_sde_cache", None)
        if hit is None or hit[0] != key:
            with torch.no_grad():
                hit = (key, self.kernel.get_sde())
            self._sde_cache = hit
        return hit[1]

    def maximum_log_likelihood_objective(self):
        """model.py:113-117; differentiable w.r.t. trainable_variables."""
        ts, Y = self._data
        sde3 = None
        if pssgp_config.NATIVE_SDE:
            from .kernels import native
            sde3 = native.native_sde(self.kernel)      # None outside the native grammar
        if sde3 is None:
            sde = self.kernel.get_sde()
            sde3 = (sde.F, sde.P0, sde.H)
        R = self.noise_variance.value.reshape(1, 1)
        dts = time_steps(ts, 0., ts.dtype, ts.device)
        return _LogLikelihood.apply(sde3[0], sde3[1], sde3[2], R, dts, Y.reshape(-1), not self.parallel)

    def synchronize(self):
        """Waits for everything this model has enqueued, including the read-back of ``predict_f(non_blocking=True)``."""
        torch.cuda.current_stream(self.device).synchronize()
        side = getattr(self, "_copy_stream", None)
        if side is not None:
            side.synchronize()

    def predict_f(self, Xnew, full_cov=False, full_output_cov=False, out=None, non_blocking=False):
        """model.py:92-111: merge query times as NaN observations, filter + smooth, project with H.
        Returns (mean[K,1], var[K,1]) — numpy for host inputs, CUDA tensors for device inputs.
        ``out=(mean_buf, var_buf)``: host torch tensors (pinned for full PCIe speed) that receive the result
        directly, without the intermediate staging copy of the numpy path; they are returned.
        ``non_blocking=True`` (with ``out``): returns as soon as the work is enqueued; the read-back runs on a copy
        stream behind the kernels, so whatever the caller enqueues next (a training step) overlaps it.  The buffers
        are valid after ``model.synchronize()`` (or ``torch.cuda.synchronize()``)."""
        ts, ys = self._data
        dtype, dev = ts.dtype, ts.device
        Xd = A.to_device(Xnew, dtype, dev, "Xnew").reshape(-1)
        K = Xd.shape[0]
        if ys.shape[1] != 1:
            raise ValueError("single-output models only (pssgp/model.py:72)")
        with torch.no_grad():
            # model.py:95-104 in one C-ABI call: merged grid, NaN observations at the queries, time steps, query rows
            _, yv, dts, q_idx = ops.merge_queries(ts.reshape(-1).contiguous(), ys.reshape(-1).contiguous(), Xd.contiguous())
            sde = self._cached_sde()
            d = sde.F.shape[0]
            # the d x d SDE goes to the device as ONE pinned asynchronous copy
            pd = A.to_device(torch.cat([sde.F.reshape(-1), sde.P0.reshape(-1), sde.H.reshape(-1),
                                        self.noise_variance.value.detach().reshape(-1)]), dtype, dev, "pred_sde")
            Fd, Pd = pd[:d * d].view(d, d), pd[d * d:2 * d * d].view(d, d)
            Hd, Rd = pd[2 * d * d:2 * d * d + d], pd[2 * d * d + d:2 * d * d + d + 1]
            Fs, Qs = ops.discretise(Fd, Pd, dts)
            ssm = LGSSM(Pd, Fs, Qs, Hd.reshape(1, -1), Rd.reshape(1, 1))
            if not self.parallel:
                # model.py:76-77: kfs = sequential filter (with predicted moments) + sequential RTS smoother
                fms, fPs, _, mps, Pps = ops.kf(ssm.P0, ssm.Fs, ssm.Qs, Hd, Rd, yv, want_ll=False, want_predicted=True)
                sms, sPs = ops.ks(ssm.Fs, fms, fPs, mps, Pps)
                rm, rP = sms.index_select(0, q_idx), sPs.index_select(0, q_idx)
                mean = rm @ Hd.reshape(-1, 1)
                var = torch.einsum("i,kij,j->k", Hd, rP, Hd).reshape(-1, 1)
            elif ops.has_projection(ssm.Fs.shape[1], dtype):
                # fused filter + smoother that emits only (H m, H P H^T) of every smoothed state
                proj = ops.pkfs(ssm.P0, ssm.Fs, ssm.Qs, Hd, Rd, yv, project=True)[3]
                # rows of the queries (the reference's boolean_mask, model.py:107-108).  The (mean, var) pairs are
                # gathered as ONE complex element each: torch's row gather of an [n, 2] array is 40x slower
                # (520 us vs 14 us for 1e6 rows, scripts/gather_test.py)
                cdt = torch.complex128 if dtype == torch.float64 else torch.complex64
                sel = torch.view_as_real(proj.view(cdt).reshape(-1).index_select(0, q_idx))
                mean, var = sel[:, 0:1].contiguous(), sel[:, 1:2].contiguous()
            else:
                # fused filter + smoother (pssgp_pkfs): the warp-level DMMA kernels for 5 <= d <= 32
                sms, sPs = ops.pkfs(ssm.P0, ssm.Fs, ssm.Qs, Hd, Rd, yv)[3:5]
                rm, rP = sms.index_select(0, q_idx), sPs.index_select(0, q_idx)
                mean = rm @ Hd.reshape(-1, 1)
                var = torch.einsum("i,kij,j->k", Hd, rP, Hd).reshape(-1, 1)
        if out is not None and non_blocking:
            cur = torch.cuda.current_stream(dev)
            side = getattr(self, "_copy_stream", None)
            if side is None:
                side = self._copy_stream = torch.cuda.Stream(device=dev)
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                out[0].view(mean.shape).copy_(mean, non_blocking=True)
                out[1].view(var.shape).copy_(var, non_blocking=True)
            mean.record_stream(side)
            var.record_stream(side)
            return out
        if out is not None:
            return A.to_host_into(mean, out[0]), A.to_host_into(var, out[1])
        if A.is_device_tensor(Xnew):
            return mean, var
        return A.to_host(mean, "pred_mean"), A.to_host(var, "pred_var")
