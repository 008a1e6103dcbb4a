"""Framework-neutral binding of the hot path at the DLPack level — what a TensorFlow maintainer wraps.

The reference is TensorFlow code; its tensors reach this library as zero-copy DLPack views
(``tf.experimental.dlpack.to_dlpack(t)``) and the results go back the same way
(``tf.experimental.dlpack.from_dlpack(capsule)``).  Every function here takes DLPack producers (objects with
``__dlpack__`` or raw capsules, memory on the CUDA device) and returns DLPack CAPSULES; nothing is copied.  torch is
only the owner of the output allocations.  ``make_tf_ops()`` builds the ``tf.custom_gradient`` wrappers on top when
TensorFlow is importable (it is not in this image: tests/test_gpu_dlpack.py drives the capsule level with another
producer).

    pkf_ll(P0, Fs, Qs, H, R, y)                  -> (fms, fPs, ll) capsules, ctx      pssgp/kalman/parallel.py:121-152
    pkf_ll_grad(ctx, g_ll)                       -> (dP0, dFs, dQs, dH, dR) capsules  (TF autodiff in the reference)
    pkfs(P0, Fs, Qs, H, R, y)                    -> (sms, sPs) capsules               parallel.py:199-201
    get_ssm(F, Pinf, dts)                        -> (Fs, Qs) capsules                 pssgp/kernels/base.py:29-47
    kf_ll(P0, Fs, Qs, H, R, y)                   -> (fms, fPs, ll) capsules, ctx      pssgp/kalman/sequential.py:11-47
    kfs(P0, Fs, Qs, H, R, y)                     -> (sms, sPs) capsules               sequential.py:71-73
"""
import torch
from torch.utils.dlpack import to_dlpack

from . import _arrays as A
from . import ops


def _view(x):
    t = A.from_dlpack(x) if not isinstance(x, torch.Tensor) else x
    if not t.is_cuda:
        raise ValueError("dlpack_binding expects device memory (zero-copy views of the framework's GPU tensors)")
    return t.contiguous()


def get_ssm(F, Pinf, dts):
    Fs, Qs = ops.discretise(_view(F), _view(Pinf), _view(dts).reshape(-1))
    return to_dlpack(Fs), to_dlpack(Qs)


def pkf_ll(P0, Fs, Qs, H, R, y):
    """Filter + log-likelihood; ctx keeps the views the gradient needs (borrowed inputs + our own outputs)."""
    P0, Fs, Qs, H, R, y = (_view(v) for v in (P0, Fs, Qs, H, R, y))
    fms, fPs, ll, _ = ops.pkf(P0, Fs, Qs, H.reshape(-1), R.reshape(-1), y.reshape(-1))
    ctx = (P0, Fs, Qs, H, R, y, fms, fPs)
    return (to_dlpack(fms), to_dlpack(fPs), to_dlpack(ll)), ctx


def pkf_ll_grad(ctx, g_ll):
    """Gradient of the log-likelihood w.r.t. (P0, Fs, Qs, H, R) given the upstream gradient g_ll [1]."""
    P0, Fs, Qs, H, R, y, fms, fPs = ctx
    g = _view(g_ll).reshape(1).to(Fs.dtype)
    dP0, dFs, dQs, dH, dR = ops.pkf_backward(P0, Fs, Qs, H.reshape(-1), R.reshape(-1), y.reshape(-1), fms, fPs, g)
    return to_dlpack(dP0), to_dlpack(dFs), to_dlpack(dQs), to_dlpack(dH.reshape(H.shape)), to_dlpack(dR.reshape(R.shape))


def pkfs(P0, Fs, Qs, H, R, y):
    P0, Fs, Qs, H, R, y = (_view(v) for v in (P0, Fs, Qs, H, R, y))
    out = ops.pkfs(P0, Fs, Qs, H.reshape(-1), R.reshape(-1), y.reshape(-1))
    return to_dlpack(out[3]), to_dlpack(out[4])


def kf_ll(P0, Fs, Qs, H, R, y):
    """Sequential filter + log-likelihood (StateSpaceGP(parallel=False), model.py:73-75); the ctx is the one of pkf_ll:
    pkf_ll_grad gives the gradient (it is a function of the filtered moments, whichever filter produced them)."""
    P0, Fs, Qs, H, R, y = (_view(v) for v in (P0, Fs, Qs, H, R, y))
    fms, fPs, ll, _, _ = ops.kf(P0, Fs, Qs, H.reshape(-1), R.reshape(-1), y.reshape(-1))
    ctx = (P0, Fs, Qs, H, R, y, fms, fPs)
    return (to_dlpack(fms), to_dlpack(fPs), to_dlpack(ll)), ctx


def kfs(P0, Fs, Qs, H, R, y):
    """Sequential filter + RTS smoother (model.py:76-77)."""
    P0, Fs, Qs, H, R, y = (_view(v) for v in (P0, Fs, Qs, H, R, y))
    fms, fPs, _, mps, Pps = ops.kf(P0, Fs, Qs, H.reshape(-1), R.reshape(-1), y.reshape(-1), want_ll=False, want_predicted=True)
    sms, sPs = ops.ks(Fs, fms, fPs, mps, Pps)
    return to_dlpack(sms), to_dlpack(sPs)


def make_tf_ops():
    """tf.custom_gradient wrappers for a TensorFlow host (the reference's stack).  Usable inside tf.function through
    tf.py_function, exactly as the reference already leaves the graph for balancing (math_utils.py:68)."""
    import tensorflow as tf  # not installable in this image; the capsule-level functions above are what is tested
    to_cap, from_cap = tf.experimental.dlpack.to_dlpack, tf.experimental.dlpack.from_dlpack

    @tf.custom_gradient
    def tf_pkf_ll(P0, Fs, Qs, H, R, y):
        caps, ctx = pkf_ll(*(to_cap(t) for t in (P0, Fs, Qs, H, R, y)))
        fms, fPs, ll = (from_cap(c) for c in caps)

        def grad(g_fms, g_fPs, g_ll):  # only the log-likelihood is differentiated (as in the reference's use)
            dP0, dFs, dQs, dH, dR = (from_cap(c) for c in pkf_ll_grad(ctx, to_cap(tf.reshape(g_ll, (1,)))))
            return dP0, dFs, dQs, dH, dR, None

        return (fms, fPs, ll), grad

    def tf_pkfs(P0, Fs, Qs, H, R, y):
        return tuple(from_cap(c) for c in pkfs(*(to_cap(t) for t in (P0, Fs, Qs, H, R, y))))

    def tf_get_ssm(F, Pinf, dts):
        return tuple(from_cap(c) for c in get_ssm(to_cap(F), to_cap(Pinf), to_cap(dts)))

    return tf_pkf_ll, tf_pkfs, tf_get_ssm
