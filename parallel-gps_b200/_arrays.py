"""Host<->device plumbing: torch is used only for device memory, pinned staging and streams."""
import numpy as np
import torch

from . import _lib

_DT = {torch.float64: _lib.PSSGP_F64, torch.float32: _lib.PSSGP_F32}


def require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("pssgp_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")


def dtype_code(t):
    try:
        return _DT[t.dtype]
    except KeyError:
        raise TypeError(f"pssgp_b200 supports float64/float32, got {t.dtype}")


def torch_dtype(x):
    if isinstance(x, torch.Tensor):
        return x.dtype
    dt = np.asarray(x).dtype
    return torch.float32 if dt == np.float32 else torch.float64


class _Staging:
    """Reusable pinned host buffers so numpy callers do not pay cudaHostAlloc on every call.

    A buffer is handed to an asynchronous H2D / D2H copy; the CUDA event recorded right after that copy was
    enqueued (``mark``) is waited for before the buffer is given out again (``get``), so a second call with the
    same key never overwrites bytes a still-queued DMA is going to read."""

    def __init__(self):
        self._bufs = {}
        self._events = {}

    def get(self, key, nbytes):
        ev = self._events.pop(key, None)
        if ev is not None:
            ev.synchronize()
        buf = self._bufs.get(key)
        if buf is None or buf.numel() < nbytes:
            buf = torch.empty(max(nbytes, 1), dtype=torch.uint8, pin_memory=True)
            self._bufs[key] = buf
        return buf

    def mark(self, key, device):
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(device))
        self._events[key] = ev


_staging = _Staging()


def is_device_tensor(x):
    return isinstance(x, torch.Tensor) and x.is_cuda


def pick_device(*xs):
    for x in xs:
        if is_device_tensor(x):
            return x.device
    require_cuda()
    return torch.device("cuda", torch.cuda.current_device())


def from_dlpack(x):
    """Zero-copy view of any DLPack producer (an object with ``__dlpack__`` — TensorFlow via
    ``tf.experimental.dlpack``, CuPy, JAX, numpy — or a raw DLPack capsule) as a torch tensor."""
    return torch.from_dlpack(x)


def to_device(x, dtype, device, key=None):
    """Returns a contiguous CUDA tensor of `dtype` on `device` (async H2D from pinned memory for host data).
    Accepts torch tensors, numpy arrays / array-likes, and any DLPack producer (zero-copy when the producer's memory
    already lives on `device` in the right dtype: the north star's "zero-copy DLPack views of TF tensors")."""
    if not isinstance(x, (torch.Tensor, np.ndarray)) and (hasattr(x, "__dlpack__") or type(x).__name__ == "PyCapsule"):
        x = from_dlpack(x)
    if isinstance(x, torch.Tensor):
        if x.is_cuda:
            t = x.detach().to(device=device, dtype=dtype).contiguous()
            # the streaming kernels move 16-byte pieces: a sliced view may start off a 16-byte boundary
            return t if t.data_ptr() % 16 == 0 else t.clone()
        src = x.detach().to(dtype).contiguous()
    else:
        src = torch.from_numpy(np.ascontiguousarray(np.asarray(x), dtype=np.float64 if dtype == torch.float64 else np.float32))
    if src.numel() == 0:
        return torch.empty(src.shape, dtype=dtype, device=device)
    if not src.is_pinned():
        nbytes = src.numel() * src.element_size()
        if key is None:
            # no reusable slot was named: a fresh pinned tensor (torch's caching host allocator keeps the block
            # alive until the copy that uses it has run)
            stage = torch.empty(src.shape, dtype=dtype, pin_memory=True)
            stage.copy_(src)
            return stage.to(device, non_blocking=True)
        stage = _staging.get(key, nbytes)[:nbytes].view(dtype).view(src.shape)
        stage.copy_(src)
        out = stage.to(device, non_blocking=True)
        _staging.mark(key, device)
        return out
    return src.to(device, non_blocking=True)


def to_host_into(t, out):
    """Device->host read of a result straight into a caller-provided host tensor (pinned for an asynchronous
    copy at full PCIe speed); returns `out` after the copy has completed."""
    if not isinstance(out, torch.Tensor) or out.is_cuda:
        raise TypeError("out must be a host torch tensor")
    if out.numel() != t.numel() or out.dtype != t.dtype or not out.is_contiguous():
        raise ValueError(f"out must be a contiguous {t.dtype} host tensor with {t.numel()} elements")
    out.view(t.shape).copy_(t, non_blocking=True)
    torch.cuda.current_stream(t.device).synchronize()
    return out


def to_host(t, out=None):
    """Device->host read of a result (into pinned memory), returned as numpy."""
    nbytes = t.numel() * t.element_size()
    stage = _staging.get(("out", out if out is not None else t.shape, t.dtype), nbytes)[:nbytes].view(t.dtype).view(t.shape)
    stage.copy_(t, non_blocking=True)
    torch.cuda.current_stream(t.device).synchronize()
    return stage.numpy().copy()


def ptr(t):
    return None if t is None else t.data_ptr()


def stream_ptr(device):
    return torch.cuda.current_stream(device).cuda_stream
