"""ctypes binding of libpssgp_b200.so (C ABI: include/pssgp_b200.h).

There is deliberately no fallback: if the shared library is missing or fails to load, importing
this module raises, and every product entry point fails loudly.
"""
import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
# PSSGP_B200_LIB points at another build of the same library (kernel-tuning experiments)
LIB_PATH = os.environ.get("PSSGP_B200_LIB") or os.path.join(_HERE, "lib", "libpssgp_b200.so")

PSSGP_F64 = 0
PSSGP_F32 = 1

_vp = ctypes.c_void_p
_i64 = ctypes.c_int64
_int = ctypes.c_int

# name -> (restype, argtypes); must list every symbol include/pssgp_b200.h declares.
SIGNATURES = {
    "pssgp_version": (_int, []),
    "pssgp_last_error": (ctypes.c_char_p, []),
    "pssgp_create": (_int, [ctypes.POINTER(_vp), _int]),
    "pssgp_destroy": (_int, [_vp]),
    "pssgp_set_option": (_int, [_vp, ctypes.c_char_p, _i64]),
    "pssgp_launch_count": (_i64, [_vp]),
    "pssgp_timing_report": (_int, [_vp, ctypes.c_char_p, _i64]),
    "pssgp_balance_ss": (_int, [_vp, _int, _int, _vp]),
    "pssgp_discretise": (_int, [_vp, _int, _i64, _int, _vp, _vp, _vp, _vp, _vp, _vp]),
    "pssgp_pkf": (_int, [_vp, _int, _i64, _int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _int, _vp, _vp, _vp, _vp, _vp]),
    "pssgp_pks": (_int, [_vp, _int, _i64, _int, _vp, _vp, _vp, _vp, _int, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "pssgp_pkf_summary": (_int, [_vp, _int, _i64, _int, _vp, _vp, _vp, _vp, _vp, _vp, _int, _vp, _vp]),
    "pssgp_filter_fold": (_int, [_vp, _int, _int, _int, _vp, _vp, _vp, _vp, _vp]),
    "pssgp_pks_summary": (_int, [_vp, _int, _i64, _int, _vp, _vp, _vp, _vp, _int, _vp, _vp, _vp, _vp]),
    "pssgp_smoother_fold": (_int, [_vp, _int, _int, _int, _vp, _vp, _vp]),
    "pssgp_pkf_backward": (_int, [_vp, _int, _i64, _int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _int, _vp,
                                  _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "pssgp_pkfs_grad": (_int, [_vp, _int, _i64, _int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                               _vp, _vp, _vp, _vp, _vp]),
    "pssgp_set_fold": (_int, [_vp, _int, _vp, _int, _i64, _vp]),
    "pssgp_pkfs": (_int, [_vp, _int, _i64, _int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "pssgp_pkf_with_summaries": (_int, [_vp, _int, _i64, _int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _int, _int, _vp, _vp,
                                        _vp, _vp, _vp, _vp, _vp, _vp]),
    "pssgp_pkf_backward_summary": (_int, [_vp, _int, _i64, _int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _int,
                                          _vp, _vp]),
    "pssgp_adjoint_fold": (_int, [_vp, _int, _int, _int, _vp, _vp, _vp]),
    "pssgp_shard_forward": (_int, [_vp, _int, _i64, _int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _int, _vp, _vp, _vp, _vp, _vp]),
    "pssgp_rev_fold": (_int, [_vp, _int, _int, _int, _vp, _i64, _vp, _vp]),
    "pssgp_shard_reverse": (_int, [_vp, _int, _i64, _int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _int, _vp,
                                   _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "pssgp_peer_exchange": (_int, [_vp, _vp, _i64, ctypes.POINTER(_vp), ctypes.POINTER(_vp), _int, _int, _i64, _i64, _i64,
                                   _vp, _vp, _vp]),
    "pssgp_merge_queries": (_int, [_vp, _int, _i64, _i64, _vp, _vp, _vp, ctypes.c_double, _vp, _vp, _vp, _vp, _vp]),
    "pssgp_lyap_solve": (_int, [_vp, _vp, _int, _vp]),
    "pssgp_sde_dim": (_int, [_vp, _int, _vp, _vp]),
    "pssgp_sde_batch": (_int, [_vp, _int, _i64, _vp, _i64, _vp, _vp, _vp, _int]),
    "pssgp_sde_batch_jac": (_int, [_vp, _int, _i64, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _int]),
    "pssgp_grid_loglik": (_int, [_vp, _int, _i64, _i64, _int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "pssgp_grid_loglik_grad": (_int, [_vp, _int, _i64, _i64, _int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                      _vp, _vp]),
    "pssgp_kf": (_int, [_vp, _int, _i64, _i64, _int, _int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "pssgp_ks": (_int, [_vp, _int, _i64, _i64, _int, _int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "pssgp_discretise_backward": (_int, [_vp, _int, _i64, _int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
}


class PssgpError(RuntimeError):
    pass


def load_library(path=LIB_PATH):
    if not os.path.exists(path):
        raise ImportError(
            f"pssgp_b200: CUDA library not found at {path}. Build it with "
            f"`python -c 'import __graft_entry__ as g; g.build()'` (nvcc, sm_100a). There is no CPU fallback.")
    lib = ctypes.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    return lib


_lib = None
_lib_lock = threading.Lock()


def lib():
    global _lib
    if _lib is None:
        with _lib_lock:
            if _lib is None:
                _lib = load_library()
    return _lib


def check(rc):
    if rc != 0:
        msg = lib().pssgp_last_error()
        raise PssgpError(f"pssgp_b200 error {rc}: {msg.decode() if msg else '?'}")


class Handle:
    """Owns one pssgp_handle (device workspace) for one CUDA device."""

    def __init__(self, device):
        self.device = int(device)
        self._h = _vp()
        check(lib().pssgp_create(ctypes.byref(self._h), self.device))

    @property
    def ptr(self):
        return self._h

    def set_option(self, name, value):
        check(lib().pssgp_set_option(self._h, name.encode(), int(value)))

    def launch_count(self):
        return int(lib().pssgp_launch_count(self._h))

    def timing_report(self):
        """{kernel name: (launch count, total ms)} since the last report (option "timing" must be 1)."""
        buf = ctypes.create_string_buffer(1 << 16)
        check(lib().pssgp_timing_report(self._h, buf, len(buf)))
        out = {}
        for line in buf.value.decode().splitlines():
            name, cnt, ms = line.split()
            out[name] = (int(cnt), float(ms))
        return out

    def close(self):
        if self._h:
            lib().pssgp_destroy(self._h)
            self._h = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_handles = {}


def handle(device):
    """One handle (device workspace) per (device, host thread): handles are not shared between threads."""
    key = (int(device), threading.get_ident())
    h = _handles.get(key)
    if h is None:
        h = Handle(int(device))
        _handles[key] = h
    return h
