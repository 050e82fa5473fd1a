# -*- coding: utf-8 -*-
"""
ctypes binding of ``include/tricolour_b200.h`` (libtricolour_b200.so).

The library is the hand-written sm_100a CUDA implementation; there is no other
implementation behind this module.  If the shared object has not been built
(``python -c "import __graft_entry__ as g; g.build()"`` or
``make -C tricolour_b200/csrc``) or no GPU is visible, calls raise
``RuntimeError`` -- nothing falls back to the CPU.
"""
import ctypes
import os
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtricolour_b200.so")

TC_OK, TC_ERR_VALUE, TC_ERR_CUDA, TC_ERR_NOGPU = 0, 1, 2, 3
HOST, DEVICE = 0, 1
VIS_COMPLEX64, VIS_FLOAT32 = 0, 1

_vp = ctypes.c_void_p
_i64 = ctypes.c_int64
_i32 = ctypes.c_int32
_int = ctypes.c_int
_dbl = ctypes.c_double


class StParams(ctypes.Structure):
    """mirror of ``tc_st_params``"""
    _fields_ = [
        ("outlier_nsigma", _dbl),
        ("nwin_time", _i32),
        ("nwin_freq", _i32),
        ("windows_time", _vp),
        ("tf_time", _vp),
        ("scale_time", _vp),
        ("windows_freq", _vp),
        ("tf_freq", _vp),
        ("scale_freq", _vp),
        ("background_reject", _dbl),
        ("background_iterations", _i32),
        ("nchunk_ends", _i32),
        ("radii_spec", _vp),
        ("radii_2d", _vp),
        ("freq_chunk_ends", _vp),
        ("time_extend", _i64),
        ("freq_extend", _i64),
        ("average_freq", _i64),
        ("flag_all_time_frac", _dbl),
        ("flag_all_freq_frac", _dbl),
        ("num_major_iterations", _i32),
        ("reserved", _i32),
    ]


_SIGNATURES = {
    "tc_last_error": (ctypes.c_char_p, []),
    "tc_device_count": (_int, []),
    "tc_context_create": (_int, [_int, _vp, ctypes.POINTER(_vp)]),
    "tc_context_destroy": (None, [_vp]),
    "tc_synchronize": (_int, [_vp]),
    "tc_launch_count": (ctypes.c_ulonglong, [_vp]),
    "tc_workspace_peak": (ctypes.c_size_t, [_vp]),
    "tc_workspace_held": (ctypes.c_size_t, [_vp]),
    "tc_workspace_share": (ctypes.c_size_t, [_vp]),
    "tc_context_trim": (_int, [_vp]),
    "tc_profile_enable": (_int, [_vp, _int]),
    "tc_profile_reset": (_int, [_vp]),
    "tc_profile_count": (_int, []),
    "tc_profile_name": (ctypes.c_char_p, [_int]),
    "tc_profile_read": (_int, [_vp, _int, ctypes.POINTER(_dbl), ctypes.POINTER(ctypes.c_longlong)]),
    "tc_alloc_pinned": (_int, [ctypes.c_size_t, ctypes.POINTER(_vp)]),
    "tc_free_pinned": (_int, [_vp]),
    "tc_memcpy_async": (_int, [_vp, _vp, _vp, ctypes.c_size_t, _int]),
    "tc_is_emulated": (_int, []),
    "tc_flag_nans_zeros": (_int, [_vp, _vp, _vp, _vp, _i64, _int]),
    "tc_flag_autos": (_int, [_vp, _vp, _vp, _i64, _i64, _vp, _int]),
    "tc_apply_channel_mask": (_int, [_vp, _vp, _vp, _vp, _int, _i64, _i64, _i64, _vp, _int]),
    "tc_sum_threshold": (_int, [_vp, ctypes.POINTER(StParams), _vp, _int, _vp, _i64, _i64, _i64, _vp, _int]),
    "tc_uvcontsub": (_int, [_vp, _vp, _vp, _i64, _i64, _i64, _int, _int, _int, _dbl, _vp, _int]),
    "tc_polarised_intensity": (_int, [_vp, _vp, _i64, _int, _vp, _vp, _int, _vp, _int]),
    "tc_unpolarised_intensity": (_int, [_vp, _vp, _i64, _int, _vp, _vp, _int, _vp, _vp, _int, _vp, _int]),
    "tc_pack": (_int, [_vp, _vp, _vp, _i64, _vp, _vp, _i64, _i64, _i64, _i64, _vp, _vp, _int, _int]),
    "tc_unpack": (_int, [_vp, _vp, _vp, _i64, _vp, _int, _i64, _i64, _i64, _i64, _vp, _int]),
    "tc_unpack_flags_any_corr": (_int, [_vp, _vp, _vp, _i64, _vp, _i64, _i64, _i64, _i64, _vp, _int]),
    "tc_unpack_flags_broadcast": (_int, [_vp, _vp, _vp, _i64, _vp, _i64, _i64, _i64, _i64, _i64, _vp, _int]),
    "tc_stokes_pack": (_int, [_vp, _vp, _vp, _i64, _vp, _vp, _i64, _i64, _i64, _i64, _vp, _vp, _int, _vp, _vp, _int,
                              _vp, _vp, _int, _int]),
    "tc_window_counts": (_int, [_vp, _vp, _i64, _i64, _i64, _i64, _vp, _vp, _int]),
    "tc_flags_or": (_int, [_vp, _vp, _vp, _vp, _i64, _int]),
    "tc_stage_average_freq": (_int, [_vp, _vp, _int, _vp, _i64, _i64, _i64, _i64, _vp, _vp, _int]),
    "tc_stage_time_median": (_int, [_vp, _vp, _vp, _i64, _i64, _i64, _vp, _vp, _int]),
    "tc_stage_chunk_median_abs": (_int, [_vp, _vp, _vp, _i64, _i64, _i64, _vp, _int, _vp, _int]),
    "tc_stage_masked_filter": (_int, [_vp, _vp, _vp, _i64, _i64, _i64, _i64, _i64, _vp, _int]),
    "tc_stage_interp_nans": (_int, [_vp, _vp, _i64, _i64, _i64, _vp, _int]),
    "tc_stage_background2d": (_int, [_vp, _vp, _vp, _i64, _i64, _i64, _int, _vp, _dbl, _vp, _int, _vp, _int]),
    "tc_stage_sum_threshold": (_int, [_vp, _vp, _vp, _i64, _i64, _i64, _int, _vp, _vp, _vp, _int, _dbl, _vp,
                                      _int, _vp, _int]),
    "tc_stage_combine_unaverage": (_int, [_vp, _vp, _vp, _vp, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _dbl,
                                          _dbl, _vp, _int]),
}

EXPORTED_SYMBOLS = tuple(sorted(_SIGNATURES))

_lib = None
_lib_lock = threading.Lock()


def _bind(lib):
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib


def load(path=None):
    """Load (once) and return the CUDA library.  Fails loudly when missing."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    with _lib_lock:
        if _lib is not None and path is None:
            return _lib
        p = path or LIB_PATH
        if not os.path.exists(p):
            raise RuntimeError(
                "tricolour_b200: CUDA library %s not found. Build it with "
                "`make -C tricolour_b200/csrc` (nvcc, sm_100a). There is no CPU "
                "fallback." % p)
        lib = _bind(ctypes.CDLL(p))
        if path is None:
            _lib = lib
        return lib


def _set_library_for_testing(lib):
    """tests/ only: route the host logic through another build of the same
    C ABI (the CPU-emulated kernels).  Never used by the package itself."""
    global _lib
    _lib = lib
    _tls.__dict__.clear()


def check(rc):
    if rc == TC_OK:
        return
    msg = load().tc_last_error()
    msg = msg.decode("utf-8", "replace") if msg else "unknown error"
    if rc == TC_ERR_VALUE:
        raise ValueError(msg)
    raise RuntimeError("tricolour_b200: " + msg)


class Context(object):
    """Owns a ``tc_context`` (stream + workspace).  One per thread."""

    def __init__(self, device=None, stream=None):
        lib = load()
        if device is None:
            device = default_device()
        h = _vp()
        check(lib.tc_context_create(int(device), _vp(stream) if stream else None,
                                    ctypes.byref(h)))
        self._h = h
        self.device = int(device)
        self.stream = stream

    @property
    def handle(self):
        return self._h

    def synchronize(self):
        check(load().tc_synchronize(self._h))

    def launch_count(self):
        return int(load().tc_launch_count(self._h))

    def workspace_peak(self):
        return int(load().tc_workspace_peak(self._h))

    def workspace_held(self):
        """bytes of device memory this context's arena holds right now"""
        return int(load().tc_workspace_held(self._h))

    def workspace_share(self):
        """this context's share of the device-wide workspace budget (TC_WORKSPACE_MB
        divided between the contexts that hold an arena on the device)"""
        return int(load().tc_workspace_share(self._h))

    def trim(self):
        """hand the arena back to the driver (waits for the context's stream)"""
        check(load().tc_context_trim(self._h))

    def profile(self, on=True):
        check(load().tc_profile_enable(self._h, 1 if on else 0))

    def profile_reset(self):
        check(load().tc_profile_reset(self._h))

    def profile_read(self):
        """{kernel family: (total_ms, launches)} measured with CUDA events"""
        lib = load()
        out = {}
        for i in range(lib.tc_profile_count()):
            ms, n = _dbl(0), ctypes.c_longlong(0)
            check(lib.tc_profile_read(self._h, i, ctypes.byref(ms), ctypes.byref(n)))
            out[lib.tc_profile_name(i).decode()] = (ms.value, n.value)
        return out

    def close(self):
        if getattr(self, "_h", None):
            load().tc_context_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def default_device():
    for key in ("TRICOLOUR_B200_DEVICE", "LOCAL_RANK"):
        v = os.environ.get(key)
        if v is not None and v != "":
            return int(v)
    return 0


_tls = threading.local()
MAX_CACHED_CONTEXTS = 8      # per thread; the least recently used one is closed beyond that


def get_context(device=None, stream=None):
    """Thread-local context for (device, stream); the reference is called from a
    dask ThreadPool, so every worker thread gets its own stream and arena.  A
    thread keeps at most ``MAX_CACHED_CONTEXTS`` contexts (callers that bind work
    to short-lived torch streams would otherwise pile up arenas): the least
    recently used one is closed, which returns its arena to the driver."""
    if device is None:
        device = default_device()
    key = (int(device), int(stream) if stream else 0)
    cache = _tls.__dict__.setdefault("ctx", {})
    ctx = cache.pop(key, None)
    if ctx is None:
        ctx = Context(device, stream)
        while len(cache) >= MAX_CACHED_CONTEXTS:
            old = cache.pop(next(iter(cache)))
            old.close()
    cache[key] = ctx         # most recently used last
    return ctx


def release_contexts(trim_only=False):
    """Close (or, with ``trim_only``, just empty the arenas of) the calling
    thread's cached contexts.  Worker threads that are done flagging call this to
    give their share of the device workspace back."""
    cache = _tls.__dict__.get("ctx", {})
    for key in list(cache):
        if trim_only:
            cache[key].trim()
        else:
            cache.pop(key).close()


# ---------------------------------------------------------------------------
# array plumbing: numpy arrays (host) or torch CUDA tensors (device)
# ---------------------------------------------------------------------------
def is_device_array(x):
    return hasattr(x, "data_ptr") and hasattr(x, "is_cuda") and bool(x.is_cuda)


def ptr(a):
    if a is None:
        return None
    if is_device_array(a):
        return _vp(a.data_ptr())
    return _vp(a.ctypes.data)


CUDA_STREAM_LEGACY = 1   # cudaStreamLegacy: the handle that names torch's default stream


def torch_stream_handle(device_index):
    """cudaStream_t of torch's current stream on ``device_index``.  torch reports
    the default stream as 0; the library treats NULL as "create your own stream",
    so the legacy default stream is passed by its explicit handle instead."""
    import torch
    h = int(torch.cuda.current_stream(device_index).cuda_stream)
    return h if h else CUDA_STREAM_LEGACY


def context_for(*arrays):
    """Host arrays -> thread default context; device tensors -> a context bound
    to torch's current stream on the tensors' device."""
    arrays = [a for a in arrays if a is not None]
    on_dev = [a for a in arrays if is_device_array(a)]
    if not on_dev:
        return get_context(), HOST
    # the kernels take raw pointers: one host array (or a tensor of another GPU) among
    # device tensors would be dereferenced as a device pointer
    if len(on_dev) != len(arrays):
        raise TypeError("tricolour_b200: arrays of one call must all be numpy arrays or all be "
                        "CUDA tensors (got a mix of host and device arrays)")
    import torch
    devs = {a.device.index if a.device.index is not None else torch.cuda.current_device() for a in on_dev}
    if len(devs) != 1:
        raise ValueError("tricolour_b200: the CUDA tensors of one call live on different devices: %s"
                         % sorted(devs))
    dev = devs.pop()
    return get_context(dev, torch_stream_handle(dev)), DEVICE


_pinned = {}


def pinned_empty(shape, dtype):
    """numpy array backed by page-locked host memory (fast H2D/D2H staging).
    The block lives until ``free_pinned(arr)`` or process exit."""
    lib = load()
    dtype = np.dtype(dtype)
    count = int(np.prod(shape))
    n = count * dtype.itemsize
    p = _vp()
    check(lib.tc_alloc_pinned(n, ctypes.byref(p)))
    buf = (ctypes.c_char * max(n, 1)).from_address(p.value)
    arr = np.frombuffer(buf, dtype=dtype, count=count).reshape(shape)
    _pinned[p.value] = buf
    return arr


def free_pinned(arr):
    addr = arr.ctypes.data
    if _pinned.pop(addr, None) is not None:
        check(load().tc_free_pinned(_vp(addr)))
