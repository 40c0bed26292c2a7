"""ctypes binding of include/sgs.h.  The library is built in-tree (csrc/libsgs.so) by
`__graft_entry__.build()` / `make -C csrc`; nothing here falls back to another implementation."""
import ctypes as C
import os
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), 'csrc', 'libsgs.so')

_lib = None
_lock = threading.Lock()
_init_pid = None

c_void_p, c_int, c_int64, c_double_p = C.c_void_p, C.c_int, C.c_int64, C.POINTER(C.c_double)


class SgsError(RuntimeError):
    pass


def _declare(lib):
    def fn(name, restype, *argtypes):
        f = getattr(lib, name)
        f.restype = restype
        f.argtypes = list(argtypes)
        return f

    fn('sgs_abi_version', c_int)
    fn('sgs_last_error', C.c_char_p)
    fn('sgs_init', c_int, c_int)
    fn('sgs_device_count', c_int, C.POINTER(c_int))
    fn('sgs_synchronize', c_int, c_void_p)
    fn('sgs_launch_count', C.c_ulonglong)
    fn('sgs_profile_enable', c_int, c_int)
    fn('sgs_profile_read', c_int, C.c_char_p, C.POINTER(C.c_double), C.POINTER(C.c_ulonglong))
    fn('sgs_feat_plan_create', c_int, C.POINTER(c_void_p), c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int)
    fn('sgs_feat_plan_destroy', None, c_void_p)
    fn('sgs_feat_plan_set_tail', c_int, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p)
    fn('sgs_feat_extract', c_int, c_void_p, c_void_p, c_int, c_int64, c_int, c_int, c_int64, c_void_p, c_int, c_int,
       c_int, c_int64, c_int, c_void_p, c_void_p, c_void_p)
    fn('sgs_feat_stack', c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p)
    c_i32p = c_void_p
    fn('sgs_lda_model_create', c_int, C.POINTER(c_void_p), c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_i32p,
       c_void_p, c_int, c_void_p, c_int)
    fn('sgs_lda_model_destroy', None, c_void_p)
    fn('sgs_lda_decode', c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
       c_int, c_void_p)
    fn('sgs_lda_last_rescored', c_int, c_void_p, C.POINTER(c_int))
    fn('sgs_gl_node_create', c_int, C.POINTER(c_void_p), c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
       c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p, C.c_double, c_int)
    fn('sgs_gl_node_destroy', None, c_void_p)
    fn('sgs_gl_node_synthesize', c_int, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, C.c_uint64, c_void_p,
       c_void_p, c_void_p, c_void_p, c_void_p)
    fn('sgs_feat_stream_create', c_int, C.POINTER(c_void_p), c_void_p, c_int, c_int, c_int, c_int)
    fn('sgs_feat_stream_destroy', None, c_void_p)
    fn('sgs_feat_stream_set_cold_start', c_int, c_void_p, c_int)
    fn('sgs_feat_stream_push', c_int, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p)
    fn('sgs_sos_stream_create', c_int, C.POINTER(c_void_p), c_void_p, c_void_p, c_int, c_int, c_int)
    fn('sgs_sos_stream_destroy', None, c_void_p)
    fn('sgs_sos_stream_push', c_int, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p)
    fn('sgs_gl_node_rebase', c_int, c_void_p, C.c_int32)
    fn('sgs_gl_node_set_log_mels', c_int, c_void_p, c_int)
    fn('sgs_gl_node_push', c_int, c_void_p, c_void_p, c_int, c_void_p, C.c_int32, c_void_p, C.c_uint64, c_void_p,
       C.POINTER(c_int), c_void_p)
    fn('sgs_chain_create', c_int, C.POINTER(c_void_p), c_void_p, c_int, c_void_p, c_void_p)
    fn('sgs_chain_destroy', None, c_void_p)
    fn('sgs_chain_push', c_int, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p, C.c_int32, c_void_p,
       C.c_uint64, c_void_p, c_void_p, c_void_p, c_void_p, C.POINTER(c_int), c_void_p)
    fn('sgs_gl_batch_create', c_int, C.POINTER(c_void_p), c_int, c_int, c_int, c_void_p, c_void_p, c_void_p)
    fn('sgs_gl_batch_destroy', None, c_void_p)
    fn('sgs_gl_batch_synthesize', c_int, c_void_p, c_void_p, c_int, c_int, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p)
    fn('sgs_logmel', c_int, c_void_p, c_int64, c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_int64, c_void_p, c_void_p)
    fn('sgs_col_minmax', c_int, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p)
    fn('sgs_quantize', c_int, c_void_p, c_int64, c_int, c_void_p, c_int, c_void_p, c_void_p)
    fn('sgs_spearman', c_int, c_void_p, c_int64, c_int, c_int64, c_void_p, c_int, c_void_p, c_void_p, c_void_p)
    fn('sgs_col_means', c_int, c_void_p, c_int64, c_int64, c_void_p, c_int, c_void_p, c_void_p)
    fn('sgs_lda_stats', c_int, c_void_p, c_int64, c_int64, c_void_p, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p,
       c_void_p, c_void_p, c_void_p, c_void_p)
    fn('sgs_exp_angle', c_int, c_void_p, c_void_p, c_int64, c_void_p, c_void_p)
    fn('sgs_decimate', c_int, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p)
    fn('sgs_dequantize', c_int, c_void_p, c_int, c_int, c_void_p, c_int, c_void_p, c_int64, c_int, c_void_p, c_void_p)
    return lib


def lib():
    """The loaded library (no device needed to load it or to look up symbols)."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise SgsError("libsgs.so not built (%s): run `python -c 'import __graft_entry__ as g; g.build()'` "
                                   "or `make -C closed-loop-seeg-speech-synthesis_b200/csrc`; there is no CPU fallback" % LIB_PATH)
                _lib = _declare(C.CDLL(LIB_PATH))
    return _lib


def check(rc):
    if rc != 0:
        raise SgsError(lib().sgs_last_error().decode('utf-8', 'replace') or ('libsgs error %d' % rc))


def ensure_init(device=None):
    """Creates the CUDA context lazily and per process: the reference forks the whole node graph into a
    child (Sender.py:57-64), so nothing may touch CUDA before the first add_data in that child."""
    global _init_pid
    pid = os.getpid()
    if _init_pid != pid:
        if device is None:
            device = int(os.environ.get('SGS_DEVICE', os.environ.get('LOCAL_RANK', '0')))
        check(lib().sgs_init(int(device)))
        _init_pid = pid


def launch_count():
    return int(lib().sgs_launch_count())


def profile_enable(on=True):
    check(lib().sgs_profile_enable(int(on)))


def profile_read(name):
    """(total device milliseconds, launches) of one kernel class since profile_enable(True)."""
    ms, n = C.c_double(0), C.c_ulonglong(0)
    check(lib().sgs_profile_read(name.encode(), C.byref(ms), C.byref(n)))
    return ms.value, int(n.value)


# ---- pointer helpers --------------------------------------------------------------------------
def _is_torch(a):
    return type(a).__module__.startswith('torch')


def ptr(a):
    """Raw address of a numpy array (host) or a torch tensor (host or device)."""
    if a is None:
        return None
    if _is_torch(a):
        assert a.is_contiguous()
        return c_void_p(a.data_ptr())
    assert a.flags['C_CONTIGUOUS']
    return c_void_p(a.ctypes.data)


def host(a, dtype):
    """Small host-side table as a contiguous numpy array of `dtype`."""
    return np.ascontiguousarray(a, dtype=dtype)


def current_stream(like=None):
    """cudaStream_t to launch on: torch's current stream when the data is a CUDA tensor, else the default stream."""
    if like is not None and _is_torch(like) and like.is_cuda:
        import torch
        return c_void_p(torch.cuda.current_stream(like.device).cuda_stream)
    return c_void_p(0)
