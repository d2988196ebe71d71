"""ctypes binding of ``libsequitr_b200.so`` (the C ABI in ``include/sequitr_b200.h``).

There is NO CPU fallback: if the shared library is missing, or no sm_100 device
is present, every product entry point raises.  torch is used only to own device
buffers and streams.
"""
import ctypes
import os
import re

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libsequitr_b200.so')
HEADER_PATH = os.path.join(os.path.dirname(_HERE), 'include', 'sequitr_b200.h')

SQ_OK, SQ_EINVAL, SQ_ECUDA, SQ_ENOMEM, SQ_EOVERFLOW, SQ_EUNSUPPORTED, SQ_ESTATE = 0, -1, -2, -3, -4, -5, -6
BRIDGE_CODES = {None: 0, 'eltwise_add': 1, 'eltwise_mul': 2, 'eltwise_sub': 3, 'concat': 4}
MODE_FP32_EXACT, MODE_BF16_TC = 0, 1
F32, F64, U8, U16 = 0, 1, 2, 3

c_int, c_void_p, c_size_t, c_double, c_char_p = (ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t,
                                                 ctypes.c_double, ctypes.c_char_p)
_P = ctypes.POINTER

_SIGNATURES = {
    'sq_version': (c_char_p, []),
    'sq_last_error': (c_char_p, []),
    'sq_create': (c_int, [c_int, _P(c_void_p)]),
    'sq_destroy': (c_int, [c_void_p]),
    'sq_device_info': (c_int, [c_void_p, _P(c_int), _P(c_int), _P(c_int), _P(c_size_t)]),
    'sq_host_register': (c_int, [c_void_p, c_size_t]),
    'sq_host_unregister': (c_int, [c_void_p]),
    'sq_label_workspace_bytes': (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, _P(c_size_t)]),
    'sq_label_centroids': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p,
                                   c_void_p, c_void_p, c_int, c_void_p, c_size_t, c_void_p]),
    'sq_label_centroids_host': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                        c_void_p, c_void_p, c_void_p, c_int]),
    'sq_weightmap_workspace_bytes': (c_int, [c_void_p, c_int, c_int, c_int, c_int, _P(c_size_t)]),
    'sq_weightmap_edt': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_double, c_double, c_int,
                                 c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    'sq_weightmap_unet': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_double, c_double,
                                  c_void_p, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    'sq_weightmap_edt_host': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_double, c_double,
                                      c_int, c_void_p, c_void_p]),
    'sq_weightmap_unet_host': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_double, c_double,
                                       c_void_p, c_int, c_void_p]),
    'sq_unet_create': (c_int, [c_void_p, c_int, c_int, c_int, _P(c_int), c_int, c_int, c_int,
                               _P(c_void_p)]),
    'sq_unet_destroy': (c_int, [c_void_p]),
    'sq_unet_load_weights': (c_int, [c_void_p, c_char_p, c_void_p, _P(ctypes.c_int64), c_int]),
    'sq_unet_finalize': (c_int, [c_void_p]),
    'sq_unet_workspace_bytes': (c_int, [c_void_p, c_int, c_int, c_int, c_int, _P(c_size_t)]),
    'sq_unet_forward': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                                c_void_p, c_void_p, c_size_t, c_void_p]),
    'sq_unet_last_launches': (c_int, [c_void_p, _P(c_int)]),
    'sq_unet_profile': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_size_t,
                                c_void_p, _P(c_char_p), _P(ctypes.c_float), _P(c_double), c_int,
                                _P(c_int)]),
    'sq_weighted_ce_workspace_bytes': (c_int, [c_void_p, _P(c_size_t)]),
    'sq_weighted_ce': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, ctypes.c_longlong, c_int, c_void_p,
                       c_void_p, c_void_p, c_size_t, c_void_p]),
    'sq_trainer_create': (c_int, [c_void_p, c_int, ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_float,
                                  ctypes.c_float, ctypes.c_ulonglong, _P(c_void_p)]),
    'sq_trainer_destroy': (c_int, [c_void_p]),
    'sq_trainer_workspace_bytes': (c_int, [c_void_p, c_int, c_int, c_int, c_int, _P(c_size_t)]),
    'sq_trainer_step': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p,
                                c_void_p, c_size_t, c_void_p]),
    'sq_trainer_apply': (c_int, [c_void_p, c_void_p]),
    'sq_trainer_grad_arena': (c_int, [c_void_p, _P(c_void_p), _P(c_size_t)]),
    'sq_trainer_read': (c_int, [c_void_p, c_char_p, c_int, c_void_p, c_size_t]),
    'sq_tr_augment': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p,
                              c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    'sq_prep_workspace_bytes': (c_int, [c_void_p, c_int, c_int, _P(c_size_t)]),
    'sq_image_norm': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_size_t,
                              c_void_p]),
    'sq_image_norm_raw': (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p,
                                  c_size_t, c_void_p]),
    'sq_image_outliers': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_double,
                                  c_void_p]),
    'sq_image_bgsubtract': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p,
                                    c_size_t, c_void_p]),
    'sq_image_pipe_host': (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                   c_int, c_double]),
    'sq_segment_localise_raw_host': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p,
                                             c_void_p, c_int, c_void_p]),
    'sq_image_cast': (c_int, [c_void_p, c_void_p, c_int, c_void_p, ctypes.c_longlong, c_void_p]),
    'sq_segment_localise_host': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p,
                                         c_void_p, c_int, c_void_p]),
}


def declared_symbols():
    """Names of every function include/sequitr_b200.h declares."""
    with open(HEADER_PATH) as f:
        text = f.read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(sq_[a-z0-9_]+)\s*\(', text)))


class SequitrError(RuntimeError):
    pass


_lib = None


def load():
    """Load the shared library (no GPU needed for this step)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "sequitr_b200: %s is missing -- build it with "
            "`python -c 'import __graft_entry__ as g; g.build()'` (there is no CPU fallback)"
            % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status):
    if status == SQ_OK:
        return
    msg = load().sq_last_error().decode('utf-8', 'replace')
    if status == SQ_EINVAL:
        raise ValueError(msg)
    if status == SQ_ENOMEM:
        raise MemoryError(msg)
    if status == SQ_EOVERFLOW:
        raise OverflowError(msg)
    if status == SQ_EUNSUPPORTED:
        raise NotImplementedError(msg)
    raise SequitrError('%s (status %d)' % (msg, status))


_handles = {}


def handle(device=None):
    """Per-device library handle (creates it on first use; raises without a B200)."""
    import torch
    if not torch.cuda.is_available():
        raise SequitrError("sequitr_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    if device is None:
        device = torch.cuda.current_device()
    device = int(device)
    if device not in _handles:
        h = c_void_p()
        check(load().sq_create(device, ctypes.byref(h)))
        _handles[device] = h
    return _handles[device]


def new_handle(device):
    """A library handle of its own (own streams, own device arena) on `device`: host calls through plans built on
    different handles can run concurrently.  The caller destroys it (sq_destroy)."""
    handle(device)                                  # same availability checks as the shared handle
    h = c_void_p()
    check(load().sq_create(int(device), ctypes.byref(h)))
    return h


def ptr(t):
    """Raw pointer of a torch tensor / numpy array / None."""
    if t is None:
        return None
    if isinstance(t, np.ndarray):
        return t.ctypes.data
    return t.data_ptr()


def stream_ptr(device=None):
    """Raw handle of torch's current stream ON `device` (a tensor's device / index / None = the
    current device): kernels must run on a stream of the GPU that owns their buffers."""
    import torch
    return torch.cuda.current_stream(device).cuda_stream
