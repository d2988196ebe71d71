"""Device-level wrappers over the C ABI (torch tensors own the memory).

These are the calls ``bench.py`` times with inputs resident in HBM; the
reference-facing classes in ``utils`` / ``pipeline`` / ``networks.unet`` use the
``*_host`` entry points (host buffers in, host buffers out).
"""
import ctypes

import numpy as np

from . import _lib


def _torch():
    import torch
    return torch


class Workspace(object):
    """Grow-only device scratch buffer (allocated through torch)."""

    def __init__(self, device=None):
        self.buf = None
        self.device = device

    def get(self, nbytes):
        torch = _torch()
        if self.buf is None or self.buf.numel() < nbytes:
            self.buf = None
            self.buf = torch.empty(max(int(nbytes), 256), dtype=torch.uint8,
                                   device=self.device or 'cuda')
        return self.buf


_default_ws = {}


def _ws(key):
    """Default scratch buffer of (purpose, device index): allocated ON that device, whatever the
    current device is."""
    if key not in _default_ws:
        _default_ws[key] = Workspace(device='cuda:%d' % key[1])
    return _default_ws[key]


def label_centroids(mask, max_rows=4096, frame0=0, want_labels=False, workspace=None):
    """mask: uint8 cuda tensor (N,H,W) or (N,D,H,W) (already in utils.py:519's swapped
    order for volumes).  Returns (table (N,max_rows,5) f32, counts (N) i32[, labels i32])."""
    torch = _torch()
    lib = _lib.load()
    assert mask.is_cuda and mask.dtype == torch.uint8 and mask.is_contiguous()
    if mask.dim() == 3:
        n, (d, h, w) = mask.shape[0], (1,) + tuple(mask.shape[1:])
    elif mask.dim() == 4:
        n, d, h, w = mask.shape
    else:
        raise ValueError("Incorrect image data shape.")
    hd = _lib.handle(mask.device.index)
    need = ctypes.c_size_t()
    _lib.check(lib.sq_label_workspace_bytes(hd, n, d, h, w, max_rows, ctypes.byref(need)))
    ws = (workspace or _ws(('label', mask.device.index))).get(need.value)
    table = torch.empty((n, max_rows, 5), dtype=torch.float32, device=mask.device)
    counts = torch.empty((n,), dtype=torch.int32, device=mask.device)
    labels = torch.empty(mask.shape, dtype=torch.int32, device=mask.device) if want_labels else None
    _lib.check(lib.sq_label_centroids(hd, mask.data_ptr(), n, d, h, w, frame0, _lib.ptr(labels),
                                      table.data_ptr(), counts.data_ptr(), max_rows,
                                      ws.data_ptr(), ws.numel(), _lib.stream_ptr(mask.device)))
    return (table, counts, labels) if want_labels else (table, counts)


def weightmap_edt(mask, w0=10., sigma=5., out_dtype='float32', want_d2=False, workspace=None):
    """W1 on a uint8 cuda tensor (N,H,W) -> weights (N,H,W) [, exact squared distances int32]."""
    torch = _torch()
    lib = _lib.load()
    assert mask.is_cuda and mask.dtype == torch.uint8 and mask.is_contiguous() and mask.dim() == 3
    n, h, w = mask.shape
    hd = _lib.handle(mask.device.index)
    need = ctypes.c_size_t()
    _lib.check(lib.sq_weightmap_workspace_bytes(hd, n, h, w, 0, ctypes.byref(need)))
    ws = (workspace or _ws(('wm', mask.device.index))).get(need.value)
    tdt = torch.float32 if out_dtype == 'float32' else torch.float64
    out = torch.empty((n, h, w), dtype=tdt, device=mask.device)
    d2 = torch.empty((n, h, w), dtype=torch.int32, device=mask.device) if want_d2 else None
    _lib.check(lib.sq_weightmap_edt(hd, mask.data_ptr(), n, h, w, float(w0), float(sigma),
                                    _lib.F32 if out_dtype == 'float32' else _lib.F64,
                                    out.data_ptr(), _lib.ptr(d2), ws.data_ptr(), ws.numel(),
                                    _lib.stream_ptr(mask.device)))
    return (out, d2) if want_d2 else out


def weightmap_unet(labels, w0=10., sigma=5., wc=None, out_dtype='float32', workspace=None):
    """W3 on an int32 cuda tensor of instance labels (N,H,W) -> weights (N,H,W)."""
    torch = _torch()
    lib = _lib.load()
    assert labels.is_cuda and labels.dtype == torch.int32 and labels.is_contiguous() and labels.dim() == 3
    n, h, w = labels.shape
    hd = _lib.handle(labels.device.index)
    need = ctypes.c_size_t()
    _lib.check(lib.sq_weightmap_workspace_bytes(hd, n, h, w, 1, ctypes.byref(need)))
    ws = (workspace or _ws(('wm', labels.device.index))).get(need.value)
    tdt = torch.float32 if out_dtype == 'float32' else torch.float64
    out = torch.empty((n, h, w), dtype=tdt, device=labels.device)
    wc_arr = None if wc is None else (ctypes.c_double * 2)(float(wc[0]), float(wc[1]))
    _lib.check(lib.sq_weightmap_unet(hd, labels.data_ptr(), n, h, w, float(w0), float(sigma),
                                     None if wc_arr is None else ctypes.cast(wc_arr, ctypes.c_void_p),
                                     _lib.F32 if out_dtype == 'float32' else _lib.F64,
                                     out.data_ptr(), ws.data_ptr(), ws.numel(), _lib.stream_ptr(labels.device)))
    return out


def weighted_cross_entropy(logits, labels, weights, want_grad=True, workspace=None):
    """Weighted softmax cross-entropy of the head on cuda tensors: logits float32 (...,K), labels
    uint8 (...), weights float32 (...) (e.g. the GPU weight map).  Returns (loss float64 0-d tensor,
    grad float32 like logits or None)."""
    torch = _torch()
    lib = _lib.load()
    k = logits.shape[-1]
    npix = logits.numel() // k
    assert logits.is_cuda and logits.dtype == torch.float32 and logits.is_contiguous()
    assert labels.dtype == torch.uint8 and labels.numel() == npix and labels.is_contiguous()
    assert weights.dtype == torch.float32 and weights.numel() == npix and weights.is_contiguous()
    hd = _lib.handle(logits.device.index)
    need = ctypes.c_size_t()
    _lib.check(lib.sq_weighted_ce_workspace_bytes(hd, ctypes.byref(need)))
    ws = (workspace or _ws(('ce', logits.device.index))).get(need.value)
    loss = torch.empty((), dtype=torch.float64, device=logits.device)
    grad = torch.empty_like(logits) if want_grad else None
    _lib.check(lib.sq_weighted_ce(hd, logits.data_ptr(), labels.data_ptr(), weights.data_ptr(), npix, k,
                                  loss.data_ptr(), _lib.ptr(grad), ws.data_ptr(), ws.numel(),
                                  _lib.stream_ptr(logits.device)))
    return loss, grad


def rotation_transform(theta, height, width):
    """TensorFlow 1.x ``angles_to_projective_transforms`` for one angle (float32 arithmetic):
    (cos, -sin, x_off, sin, cos, y_off) maps an output pixel (x, y) to its input sample point, a
    rotation about the image centre as ``tf.contrib.image.rotate`` (reference networks/unet.py:373)."""
    f32 = np.float32
    theta = f32(theta)
    c, s = np.cos(theta, dtype=f32), np.sin(theta, dtype=f32)
    wm1, hm1 = f32(width) - f32(1), f32(height) - f32(1)
    x_off = (wm1 - (c * wm1 - s * hm1)) / f32(2)
    y_off = (hm1 - (s * wm1 + c * hm1)) / f32(2)
    return np.array([c, -s, x_off, s, c, y_off], dtype=f32)


def tr_augment(image, label, weights, transforms, crops, ch, cw, num_outputs=2):
    """Fused rotate + crop + one-hot of the training pipeline (reference networks/unet.py:348-401)
    on cuda tensors: image float32 (N,H,W,C), label uint8 (N,H,W), weights float32 (N,H,W);
    transforms float32 ndarray (N,6) (see ``rotation_transform``), crops int ndarray (N,2) = (rh, rw).
    Returns (image (N,ch,cw,C) float32, label (N,ch,cw,num_outputs) uint8, weights (N,ch,cw) float32)."""
    torch = _torch()
    lib = _lib.load()
    assert image.is_cuda and image.dtype == torch.float32 and image.is_contiguous() and image.dim() == 4
    n, h, w, c = (int(v) for v in image.shape)
    assert label.dtype == torch.uint8 and tuple(label.shape) == (n, h, w) and label.is_contiguous()
    assert weights.dtype == torch.float32 and tuple(weights.shape) == (n, h, w) and weights.is_contiguous()
    transforms = np.ascontiguousarray(transforms, dtype=np.float32).reshape(n, 6)
    crops = np.ascontiguousarray(crops, dtype=np.int32).reshape(n, 2)
    img_o = torch.empty((n, ch, cw, c), dtype=torch.float32, device=image.device)
    lab_o = torch.empty((n, ch, cw, num_outputs), dtype=torch.uint8, device=image.device)
    wgt_o = torch.empty((n, ch, cw), dtype=torch.float32, device=image.device)
    _lib.check(lib.sq_tr_augment(_lib.handle(image.device.index), image.data_ptr(), label.data_ptr(),
                                 weights.data_ptr(), n, h, w, c, transforms.ctypes.data, crops.ctypes.data,
                                 int(ch), int(cw), int(num_outputs), img_o.data_ptr(), lab_o.data_ptr(),
                                 wgt_o.data_ptr(), _lib.stream_ptr(image.device)))
    return img_o, lab_o, wgt_o


def _stack_geometry(x):
    torch = _torch()
    assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous() and x.dim() == 4
    return tuple(int(v) for v in x.shape)


def image_norm(x, out=None, workspace=None):
    """ImageNorm on a float32 cuda stack (N,H,W,C): (x - mean) / std per image and channel."""
    torch = _torch()
    lib = _lib.load()
    n, h, w, c = _stack_geometry(x)
    hd = _lib.handle(x.device.index)
    need = ctypes.c_size_t()
    _lib.check(lib.sq_prep_workspace_bytes(hd, n, c, ctypes.byref(need)))
    ws = (workspace or _ws(('prep', x.device.index))).get(need.value)
    out = torch.empty_like(x) if out is None else out
    _lib.check(lib.sq_image_norm(hd, x.data_ptr(), out.data_ptr(), n, h, w, c, ws.data_ptr(), ws.numel(),
                                 _lib.stream_ptr(x.device)))
    return out


def image_outliers(x, size=2, threshold=5.):
    """ImageOutliers on a float32 cuda stack (N,H,W,C): hot pixels replaced by the local median."""
    torch = _torch()
    lib = _lib.load()
    n, h, w, c = _stack_geometry(x)
    out = torch.empty_like(x)
    _lib.check(lib.sq_image_outliers(_lib.handle(x.device.index), x.data_ptr(), out.data_ptr(), n, h, w, c,
                                     int(size), float(threshold), _lib.stream_ptr(x.device)))
    return out


def image_bgsubtract(x, out_dtype='float32', workspace=None):
    """ImageBGSubtract on a float32 cuda stack (N,H,W,1): minus the least-squares quadratic surface."""
    torch = _torch()
    lib = _lib.load()
    n, h, w, c = _stack_geometry(x)
    if c != 1:
        raise ValueError('image_bgsubtract takes single-channel stacks (N,H,W,1)')
    hd = _lib.handle(x.device.index)
    need = ctypes.c_size_t()
    _lib.check(lib.sq_prep_workspace_bytes(hd, n, c, ctypes.byref(need)))
    ws = (workspace or _ws(('prep', x.device.index))).get(need.value)
    out = torch.empty(x.shape, dtype=torch.float32 if out_dtype == 'float32' else torch.float64, device=x.device)
    _lib.check(lib.sq_image_bgsubtract(hd, x.data_ptr(), out.data_ptr(),
                                       _lib.F32 if out_dtype == 'float32' else _lib.F64, n, h, w,
                                       ws.data_ptr(), ws.numel(), _lib.stream_ptr(x.device)))
    return out


# ------------------------------------------------------------- host-buffer calls

def image_pipe_host(which, image, size=2, threshold=5., out_dtype='float32', device=None):
    """which in ('norm', 'outliers', 'bgsubtract'); image float32 ndarray (H,W,C) or (N,H,W,C)."""
    lib = _lib.load()
    image = np.ascontiguousarray(image, dtype=np.float32)
    squeeze = image.ndim == 3
    if squeeze:
        image = image[None]
    if image.ndim != 4:
        raise ValueError('image_pipe_host: (H,W,C) or (N,H,W,C) float32 images')
    n, h, w, c = image.shape
    code = {'norm': 0, 'outliers': 1, 'bgsubtract': 2}[which]
    f64 = (code == 2 and out_dtype == 'float64')
    out = np.empty(image.shape, dtype=np.float64 if f64 else np.float32)
    _lib.check(lib.sq_image_pipe_host(_lib.handle(device), code, image.ctypes.data, out.ctypes.data,
                                      _lib.F64 if f64 else _lib.F32, n, h, w, c, int(size), float(threshold)))
    return out[0] if squeeze else out



def label_centroids_host(mask, max_rows=4096, frame0=0, want_labels=False, device=None):
    """mask: uint8 ndarray (N,H,W) / (N,D,H,W).  Returns list of per-frame (n_i,5) float32
    tables [, labels int32 ndarray].  Retries with a larger table on overflow."""
    lib = _lib.load()
    mask = np.ascontiguousarray(mask, dtype=np.uint8)
    if mask.ndim == 3:
        n, (d, h, w) = mask.shape[0], (1,) + mask.shape[1:]
    elif mask.ndim == 4:
        n, d, h, w = mask.shape
    else:
        raise ValueError("Incorrect image data shape.")
    hd = _lib.handle(device)
    labels = np.empty(mask.shape, dtype=np.int32) if want_labels else None
    while True:
        table = np.empty((n, max_rows, 5), dtype=np.float32)
        counts = np.empty((n,), dtype=np.int32)
        st = lib.sq_label_centroids_host(hd, mask.ctypes.data, n, d, h, w, frame0,
                                         _lib.ptr(labels), table.ctypes.data, counts.ctypes.data,
                                         max_rows)
        if st == _lib.SQ_EOVERFLOW:
            max_rows = int(counts.max())
            continue
        _lib.check(st)
        break
    tables = [table[i, :counts[i]].copy() for i in range(n)]
    return (tables, labels) if want_labels else tables


def weightmap_edt_host(mask, w0=10., sigma=5., out_dtype='float64', want_d2=False, device=None):
    lib = _lib.load()
    mask = np.ascontiguousarray(mask)
    if mask.dtype != np.uint8:
        mask = (mask != 0).astype(np.uint8)
    squeeze = mask.ndim == 2
    if squeeze:
        mask = mask[None]
    n, h, w = mask.shape
    out = np.empty((n, h, w), dtype=np.float32 if out_dtype == 'float32' else np.float64)
    d2 = np.empty((n, h, w), dtype=np.int32) if want_d2 else None
    _lib.check(lib.sq_weightmap_edt_host(_lib.handle(device), mask.ctypes.data, n, h, w,
                                         float(w0), float(sigma),
                                         _lib.F32 if out_dtype == 'float32' else _lib.F64,
                                         out.ctypes.data, _lib.ptr(d2)))
    if squeeze:
        out = out[0]
        d2 = d2[0] if want_d2 else None
    return (out, d2) if want_d2 else out


def weightmap_unet_host(labels, w0=10., sigma=5., wc=None, out_dtype='float64', device=None):
    lib = _lib.load()
    labels = np.ascontiguousarray(labels, dtype=np.int32)
    squeeze = labels.ndim == 2
    if squeeze:
        labels = labels[None]
    n, h, w = labels.shape
    out = np.empty((n, h, w), dtype=np.float32 if out_dtype == 'float32' else np.float64)
    wc_arr = None if wc is None else (ctypes.c_double * 2)(float(wc[0]), float(wc[1]))
    _lib.check(lib.sq_weightmap_unet_host(_lib.handle(device), labels.ctypes.data, n, h, w,
                                          float(w0), float(sigma),
                                          None if wc_arr is None else ctypes.cast(wc_arr, ctypes.c_void_p),
                                          _lib.F32 if out_dtype == 'float32' else _lib.F64,
                                          out.ctypes.data))
    return out[0] if squeeze else out
