"""Driver for the plain-C UNet oracle (TEST INFRASTRUCTURE ONLY).

Walks the reference topology -- ``UNet.build`` (/root/reference/sequitr/
networks/unet.py:224-262), ``conv_block`` (:265-277), ``down_layer`` (:282-296),
``up_layer`` (:299-322), bridges (:182-202) -- calling the layer primitives of
``oracle/unet_ref.c``.  Dropout (:274-276) is the identity at inference.

``contract='fp32'`` is the bit-exact contract of the GPU "fp32 exact" mode.
``contract='bf16'`` additionally rounds the input, every kernel and every stored
activation to bf16 (fp32 accumulation) -- the storage points of the tensor-core
path -- so that path can be compared at a tolerance that only has to absorb the
accumulation order.
"""
import ctypes
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, '_build', 'libsqref_unet.so')
_lib = None
_pool = None


def build(force=False):
    if force or not os.path.exists(_SO) or \
            os.path.getmtime(_SO) < os.path.getmtime(os.path.join(_HERE, 'unet_ref.c')):
        subprocess.check_call(['make', '-C', _HERE, '-s'])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
    return _lib


def num_threads():
    return int(os.environ.get('SQREF_THREADS', os.cpu_count() or 1))


def _run_rows(fn, args, nrows):
    """Call fn(*args, r0, r1) over row chunks on a thread pool."""
    global _pool
    nt = max(1, min(num_threads(), nrows))
    if nt == 1:
        fn(*args, ctypes.c_long(0), ctypes.c_long(nrows))
        return
    if _pool is None:
        _pool = ThreadPoolExecutor(max_workers=num_threads())
    chunks = nt * 4
    futs = []
    for t in range(chunks):
        r0, r1 = nrows * t // chunks, nrows * (t + 1) // chunks
        if r1 > r0:
            futs.append(_pool.submit(fn, *args, ctypes.c_long(r0), ctypes.c_long(r1)))
    for f in futs:
        f.result()


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def round_bf16(a):
    a = _f32(a)
    out = np.empty_like(a)
    _run_rows(lib().sqref_round_bf16, (_p(a), _p(out)), a.size)
    return out


def conv(x0, x1, kernel, scale, shift, relu):
    """SAME conv over concat(x0, x1) (x1 may be None); NHWC or NDHWC."""
    x0 = _f32(x0)
    kernel = _f32(kernel)
    nd = x0.ndim - 2
    c0 = x0.shape[-1]
    c1 = 0 if x1 is None else x1.shape[-1]
    if x1 is not None:
        x1 = _f32(x1)
    co = kernel.shape[-1]
    assert kernel.shape[-2] == c0 + c1, (kernel.shape, c0, c1)
    out = np.empty(x0.shape[:-1] + (co,), dtype=np.float32)
    scale, shift = _f32(scale), _f32(shift)
    p1 = _p(x1) if x1 is not None else None
    if nd == 2:
        n, h, w = x0.shape[:3]
        args = (_p(x0), c0, p1, c1, h, w, _p(kernel), kernel.shape[0], kernel.shape[1], co,
                _p(scale), _p(shift), int(relu), _p(out))
        _run_rows(lib().sqref_conv2d, args, n * h)
    else:
        n, d, h, w = x0.shape[:4]
        args = (_p(x0), c0, p1, c1, d, h, w, _p(kernel), kernel.shape[0], kernel.shape[1],
                kernel.shape[2], co, _p(scale), _p(shift), int(relu), _p(out))
        _run_rows(lib().sqref_conv3d, args, n * d * h)
    return out


def maxpool(x):
    x = _f32(x)
    if x.ndim == 4:
        n, h, w, c = x.shape
        out = np.empty((n, h // 2, w // 2, c), dtype=np.float32)
        _run_rows(lib().sqref_maxpool2d, (_p(x), h, w, c, _p(out)), n * (h // 2))
    else:
        n, d, h, w, c = x.shape
        out = np.empty((n, d // 2, h // 2, w // 2, c), dtype=np.float32)
        _run_rows(lib().sqref_maxpool3d, (_p(x), d, h, w, c, _p(out)), n * (d // 2) * (h // 2))
    return out


def upconv(x, kernel, bias):
    x, kernel, bias = _f32(x), _f32(kernel), _f32(bias)
    co, ci = kernel.shape[-2], kernel.shape[-1]
    assert ci == x.shape[-1]
    if x.ndim == 4:
        n, h, w, _ = x.shape
        out = np.empty((n, 2 * h, 2 * w, co), dtype=np.float32)
        _run_rows(lib().sqref_upconv2d, (_p(x), h, w, ci, _p(kernel), _p(bias), co, _p(out)), n * h)
    else:
        n, d, h, w, _ = x.shape
        out = np.empty((n, 2 * d, 2 * h, 2 * w, co), dtype=np.float32)
        _run_rows(lib().sqref_upconv3d, (_p(x), d, h, w, ci, _p(kernel), _p(bias), co, _p(out)),
                  n * d * h)
    return out


def eltwise(a, b, op):
    a, b = _f32(a), _f32(b)
    out = np.empty_like(a)
    _run_rows(lib().sqref_eltwise, (_p(a), _p(b), {'eltwise_add': 0, 'eltwise_mul': 1,
                                                  'eltwise_sub': 2}[op], _p(out)), a.size)
    return out


def softmax_argmax(logits):
    logits = _f32(logits)
    k = logits.shape[-1]
    npix = logits.size // k
    probs = np.empty_like(logits)
    mask = np.empty(logits.shape[:-1], dtype=np.uint8)
    _run_rows(lib().sqref_softmax_argmax, (_p(logits), k, _p(probs), _p(mask)), npix)
    return probs, mask


def _affine(weights, scope):
    """(scale, shift) of the conv epilogue y = relu(acc*scale + shift)."""
    bias = _f32(weights[scope + '/bias'])
    if scope + '/scale' in weights:
        s = _f32(weights[scope + '/scale'])
        # folded in fp32 exactly as the product loader does: shift' = bias*scale + shift
        return s, (bias * s + _f32(weights[scope + '/shift'])).astype(np.float32)
    return np.ones_like(bias), bias


def unet_forward(x, weights, filters, bridge='concat', contract='fp32', return_all=False, exact_tail=False):
    """x: (N,H,W,Cin) or (N,D,H,W,Cin) float32.  Returns dict with 'logits',
    'probs', 'mask' (and 'net', the per-layer outputs, if return_all).
    ``exact_tail`` (experiment, contract 'bf16' only): the last block (up0/conv1, up0/conv2) and the head run
    in full fp32 on unrounded weights -- the limit of a tf32 / 3xbf16-split variant of those layers -- while
    everything before keeps the bf16 contract (scripts/exact_tail_experiment.py)."""
    rb = round_bf16 if contract == 'bf16' else (lambda a: _f32(a))
    exact = (lambda a: _f32(a))
    tail = ('UNet/up0/conv1/kernel', 'UNet/up0/conv2/kernel', 'UNet/to_image/kernel') if exact_tail else ()
    W = {k: (rb(v) if k.endswith('/kernel') and k not in tail else _f32(v)) for k, v in weights.items()}

    def block(x0, x1, scope):
        r = exact if (exact_tail and scope == 'UNet/up0') else rb
        s, t = _affine(W, scope + '/conv1')
        y = r(conv(x0, x1, W[scope + '/conv1/kernel'], s, t, True))
        s, t = _affine(W, scope + '/conv2')
        return r(conv(y, None, W[scope + '/conv2/kernel'], s, t, True))

    x = rb(x)
    net = [block(x, None, 'UNet/down0')]
    for i in range(1, len(filters)):
        net.append(block(maxpool(net[-1]), None, 'UNet/down%d' % i))
    for i in reversed(range(len(filters) - 1)):
        scope = 'UNet/up%d' % i
        up = rb(upconv(net[-1], W[scope + '/upscale/kernel'], W[scope + '/upscale/bias']))
        skip = net[i]
        if bridge == 'concat':
            net.append(block(up, skip, scope))
        elif bridge is None:
            net.append(block(up, None, scope))
        else:
            net.append(block(rb(eltwise(up, skip, bridge)), None, scope))
    k = W['UNet/to_image/kernel']
    logits = conv(net[-1], None, k, np.ones(k.shape[-1], np.float32),
                  W['UNet/to_image/bias'], False)
    probs, mask = softmax_argmax(logits)
    res = {'logits': logits, 'probs': probs, 'mask': mask}
    if return_all:
        res['net'] = net
    return res
