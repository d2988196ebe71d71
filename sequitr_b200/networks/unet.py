"""Drop-in for the reference ``sequitr/networks/unet.py`` (``UNet`` base class,
:53-342) plus the concrete ``UNet2D`` / ``UNet3D`` the reference names (:56) but does
not ship.  TensorFlow is replaced by a CUDA plan behind the C ABI
(``sq_unet_create / load_weights / finalize / forward``).

Same constructor (``params`` dict keys ``name, filters, dropout, num_inputs,
num_outputs, shape, bridge, kernel`` with the reference defaults, :132-139), same
topology walk (``build`` :224-262, ``conv_block`` :265-277, ``down_layer`` :282-296,
``up_layer`` :299-322), same subclass hooks (``conv_layer``, ``conv_layer_1x1``,
``conv_transpose_layer``, ``max_pool_layer``, ``reshape_input``), same errors
(``ValueError('Bridge type not recognized')`` :186-187, ``NotImplementedError`` from
the abstract primitives :326-342).  Where TF builds a symbolic graph and runs it
in a session, ``build(features)`` here traces the same hooks symbolically (to check
the topology the plan implements) and then executes the plan, returning the logits.

Layer definitions (the reference leaves them abstract; see include/sequitr_b200.h):
3x3 SAME conv + bias (+ folded-BN affine) + ReLU; 2x2 max-pool; 2x2 stride-2
transposed conv + bias; 1x1 conv head.  Extra ``params`` key ``compute``:
``'bf16'`` (tcgen05 tensor cores, default) or ``'fp32'`` (bit-exact verification mode).
"""
import contextlib
import ctypes
import logging

import numpy as np

from .. import _lib, ops, synth

DEFAULT_FILTERS = (16, 32, 64, 128, 256)
DEFAULT_DROPOUT = 0.4
BRIDGE_TYPES = ('eltwise_add', 'eltwise_mul', 'eltwise_sub', 'concat', None)

logger = logging.getLogger('worker_process')


class ModeKeys(object):
    """tf.estimator.ModeKeys values (reference networks/unet.py:172)."""
    TRAIN = 'train'
    EVAL = 'eval'
    PREDICT = 'infer'


class _Sym(object):
    """Symbolic channels-last tensor used while tracing the topology."""

    def __init__(self, shape, op=None, scope=None, inputs=()):
        self.shape = tuple(shape)
        self.op, self.scope, self.inputs = op, scope, inputs

    @property
    def channels(self):
        return self.shape[-1]


class UNet(object):
    """ UNet

    ** This is the Base Class, use the sublasses UNet2D or UNet3D **
    (reference networks/unet.py:53-124 for the full description.)
    """

    def __init__(self, params, mode=ModeKeys.PREDICT):
        self._mode = mode
        self.params = dict(params)
        self._handle = None                     # a library handle of this network's own (params['private_handle'])
        self.name = params.get('name', 'UNet2d_test')
        self.filters = tuple(params.get('filters', DEFAULT_FILTERS))
        self.dropout = params.get('dropout', DEFAULT_DROPOUT)
        self.n_inputs = params.get('num_inputs', 1)
        self.n_outputs = params.get('num_outputs', 2)
        self.shape = tuple(params.get('shape', (1024, 1024)))
        self.bridge_type = params.get('bridge', 'eltwise_mul')
        self.kernel = tuple(params.get('kernel', (3, 3)))
        self.compute = params.get('compute', 'bf16')
        if self.compute not in ('bf16', 'fp32'):
            raise ValueError("compute must be 'bf16' or 'fp32'")
        self._activation = 'relu'
        self._initializer = 'variance_scaling'
        self._net = None
        self._scopes = []
        self._trace = []
        self._weights = None
        self._plan = None
        self._device = params.get('device', None)   # CUDA device index; None = the current device at first use
        self._ws = None

    # ------------------------------------------------------------ properties
    @property
    def width(self):
        """ width of the image volume """
        return self.shape[0]

    @property
    def height(self):
        """ height of the image volume """
        return self.shape[1]

    @property
    def slices(self):
        """ depth (number of slices) of the image volume """
        if self.ndim < 3:
            return 0
        return self.shape[2]

    @property
    def ndim(self):
        """ number of dimensions of image volume """
        return len(self.shape)

    @property
    def training(self):
        """ training mode flag """
        return self._mode == ModeKeys.TRAIN

    @property
    def btype(self):
        """ DEPRECATED: bridge type  """
        raise DeprecationWarning("Use @bridge_type")

    @property
    def bridge_type(self):
        return self._bridge_type

    @bridge_type.setter
    def bridge_type(self, bridge):
        """ Set the bridge type """
        if bridge not in BRIDGE_TYPES:
            raise ValueError('Bridge type not recognized')
        if bridge == 'concat':
            self.bridge = lambda x, y: _Sym(x.shape[:-1] + (x.channels + y.channels,), 'concat',
                                            self._scope(), (x, y))
        elif bridge is None:
            logger.warning('Bridge function in UNet not recognized')
            self.bridge = lambda x, y: x
        else:
            self.bridge = lambda x, y: _Sym(x.shape, bridge, self._scope(), (x, y))
        self._bridge_type = bridge
        self._drop_plan()

    # ------------------------------------------------------------ graph walk
    @contextlib.contextmanager
    def variable_scope(self, name):
        self._scopes.append(name)
        try:
            yield
        finally:
            self._scopes.pop()

    def _scope(self):
        return '/'.join(self._scopes)

    def reshape_input(self, features):
        """ Reshape the input layer from the dataset features:
        (batch, depth (aka slices), height, width, channels) """
        full_shape = [-1, self.slices, self.width, self.height, self.n_inputs]
        input_shape = [d for d in full_shape if d != 0]
        return features.reshape(input_shape)

    def logits(self):
        """ return the un-normalized logits (i.e. last) layer of the network """
        return self._net[-1]

    def build(self, features):
        """ build

        Build the network using the given parameters and the features and run it.
        Returns the final output layer (logits, channels-last, float32); a numpy
        array for numpy features, a cuda tensor for cuda-tensor features.
        """
        logger.info('Building UNet ({0:s})...'.format(self.__class__.__name__))
        self._trace = []
        with self.variable_scope('UNet'):
            input_layer = self.reshape_input(features)
            x = _Sym(tuple(input_layer.shape), 'input', self._scope())

            # BUILD THE NET!
            self._net = [self.down_layer(x, self.filters[0], name=0)]

            # do the down layers
            for i, f in enumerate(self.filters[1:]):
                prev_layer = self.max_pool_layer(self._net[-1])
                self._net.append(self.down_layer(prev_layer, f, name=i + 1))

            # now add the up layers
            for i, f in reversed(list(enumerate(self.filters[:-1]))):
                prev_layer = self._net[-1]  # layer below
                bridge = self._net[i]       # bridge information
                self._net.append(self.up_layer(prev_layer, f, bridge, name=i))

            # make an output layer with a 1x1 convolution
            with self.variable_scope('to_image'):
                logits_sym = self.conv_layer_1x1(self._net[-1], self.n_outputs)

        logger.info('Output layer -> shape {0:s}'.format(str(logits_sym.shape)))
        self._check_trace()
        logits = self._execute(input_layer, want=('logits',))['logits']
        # append this layer for completeness
        self._net.append(logits)
        logger.info('...Done')
        return logits

    def conv_block(self, x, filters):
        """ convolutional block """
        with self.variable_scope('conv1'):
            conv1 = self.conv_layer(x, filters)
        with self.variable_scope('conv2'):
            conv2 = self.conv_layer(conv1, filters)
        # Dropout (tf.layers.dropout, networks/unet.py:274-276) is the identity unless training; in TRAIN mode the
        # device applies it after conv2 of every block inside the training step (see ``Trainer``)
        return conv2

    def down_layer(self, x, filters, name=None):
        """ down_layer: 2x [3x3 convolution, ReLu]; tensor shape NHWC """
        logger.info('Down layer -> shape {0:s}'.format(str(x.shape)))
        with self.variable_scope('down{0:d}'.format(name)):
            out = self.conv_block(x, filters)
        return out

    def up_layer(self, x, filters, bridge, name=None):
        """ up_layer: transpose convolution, bridge, conv block """
        logger.info('Up layer -> shape {0:s} (bridge: {1:s})'.format(str(x.shape),
                                                                      str(self.bridge_type)))
        with self.variable_scope('up{0:d}'.format(name)):
            # scale up the image
            with self.variable_scope('upscale'):
                upscale = self.conv_transpose_layer(x, filters)
            # now we need to incorporate the filters using the bridge
            with self.variable_scope('bridge'):
                bridge = self.bridge(upscale, bridge)
            out = self.conv_block(bridge, filters)
        return out

    def conv_layer(self, x, filters):
        """ Convolution layer, conv-relu with padding """
        raise NotImplementedError

    def conv_layer_1x1(self, x, filters):
        """ Return a 1x1 convolution layer """
        raise NotImplementedError

    def conv_transpose_layer(self, x, filters):
        """ Transpose convolution (aka deconvolution) layer """
        raise NotImplementedError

    def max_pool_layer(self, x):
        """ Max pool operation """
        raise NotImplementedError

    # the reference declares it under this name (:340) but calls max_pool_layer (:242)
    pool_layer = max_pool_layer

    # ------------------------------------------------------------ weights
    def load_weights(self, weights):
        """weights: dict TF-scope name -> float32 ndarray (see synth.unet_weights)."""
        self._weights = {k: np.ascontiguousarray(v, dtype=np.float32) for k, v in weights.items()}
        self._drop_plan()

    def _drop_plan(self):
        plan = getattr(self, '_plan', None)
        if plan is not None:
            _lib.load().sq_unet_destroy(plan)
        self._plan = None

    def __del__(self):
        try:
            self._drop_plan()
            h = getattr(self, '_handle', None)
            if h is not None:
                _lib.load().sq_destroy(h)
                self._handle = None
        except Exception:
            pass

    def trainer(self, **kwargs):
        """A ``Trainer`` bound to this network (``compute='fp32'``): ``loss = net.trainer().step(img, target)``."""
        return Trainer(self, **kwargs)

    def twin(self):
        """A second network with the same parameters and weights on the same GPU but with a library handle of its
        own: its host calls (``segment_and_localise``) run concurrently with this one's, so the copy ramp and the
        label / read-back tail of one call hide under the other's convolutions (``shard.segment_stack(overlap=True)``)."""
        self._ensure_plan()
        params = dict(self.params)
        params['device'] = self._device
        params['private_handle'] = True
        other = type(self)(params)
        other.load_weights(self._weights)
        return other

    def _ensure_plan(self, device=None):
        """The CUDA plan lives on ONE device: the first one it is used on (default: the current
        device).  Buffers of another GPU are refused instead of being handed to the wrong stream."""
        import torch
        if device is not None:
            device = torch.device(device).index
            if device is None:
                device = torch.cuda.current_device()
        if self._plan is not None:
            if device is not None and device != self._device:
                raise ValueError('this network is bound to cuda:%d, the features live on cuda:%d '
                                 '(build one network per GPU)' % (self._device, device))
            return self._plan
        if any(k != 3 for k in self.kernel):
            raise NotImplementedError('only 3x3(x3) kernels are implemented')
        if self._weights is None:
            # the reference initialises variables with variance_scaling (networks/unet.py:143)
            self._weights = synth.unet_weights(self.filters, self.n_inputs, self.n_outputs,
                                               ndim=self.ndim, bridge=self.bridge_type)
        lib = _lib.load()
        filt = (ctypes.c_int * len(self.filters))(*self.filters)
        plan = ctypes.c_void_p()
        mode = _lib.MODE_BF16_TC if self.compute == 'bf16' else _lib.MODE_FP32_EXACT
        if device is None:
            device = self._device
        if device is None:
            _lib.handle()                       # raises without a CUDA device
            device = torch.cuda.current_device()
        self._device = int(device)
        device = self._device
        self._ws = ops.Workspace(device='cuda:%d' % device)
        if self.params.get('private_handle') and getattr(self, '_handle', None) is None:
            self._handle = _lib.new_handle(device)      # this network's own streams / arena (see twin())
        hnd = getattr(self, '_handle', None) or _lib.handle(device)
        _lib.check(lib.sq_unet_create(hnd, self.ndim, self.n_inputs, self.n_outputs, filt,
                                      len(self.filters), _lib.BRIDGE_CODES[self.bridge_type], mode,
                                      ctypes.byref(plan)))
        try:
            for name, arr in self._weights.items():
                shape = (ctypes.c_int64 * arr.ndim)(*arr.shape)
                _lib.check(lib.sq_unet_load_weights(plan, name.encode(), arr.ctypes.data, shape,
                                                    arr.ndim))
            _lib.check(lib.sq_unet_finalize(plan))
        except Exception:
            lib.sq_unet_destroy(plan)
            raise
        self._plan = plan
        return plan

    def _expected_trace(self):
        ops_ = []
        cin = self.n_inputs
        for i, f in enumerate(self.filters):
            if i > 0:
                ops_.append(('pool', 'UNet', cin, cin))
            ops_.append(('conv', 'UNet/down%d/conv1' % i, cin, f))
            ops_.append(('conv', 'UNet/down%d/conv2' % i, f, f))
            cin = f
        for i in reversed(range(len(self.filters) - 1)):
            f = self.filters[i]
            ops_.append(('upconv', 'UNet/up%d/upscale' % i, cin, f))
            merged = 2 * f if self.bridge_type == 'concat' else f
            ops_.append(('conv', 'UNet/up%d/conv1' % i, merged, f))
            ops_.append(('conv', 'UNet/up%d/conv2' % i, f, f))
            cin = f
        ops_.append(('conv1x1', 'UNet/to_image', cin, self.n_outputs))
        return ops_

    def _check_trace(self):
        if self._trace != self._expected_trace():
            raise NotImplementedError('the traced topology differs from the one the CUDA plan '
                                      'implements (reference networks/unet.py:224-262)')

    # ------------------------------------------------------------ execution
    def _execute(self, input_layer, want=('logits', 'probs', 'mask')):
        import torch
        lib = _lib.load()
        is_numpy = isinstance(input_layer, np.ndarray)
        if is_numpy:
            plan = self._ensure_plan()
            x = torch.from_numpy(np.ascontiguousarray(input_layer, dtype=np.float32)).to('cuda:%d' % self._device)
        else:
            x = input_layer.contiguous().float()
            if not x.is_cuda:
                raise ValueError('features must be a numpy array or a cuda tensor')
            plan = self._ensure_plan(x.device)
        if x.dim() != self.ndim + 2 or x.shape[-1] != self.n_inputs:
            raise ValueError('features have shape %s, expected (N,%s%d)' %
                             (tuple(x.shape), 'D,H,W,' if self.ndim == 3 else 'H,W,', self.n_inputs))
        n = x.shape[0]
        d, h, w = (1,) + tuple(x.shape[1:3]) if self.ndim == 2 else tuple(x.shape[1:4])
        need = ctypes.c_size_t()
        _lib.check(lib.sq_unet_workspace_bytes(plan, n, d, h, w, ctypes.byref(need)))
        ws = self._ws.get(need.value)
        sp = tuple(x.shape[:-1])
        out = {}
        if 'logits' in want:
            out['logits'] = torch.empty(sp + (self.n_outputs,), dtype=torch.float32, device=x.device)
        if 'probs' in want:
            out['probs'] = torch.empty(sp + (self.n_outputs,), dtype=torch.float32, device=x.device)
        if 'mask' in want:
            out['mask'] = torch.empty(sp, dtype=torch.uint8, device=x.device)
        _lib.check(lib.sq_unet_forward(plan, x.data_ptr(), n, d, h, w, _lib.ptr(out.get('probs')),
                                       _lib.ptr(out.get('mask')), _lib.ptr(out.get('logits')),
                                       ws.data_ptr(), ws.numel(), _lib.stream_ptr(x.device)))
        if is_numpy:
            out = {k: v.cpu().numpy() for k, v in out.items()}
        return out

    def _as_input(self, features):
        """Channels-last batches pass through; anything else goes through reshape_input."""
        if features.ndim == self.ndim + 2 and features.shape[-1] == self.n_inputs:
            return features
        return self.reshape_input(features)

    def predict(self, features, want=('logits', 'probs', 'mask')):
        """Logits + softmax probabilities + argmax mask (uint8, first max wins) -- the
        head the consumer ``utils.CentroidWriter`` implies (reference utils.py:492)."""
        return self._execute(self._as_input(features), want)

    def profile(self, features):
        """Per-layer device times of one instrumented forward pass:
        list of (scope, milliseconds, algorithmic FLOPs).  ``features`` must be a cuda tensor."""
        lib = _lib.load()
        x = self._as_input(features).contiguous().float()
        plan = self._ensure_plan(x.device)
        n = x.shape[0]
        d, h, w = (1,) + tuple(x.shape[1:3]) if self.ndim == 2 else tuple(x.shape[1:4])
        need = ctypes.c_size_t()
        _lib.check(lib.sq_unet_workspace_bytes(plan, n, d, h, w, ctypes.byref(need)))
        ws = self._ws.get(need.value)
        cap = 64
        names = (ctypes.c_char_p * cap)()
        ms = (ctypes.c_float * cap)()
        flops = (ctypes.c_double * cap)()
        nl = ctypes.c_int()
        _lib.check(lib.sq_unet_profile(plan, x.data_ptr(), n, d, h, w, ws.data_ptr(), ws.numel(),
                                       _lib.stream_ptr(x.device), names, ms, flops, cap, ctypes.byref(nl)))
        return [(names[i].decode(), float(ms[i]), float(flops[i])) for i in range(min(nl.value, cap))]

    def launches(self):
        """Kernel launches issued by the last forward pass."""
        n = ctypes.c_int()
        _lib.check(_lib.load().sq_unet_last_launches(self._ensure_plan(), ctypes.byref(n)))
        return n.value

    def segment(self, features):
        """uint8 class mask (N,[D,]H,W)."""
        return self.predict(features, want=('mask',))['mask']

    def _result_buffers(self, n, max_rows):
        key = (int(n), int(max_rows))
        cache = self.__dict__.setdefault('_result_cache', {})
        if key not in cache:
            from ..utils import pinned_array
            cache.clear()                                   # one size at a time: a stack is processed in equal calls
            cache[key] = (pinned_array((n, max_rows, 5), np.float32), pinned_array((n,), np.int32))
        return cache[key]

    def segment_and_localise(self, frames, frame0=0, max_rows=4096, return_mask=False, normalise=False):
        """The whole hot path on HOST frames (2-D): H2D -> UNet -> argmax -> label-and-
        localise -> D2H.  frames float32 (N,H,W,Cin), or RAW uint8 / uint16 camera frames as the
        readers in ``dataio`` deliver them (they then cross PCIe in 1-2 bytes per pixel and are
        widened on the device).  ``normalise`` applies ImageNorm per frame on the device first.
        Returns the list of per-frame (n_i,5) float32 centroid tables (rows as
        utils.CentroidWriter writes them)."""
        plan = self._ensure_plan()
        lib = _lib.load()
        frames = np.asarray(frames)
        if frames.dtype == np.uint8:
            code = _lib.U8
        elif frames.dtype == np.uint16:
            code = _lib.U16
        else:
            code = _lib.F32
            frames = frames.astype(np.float32, copy=False)
        frames = np.ascontiguousarray(self._as_input(frames))
        n, h, w = frames.shape[:3]
        mask = np.empty((n, h, w), dtype=np.uint8) if return_mask else None
        while True:
            # the result buffers are page-locked and kept with the net: the tables then come back by DMA instead of a
            # staged copy into pageable memory, and a time-lapse job does not allocate per call
            table, counts = self._result_buffers(n, max_rows)
            st = lib.sq_segment_localise_raw_host(plan, frames.ctypes.data, code, int(bool(normalise)), n, h, w,
                                                  frame0, table.ctypes.data, counts.ctypes.data, max_rows,
                                                  _lib.ptr(mask))
            if st == _lib.SQ_EOVERFLOW:
                max_rows = int(counts.max())
                continue
            _lib.check(st)
            break
        tables = [table[i, :counts[i]].copy() for i in range(n)]
        return (tables, mask) if return_mask else tables


class _ConcreteUNet(UNet):
    """Shared primitive definitions of UNet2D / UNet3D (symbolic tracing side)."""

    def _record(self, kind, cin, cout):
        self._trace.append((kind, self._scope() if kind != 'pool' else 'UNet', cin, cout))

    def conv_layer(self, x, filters):
        """ Convolution layer, conv-relu with SAME padding """
        self._record('conv', x.channels, filters)
        return _Sym(x.shape[:-1] + (filters,), 'conv', self._scope(), (x,))

    def conv_layer_1x1(self, x, filters):
        """ 1x1 convolution layer (logits, no activation) """
        self._record('conv1x1', x.channels, filters)
        return _Sym(x.shape[:-1] + (filters,), 'conv1x1', self._scope(), (x,))

    def conv_transpose_layer(self, x, filters):
        """ 2x2 stride-2 transpose convolution: doubles the spatial size """
        self._record('upconv', x.channels, filters)
        sp = tuple(2 * s if s and s > 0 else s for s in x.shape[1:-1])
        return _Sym((x.shape[0],) + sp + (filters,), 'upconv', self._scope(), (x,))

    def max_pool_layer(self, x):
        """ 2x2 max pool, stride 2 """
        self._record('pool', x.channels, x.channels)
        sp = tuple(s // 2 if s and s > 0 else s for s in x.shape[1:-1])
        return _Sym((x.shape[0],) + sp + (x.channels,), 'pool', self._scope(), (x,))

    pool_layer = max_pool_layer


class UNet2D(_ConcreteUNet):
    """ 2-D UNet: features (N,H,W,Cin) NHWC (or anything reshape_input accepts). """

    def __init__(self, params, mode=ModeKeys.PREDICT):
        _ConcreteUNet.__init__(self, params, mode)
        if self.ndim != 2:
            raise ValueError('UNet2D needs a 2-D shape')


class UNet3D(_ConcreteUNet):
    """ 3-D UNet: params['shape'] = (width, height, slices); features (N,D,H,W,Cin). """

    def __init__(self, params, mode=ModeKeys.PREDICT):
        p = dict(params)
        p.setdefault('kernel', (3, 3, 3))
        _ConcreteUNet.__init__(self, p, mode)
        if self.ndim != 3:
            raise ValueError('UNet3D needs a 3-D shape (width, height, slices)')



class Trainer(object):
    """The training step of a UNet on the device (BASELINE config 5): forward with dropout (rate
    ``net.dropout``, reference networks/unet.py:274-276), weighted softmax cross-entropy over the per-pixel
    ``weights`` map, backward through every layer, Adam (TensorFlow's update rule) or SGD update of the plan's kernels
    and biases in place -- ``net.predict`` right after ``step`` uses the updated weights.  The reference ships neither
    loss nor optimiser for the UNet (see ``sq_trainer_step`` in include/sequitr_b200.h); the inputs are what its
    ``tr_augment`` yields: ``image`` (N,[D,]H,W,C) float32 and ``{'label': one-hot uint8 (N,[D,]H,W,K) or class ids
    (N,[D,]H,W), 'weights': float32 (N,[D,]H,W[,1])}``.

    Kernels and biases are trained; a layer's optional per-channel affine (``scale`` / ``shift``, folded BN
    statistics) stays frozen.  The network must be built with ``compute='fp32'``; ``weights()`` returns the variables
    in the ``load_weights`` layout (e.g. to load a ``compute='bf16'`` network for tensor-core inference)."""

    def __init__(self, net, learning_rate=1e-3, optimizer='adam', beta1=0.9, beta2=0.999, epsilon=1e-8,
                 dropout=None, seed=0, data_parallel=None):
        if net.compute != 'fp32':
            raise ValueError("the training step runs on a compute='fp32' network")
        if optimizer not in ('sgd', 'adam'):
            raise ValueError("optimizer must be 'sgd' or 'adam'")
        self.net = net
        self._plan = net._ensure_plan()
        self._trainer = ctypes.c_void_p()
        rate = net.dropout if dropout is None else dropout
        _lib.check(_lib.load().sq_trainer_create(self._plan, 1 if optimizer == 'adam' else 0, learning_rate, beta1,
                                                 beta2, epsilon, float(rate), int(seed),
                                                 ctypes.byref(self._trainer)))
        self._ws = ops.Workspace(device='cuda:%d' % net._device)
        # data parallel (one process per GPU, torch.distributed initialised by the caller): every rank steps on its
        # share of the batch, the gradient arena is averaged with ONE all-reduce, every rank applies the same update.
        # None = on whenever a process group with more than one rank exists.
        self.data_parallel = data_parallel
        self._grad = None

    def _grad_tensor(self):
        """The trainer's gradient arena as a zero-copy float32 cuda tensor."""
        import torch
        if self._grad is None:
            ptr, count = ctypes.c_void_p(), ctypes.c_size_t()
            _lib.check(_lib.load().sq_trainer_grad_arena(self._trainer, ctypes.byref(ptr), ctypes.byref(count)))

            class _Arena(object):
                __cuda_array_interface__ = {'shape': (count.value,), 'typestr': '<f4', 'data': (ptr.value, False),
                                            'version': 2}
            self._grad = torch.as_tensor(_Arena(), device='cuda:%d' % self.net._device)
        return self._grad

    def _ranks(self):
        import torch.distributed as dist
        if self.data_parallel is False or not (dist.is_available() and dist.is_initialized()):
            return 1
        return dist.get_world_size()

    def close(self):
        t = getattr(self, '_trainer', None)
        self._grad = None                       # the view dies with the arena
        if t is not None and t.value:
            _lib.load().sq_trainer_destroy(t)
        self._trainer = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check_alive(self):
        if self._trainer is None or self.net._plan is None or self.net._plan.value != self._plan.value:
            raise RuntimeError('the network was rebuilt (load_weights) after this trainer was made')

    def step(self, image, target, weights=None, apply_update=True):
        """One training step on a batch; returns the loss (float).  ``target``: the dict ``tr_augment`` returns,
        or the labels with ``weights`` given separately."""
        import torch
        self._check_alive()
        net = self.net
        dev = 'cuda:%d' % net._device
        if isinstance(target, dict):
            target, weights = target['label'], target['weights']

        def to_dev(v, dtype):
            t = v if torch.is_tensor(v) else torch.from_numpy(np.ascontiguousarray(v))
            return t.to(device=dev, dtype=dtype).contiguous()

        x = to_dev(image, torch.float32)
        if x.dim() != net.ndim + 2 or x.shape[-1] != net.n_inputs:
            raise ValueError('image has shape %s, expected (N,%s%d)' %
                             (tuple(x.shape), 'D,H,W,' if net.ndim == 3 else 'H,W,', net.n_inputs))
        sp = tuple(x.shape[:-1])
        lab, wgt = self._class_ids(to_dev(target, torch.uint8), to_dev(weights, torch.float32), sp, net.n_outputs)
        n = sp[0]
        d, h, w = ((1,) + sp[1:3]) if net.ndim == 2 else sp[1:4]
        lib = _lib.load()
        need = ctypes.c_size_t()
        _lib.check(lib.sq_trainer_workspace_bytes(self._trainer, n, d, h, w, ctypes.byref(need)))
        ws = self._ws.get(need.value)
        loss = torch.zeros(1, dtype=torch.float64, device=dev)
        ranks = self._ranks()
        _lib.check(lib.sq_trainer_step(self._trainer, x.data_ptr(), lab.data_ptr(), wgt.data_ptr(), n, d, h, w,
                                       1 if (apply_update and ranks == 1) else 0, loss.data_ptr(), ws.data_ptr(),
                                       ws.numel(), _lib.stream_ptr(x.device)))
        if ranks > 1:
            from .. import shard
            with torch.cuda.device(x.device):
                shard.average_gradients_(self._grad_tensor(), loss)
            if apply_update:
                _lib.check(lib.sq_trainer_apply(self._trainer, _lib.stream_ptr(x.device)))
        return float(loss.item())

    @staticmethod
    def _class_ids(lab, wgt, sp, num_outputs):
        """Labels and weights in the layout ``sq_trainer_step`` takes (any device): class ids uint8 ``sp`` and weights
        float32 ``sp`` from class ids ``sp`` / ``sp + (1,)`` or the one-hot ``sp + (K,)`` that ``tr_augment`` yields
        (reference networks/unet.py:396-398).  A one-hot row of zeros (a class beyond the K outputs) contributes
        nothing to a softmax cross-entropy: its weight is cleared."""
        import torch
        sp = tuple(sp)
        if tuple(wgt.shape) not in (sp, sp + (1,)):
            raise ValueError('weights have shape %s, expected %s' % (tuple(wgt.shape), sp))
        wgt = wgt.reshape(sp)
        if tuple(lab.shape) == sp + (num_outputs,) and not (num_outputs == 1 and tuple(lab.shape) == sp + (1,)):
            none = lab.sum(-1) == 0
            lab = lab.argmax(-1).to(torch.uint8).contiguous()
            if bool(none.any()):
                wgt = torch.where(none, torch.zeros_like(wgt), wgt)
        elif tuple(lab.shape) == sp + (1,):
            lab = lab.reshape(sp)
        elif tuple(lab.shape) != sp:
            raise ValueError('labels have shape %s, expected %s (class ids) or %s (one-hot)' %
                             (tuple(lab.shape), sp, sp + (num_outputs,)))
        if lab.numel() and int(lab.max()) >= num_outputs:
            raise ValueError('class id %d >= num_outputs %d' % (int(lab.max()), num_outputs))
        return lab.contiguous(), wgt.contiguous()

    def _read(self, what):
        self._check_alive()
        lib = _lib.load()
        out = {}
        for name, arr in self.net._weights.items():
            if not (name.endswith('/kernel') or name.endswith('/bias')):
                if what == 0:
                    out[name] = arr.copy()          # frozen per-channel affine (scale / shift): not trained
                continue
            buf = np.empty(arr.shape, dtype=np.float32)
            _lib.check(lib.sq_trainer_read(self._trainer, name.encode(), what, buf.ctypes.data, buf.size))
            out[name] = buf
        return out

    def weights(self):
        """The current kernels and biases, in the layout ``UNet.load_weights`` takes."""
        return self._read(0)

    def gradients(self):
        """The gradients of the last step, same names and layouts."""
        return self._read(1)


def tr_augment(features, params, rng=None):
    """ Augment the dataset by random cropping, flipping and rotations

    Drop-in for the reference ``tr_augment`` (networks/unet.py:348-401): identical augmentations
    are applied to the image (bilinear), the labels (nearest) and the weight map (bilinear, plus one
    outside the rotated frame :380-383), the result is cropped to ``params['shape']`` and the label
    is expanded to ``params['num_outputs']`` one-hot channels (:396-398).  The four full-size
    ``tf.contrib.image.rotate`` calls and the crop run as ONE gather kernel (``sq_tr_augment``).

    features: dict with 'image' (N,H,W,C) float32, 'label' (N,H,W[,1]) uint8, 'weights'
    (N,H,W[,1]) float32, 'shape' = (N,H,W,C) -- NumPy arrays or cuda tensors.  The random draws
    (theta = 2*pi*U[0,1) :370, crop origin :387-388) come from ``rng`` (``numpy.random.RandomState``,
    one draw per example); pass ``features['theta']`` / ``features['crop']`` to fix them.
    Returns ``img, {'label': label, 'weights': weights}`` like the reference (cuda tensors).
    """
    import torch
    rng = rng or np.random
    dev = 'cuda'

    def to_dev(v, dtype):
        t = v if torch.is_tensor(v) else torch.from_numpy(np.ascontiguousarray(v))
        return t.to(device=dev, dtype=dtype).contiguous()

    img = to_dev(features['image'], torch.float32)
    if img.dim() == 3:
        img = img[None]
    n, height, width, _ = (int(v) for v in img.shape)
    label = to_dev(features['label'], torch.uint8).reshape(n, height, width)
    weights = to_dev(features['weights'], torch.float32).reshape(n, height, width)
    if 'shape' in features and tuple(int(v) for v in features['shape'])[1:3] != (height, width):
        raise ValueError("features['shape'] does not match the image")
    outputs = params.get('num_outputs', 2)
    ch, cw = params.get('shape', (512, 512))[0:2]
    if ch > height or cw > width:
        raise ValueError('crop (%d,%d) larger than the image (%d,%d)' % (ch, cw, height, width))
    thetas = features.get('theta')
    if thetas is None:
        thetas = [np.float32(2.) * np.float32(rng.uniform()) * np.float32(np.pi) for _ in range(n)]
    thetas = np.broadcast_to(np.asarray(thetas, dtype=np.float32), (n,))
    crops = features.get('crop')
    if crops is None:
        # the reference's `if (ch,cw != height,width)` is always true (:386); tf.random_uniform needs
        # maxval > 0, so a crop as large as the image keeps origin 0 here instead of raising
        crops = [(rng.randint(0, height - ch) if height > ch else 0,
                  rng.randint(0, width - cw) if width > cw else 0) for _ in range(n)]
    crops = np.broadcast_to(np.asarray(crops, dtype=np.int32), (n, 2))
    transforms = np.stack([ops.rotation_transform(t, height, width) for t in thetas])
    img, label, weights = ops.tr_augment(img, label, weights, transforms, crops, ch, cw, outputs)
    return img, {'label': label, 'weights': weights[..., None]}


def preprocess_norm(features):
    """ normalise images or volumes to mean 0. and std 1.0

    The reference computes the normalised image and then returns the UN-normalised ``features``
    (networks/unet.py:405-432: the assignment at :430 is commented out), so this is the identity;
    use ``pipeline.ImageNorm`` / ``segment_and_localise(normalise=True)`` for the active transform.
    """
    return features
