"""CPU oracle of the UNet training step (TEST INFRASTRUCTURE ONLY -- never imported by the product).

What it restates: the graph of reference networks/unet.py `build` :224-262 in TRAIN mode (conv_block :265-277 with
`tf.layers.dropout(..., training=True)` :274-276, down_layer :282-296, up_layer :299-322 with the bridge :315-316),
the layer definitions DESIGN.md section 1 pins (the reference's own conv_layer :326-329 raises NotImplementedError),
the weighted softmax cross-entropy of oracle/loss_oracle.py, and TensorFlow's Adam update rule (the optimiser the
reference uses for its GAN, gan.py:740-751:  lr_t = lr * sqrt(1 - b2^t) / (1 - b1^t),  m += (g - m)(1 - b1),
v += (g^2 - v)(1 - b2),  p -= lr_t * m / (sqrt(v) + eps)).

PARITY UNPINNED against TensorFlow itself: the reference holds no loss, no optimiser and no test for a UNet training
step, and TensorFlow is not in this image.  What pins this oracle instead (tests/test_oracle_train.py): float64
autograd of torch on the CPU checked against central finite differences of its own loss, and known-answer values
of the dropout hash.  The dropout mask cannot match TensorFlow's Philox stream (the op seed is graph-dependent); it
is a counter-based hash both sides compute:
    u(i) = (mix64(mix64(seed_t ^ (block << 48)) + i) >> 40) / 2^24,   keep <=> u >= rate,   y = x / (1 - rate)
with seed_t = seed + 0x9E3779B97F4A7C15 * step (mod 2^64), block = level for the down blocks and nlev + level for the
up blocks, i the flat channels-last element index, mix64 the splitmix64 finaliser.
"""
import numpy as np

_M64 = (1 << 64) - 1


def _mix64_int(z):
    z &= _M64
    z ^= z >> 30
    z = (z * 0xBF58476D1CE4E5B9) & _M64
    z ^= z >> 27
    z = (z * 0x94D049BB133111EB) & _M64
    z ^= z >> 31
    return z


def _mix64_np(z):
    with np.errstate(over='ignore'):
        z = z ^ (z >> np.uint64(30))
        z = z * np.uint64(0xBF58476D1CE4E5B9)
        z = z ^ (z >> np.uint64(27))
        z = z * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return z


def step_seed(seed, step):
    return (int(seed) + 0x9E3779B97F4A7C15 * int(step)) & _M64


def dropout_keep(seed_t, block, count, rate):
    """Boolean keep mask of `count` elements (flat channels-last order) of dropout block `block`."""
    base = _mix64_int(seed_t ^ ((int(block) << 48) & _M64))
    with np.errstate(over='ignore'):
        z = np.uint64(base) + np.arange(count, dtype=np.uint64)
    u = (_mix64_np(z) >> np.uint64(40)).astype(np.float32) * np.float32(1.0 / 16777216.0)
    return u >= np.float32(rate)


def _torch():
    import torch
    return torch


def _to_torch_params(weights, ndim):
    """TF layouts -> torch layouts, float64 leaves that require grad."""
    torch = _torch()
    out = {}
    for name, arr in weights.items():
        a = torch.tensor(np.asarray(arr, dtype=np.float64))
        if name.endswith('/kernel'):
            if '/upscale/' in name:
                # TF (k..., out, in) -> conv_transpose weight (in, out, k...)
                perm = (ndim + 1, ndim) + tuple(range(ndim))
            else:
                # HWIO -> (out, in, k...)
                perm = (ndim + 1, ndim) + tuple(range(ndim))
            a = a.permute(*perm).contiguous()
        out[name] = a.requires_grad_(True)
    return out


def _from_torch_grad(name, g, ndim):
    """torch layout -> TF layout (numpy float64)."""
    if name.endswith('/kernel'):
        # inverse of (ndim+1, ndim, 0..ndim-1)
        perm = tuple(range(2, ndim + 2)) + (1, 0)
        g = g.permute(*perm)
    return g.detach().contiguous().numpy()


def forward_loss(params, image, labels, wmap, filters, bridge='concat', ndim=2, rate=0.0, seed_t=0):
    """Loss of one batch.  image (N,[D,]H,W,C); labels (N,[D,]H,W) class ids; wmap same shape; params: torch leaves in
    torch layouts (see _to_torch_params).  Returns (loss, logits channels-last)."""
    torch = _torch()
    F = torch.nn.functional
    conv = F.conv2d if ndim == 2 else F.conv3d
    convt = F.conv_transpose2d if ndim == 2 else F.conv_transpose3d
    pool = F.max_pool2d if ndim == 2 else F.max_pool3d
    nl = len(filters)
    to_cf = (0, ndim + 1) + tuple(range(1, ndim + 1))       # channels-last -> channels-first
    to_cl = (0,) + tuple(range(2, ndim + 2)) + (1,)
    x = torch.tensor(np.asarray(image, dtype=np.float64)).permute(*to_cf)
    keep_scale = float(np.float32(1.0) - np.float32(rate))

    def layer(x, scope):
        y = conv(x, params[scope + '/kernel'], params[scope + '/bias'], padding=1)
        if scope + '/scale' in params:
            # frozen per-channel affine (folded BN): (conv + bias) * scale + shift, constants of the step
            bshape = (1, -1) + (1,) * ndim
            y = y * params[scope + '/scale'].detach().reshape(bshape) + params[scope + '/shift'].detach().reshape(bshape)
        return torch.relu(y)

    def block(x, scope, block_id):
        y = layer(layer(x, scope + '/conv1'), scope + '/conv2')
        if rate > 0.0:
            cl_shape = tuple(y.permute(*to_cl).shape)
            keep = dropout_keep(seed_t, block_id, int(np.prod(cl_shape)), rate).reshape(cl_shape)
            keep = torch.tensor(keep).permute(*to_cf)
            y = torch.where(keep, y / keep_scale, torch.zeros_like(y))
        return y

    down = []
    for l in range(nl):
        if l > 0:
            x = pool(x, 2)
        x = block(x, 'UNet/down%d' % l, l)
        down.append(x)
    for l in reversed(range(nl - 1)):
        up = convt(x, params['UNet/up%d/upscale/kernel' % l], params['UNet/up%d/upscale/bias' % l], stride=2)
        if bridge == 'concat':
            m = torch.cat([up, down[l]], dim=1)
        elif bridge == 'eltwise_add':
            m = up + down[l]
        elif bridge == 'eltwise_mul':
            m = up * down[l]
        elif bridge == 'eltwise_sub':
            m = up - down[l]
        else:
            m = up
        x = block(m, 'UNet/up%d' % l, nl + l)
    logits = conv(x, params['UNet/to_image/kernel'], params['UNet/to_image/bias']).permute(*to_cl)
    k = logits.shape[-1]
    flat = logits.reshape(-1, k)
    lab = torch.tensor(np.asarray(labels).reshape(-1).astype(np.int64))
    w = torch.tensor(np.asarray(wmap, dtype=np.float64).reshape(-1))
    lse = torch.logsumexp(flat, dim=1)
    picked = flat.gather(1, lab[:, None])[:, 0]
    loss = (w * (lse - picked)).sum() / flat.shape[0]
    return loss, logits


def gradients(weights, image, labels, wmap, filters, bridge='concat', ndim=2, rate=0.0, seed=0, step=0):
    """(loss, {name: gradient in the TF layout}, logits) of one batch, float64."""
    params = _to_torch_params(weights, ndim)
    loss, logits = forward_loss(params, image, labels, wmap, filters, bridge, ndim, rate, step_seed(seed, step))
    loss.backward()
    grads = {name: _from_torch_grad(name, p.grad, ndim) for name, p in params.items()
             if name.endswith('/kernel') or name.endswith('/bias')}        # scale / shift are frozen
    return float(loss.item()), grads, logits.detach().numpy()


class Adam(object):
    """TensorFlow's Adam update rule in float64 (optimizer='sgd': p -= lr * g)."""

    def __init__(self, weights, learning_rate=1e-3, beta1=0.9, beta2=0.999, epsilon=1e-8, optimizer='adam'):
        self.w = {k: np.asarray(v, dtype=np.float64).copy() for k, v in weights.items()}      # incl. frozen scale / shift
        self.m = {k: np.zeros_like(v) for k, v in self.w.items()}
        self.v = {k: np.zeros_like(v) for k, v in self.w.items()}
        self.lr, self.b1, self.b2, self.eps, self.optimizer = learning_rate, beta1, beta2, epsilon, optimizer
        self.t = 0

    def apply(self, grads):
        self.t += 1
        if self.optimizer == 'sgd':
            for k, g in grads.items():
                self.w[k] -= np.float64(np.float32(self.lr)) * g
            return
        b1, b2 = np.float64(np.float32(self.b1)), np.float64(np.float32(self.b2))
        lr_t = np.float64(np.float32(self.lr)) * np.sqrt(1.0 - b2 ** self.t) / (1.0 - b1 ** self.t)
        for k, g in grads.items():
            self.m[k] += (g - self.m[k]) * (1.0 - b1)
            self.v[k] += (g * g - self.v[k]) * (1.0 - b2)
            self.w[k] -= lr_t * self.m[k] / (np.sqrt(self.v[k]) + np.float64(np.float32(self.eps)))
