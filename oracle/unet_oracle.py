"""torch-CPU fp32 restatement of the reference UNet graph (TEST INFRASTRUCTURE ONLY).

Follows ``/root/reference/sequitr/networks/unet.py:224-322`` with TensorFlow
semantics (NHWC/HWIO weights, SAME zero padding, ``tf.concat([upsampled, skip],
-1)`` unet.py:197, ReLU unet.py:142, dropout off at inference unet.py:274-276).
Used (a) to pin ``oracle/unet_ref.c`` to an independent convolution
implementation (oneDNN) and (b) as the multi-threaded CPU baseline that
``bench.py`` times ("reference CPU path").

"Parity unpinned": the reference has no concrete UNet2D/UNet3D, no golden
vectors, and its arithmetic lives in TensorFlow 1.x (absent here).
"""
import numpy as np
import torch
import torch.nn.functional as F


def _to_torch_conv_kernel(k):
    """HWIO / DHWIO -> OIHW / OIDHW."""
    nd = k.ndim - 2
    perm = (nd + 1, nd) + tuple(range(nd))
    return torch.from_numpy(np.ascontiguousarray(np.transpose(k, perm)))


def _to_torch_upconv_kernel(k):
    """TF conv_transpose (k,k[,k],Cout,Cin) -> torch ConvTranspose (Cin,Cout,k,k[,k])."""
    nd = k.ndim - 2
    perm = (nd + 1, nd) + tuple(range(nd))
    return torch.from_numpy(np.ascontiguousarray(np.transpose(k, perm)))


def unet_forward(x, weights, filters, bridge='concat', threads=None):
    """x: (N,H,W,C) / (N,D,H,W,C) float32 ndarray -> dict(logits, probs, mask), channels-last."""
    if threads:
        torch.set_num_threads(int(threads))
    nd = x.ndim - 2
    conv = F.conv2d if nd == 2 else F.conv3d
    convt = F.conv_transpose2d if nd == 2 else F.conv_transpose3d
    pool = F.max_pool2d if nd == 2 else F.max_pool3d
    to_cf = (0, nd + 1) + tuple(range(1, nd + 1))
    to_cl = (0,) + tuple(range(2, nd + 2)) + (1,)
    t = torch.from_numpy(np.ascontiguousarray(np.transpose(x, to_cf))).float()

    def layer(t, scope, relu=True):
        k = _to_torch_conv_kernel(weights[scope + '/kernel'])
        b = torch.from_numpy(weights[scope + '/bias'])
        y = conv(t, k, b, padding=k.shape[-1] // 2)
        if scope + '/scale' in weights:
            shp = (1, -1) + (1,) * nd
            y = y * torch.from_numpy(weights[scope + '/scale']).view(shp) + \
                torch.from_numpy(weights[scope + '/shift']).view(shp)
        return F.relu(y) if relu else y

    def block(t, scope):
        return layer(layer(t, scope + '/conv1'), scope + '/conv2')

    with torch.no_grad():
        net = [block(t, 'UNet/down0')]
        for i in range(1, len(filters)):
            net.append(block(pool(net[-1], 2), 'UNet/down%d' % i))
        for i in reversed(range(len(filters) - 1)):
            scope = 'UNet/up%d' % i
            up = convt(net[-1], _to_torch_upconv_kernel(weights[scope + '/upscale/kernel']),
                       torch.from_numpy(weights[scope + '/upscale/bias']), stride=2)
            skip = net[i]
            if bridge == 'concat':
                m = torch.cat([up, skip], 1)
            elif bridge == 'eltwise_add':
                m = up + skip
            elif bridge == 'eltwise_mul':
                m = up * skip
            elif bridge == 'eltwise_sub':
                m = up - skip
            elif bridge is None:
                m = up
            else:
                raise ValueError('Bridge type not recognized')
            net.append(block(m, scope))
        logits = layer(net[-1], 'UNet/to_image', relu=False)
        probs = torch.softmax(logits, 1)
        mask = torch.argmax(logits, 1).to(torch.uint8)
    return {'logits': logits.permute(to_cl).contiguous().numpy(),
            'probs': probs.permute(to_cl).contiguous().numpy(),
            'mask': mask.numpy()}
