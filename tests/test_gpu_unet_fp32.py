"""fp32 'exact' UNet mode: logits and masks bit-identical to the plain-C oracle."""
import os

import numpy as np
import pytest

from oracle import unet_c
from sequitr_b200 import synth

pytestmark = pytest.mark.gpu


def _net(cls, filters, shape, bridge, cin=1, k=2, weights=None, **kw):
    net = cls({'filters': filters, 'shape': shape, 'bridge': bridge, 'num_inputs': cin,
               'num_outputs': k, 'compute': 'fp32'})
    if weights is not None:
        net.load_weights(weights)
    return net


@pytest.mark.parametrize('bridge', ['concat', 'eltwise_mul', 'eltwise_add', 'eltwise_sub', None])
def test_unet2d_bit_exact(sq, bridge):
    from sequitr_b200.networks import UNet2D
    filters = (16, 32, 64)
    w = synth.unet_weights(filters, 3, 3, bridge=bridge, affine=(bridge == 'concat'), seed=7)
    x = synth.frames(2, 48, 80, 3, seed=2, n_objects=4)
    out = _net(UNet2D, filters, (48, 80), bridge, 3, 3, w).predict(x)
    ref = unet_c.unet_forward(x, w, filters, bridge)
    np.testing.assert_array_equal(out['logits'], ref['logits'])
    np.testing.assert_array_equal(out['mask'], ref['mask'])
    assert out['mask'].dtype == np.uint8
    np.testing.assert_allclose(out['probs'], ref['probs'], atol=3e-7)     # expf: CUDA vs glibc
    np.testing.assert_allclose(out['probs'].sum(-1), 1.0, atol=1e-6)


def test_unet2d_default_filters_and_build_entry_point(sq):
    from sequitr_b200.networks import UNet2D
    filters = (16, 32, 64, 128, 256)
    w = synth.unet_weights(filters, 1, 2, bridge='concat', seed=42)
    x = synth.frames(1, 64, 96, 1, seed=5, n_objects=3)
    net = _net(UNet2D, filters, (64, 96), 'concat', weights=w)
    logits = net.build(x.reshape(1, -1))                 # reference entry point: flat features
    ref = unet_c.unet_forward(x, w, filters, 'concat')
    np.testing.assert_array_equal(logits, ref['logits'])
    assert net.logits() is logits
    np.testing.assert_array_equal(net.segment(x), ref['mask'])


def test_known_answer(sq, golden_dir):
    from sequitr_b200.networks import UNet2D
    g = np.load(os.path.join(golden_dir, 'unet_kat.npz'))
    filters = tuple(int(f) for f in g['filters'])
    w = synth.unet_weights(filters, 1, 2, ndim=2, bridge='concat', seed=5)
    out = _net(UNet2D, filters, g['x'].shape[1:3], 'concat', weights=w).predict(g['x'])
    np.testing.assert_array_equal(out['logits'], g['logits'])
    np.testing.assert_array_equal(out['mask'], g['mask'])


def test_unet3d_bit_exact(sq):
    from sequitr_b200.networks import UNet3D
    filters = (8, 16, 32)
    w = synth.unet_weights(filters, 1, 3, ndim=3, bridge='concat', seed=3)
    x = synth.volumes(1, 8, 16, 24, 1)
    net = UNet3D({'filters': filters, 'shape': (16, 24, 8), 'bridge': 'concat', 'num_outputs': 3,
                  'compute': 'fp32'})
    net.load_weights(w)
    out = net.predict(x)
    ref = unet_c.unet_forward(x, w, filters, 'concat')
    np.testing.assert_array_equal(out['logits'], ref['logits'])
    np.testing.assert_array_equal(out['mask'], ref['mask'])


def test_errors(sq):
    from sequitr_b200.networks import UNet2D
    filters = (8, 16, 32)
    net = _net(UNet2D, filters, (18, 24), 'concat')
    with pytest.raises(ValueError):                      # 18 is not divisible by 4
        net.predict(np.zeros((1, 18, 24, 1), np.float32))
    bad = synth.unet_weights(filters, 1, 2)
    bad['UNet/down0/conv1/kernel'] = np.zeros((3, 3, 2, 8), np.float32)
    net2 = _net(UNet2D, filters, (16, 16), 'concat', weights=bad)
    with pytest.raises(ValueError):
        net2.predict(np.zeros((1, 16, 16, 1), np.float32))
    missing = synth.unet_weights(filters, 1, 2)
    del missing['UNet/up0/upscale/bias']
    net3 = _net(UNet2D, filters, (16, 16), 'concat', weights=missing)
    with pytest.raises(RuntimeError):
        net3.predict(np.zeros((1, 16, 16, 1), np.float32))


def test_segment_and_localise_end_to_end(sq):
    from oracle import centroid_oracle
    from sequitr_b200.networks import UNet2D
    filters = (16, 32, 64)
    w = synth.blob_detector_weights(filters, 1, 2, seed=1)
    x = synth.frames(3, 128, 160, 1, seed=11, n_objects=8)
    net = _net(UNet2D, filters, (128, 160), 'concat', weights=w)
    tables, mask = net.segment_and_localise(x, frame0=40, return_mask=True)
    ref = unet_c.unet_forward(x, w, filters, 'concat')
    np.testing.assert_array_equal(mask, ref['mask'])
    assert 3 <= len(tables[0]) <= 12                     # the blob detector finds the discs
    want = centroid_oracle.centroid_tables(ref['mask'])
    for i in range(3):
        want[i][:, 0] += 40
        np.testing.assert_array_equal(tables[i], want[i])
