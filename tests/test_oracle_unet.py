"""Pin the plain-C UNet oracle to an independent (torch/oneDNN) convolution."""
import os

import numpy as np
import pytest

from oracle import unet_c, unet_oracle
from sequitr_b200 import synth


@pytest.mark.parametrize("nd,bridge", [(2, 'concat'), (2, 'eltwise_mul'), (2, 'eltwise_add'),
                                       (2, 'eltwise_sub'), (2, None), (3, 'concat')])
def test_c_oracle_matches_torch(nd, bridge):
    filters = (8, 16, 32)
    shape = (2, 32, 48, 3) if nd == 2 else (1, 8, 16, 16, 1)
    w = synth.unet_weights(filters, shape[-1], 3, ndim=nd, bridge=bridge, affine=(bridge == 'concat'))
    x = np.random.default_rng(3).standard_normal(shape).astype(np.float32)
    a = unet_c.unet_forward(x, w, filters, bridge)
    b = unet_oracle.unet_forward(x, w, filters, bridge)
    assert a['logits'].shape == b['logits'].shape == shape[:-1] + (3,)
    np.testing.assert_allclose(a['logits'], b['logits'], atol=2e-4 * max(1.0, np.abs(b['logits']).max()))
    np.testing.assert_allclose(a['probs'], b['probs'], atol=1e-5)
    margin = np.sort(b['logits'], -1)
    differ = a['mask'] != b['mask']
    assert (margin[..., -1] - margin[..., -2])[differ].max(initial=0.0) < 1e-4


def test_c_oracle_known_answer(golden_dir):
    g = np.load(os.path.join(golden_dir, 'unet_kat.npz'))
    filters = tuple(int(f) for f in g['filters'])
    w = synth.unet_weights(filters, 1, 2, ndim=2, bridge='concat', seed=5)
    r = unet_c.unet_forward(g['x'], w, filters, 'concat')
    np.testing.assert_array_equal(r['logits'], g['logits'])      # bit-exact, thread-count independent
    np.testing.assert_array_equal(r['mask'], g['mask'])
    rb = unet_c.unet_forward(g['x'], w, filters, 'concat', contract='bf16')
    np.testing.assert_array_equal(rb['logits'], g['logits_bf16'])


def test_thread_count_does_not_change_bits(monkeypatch):
    filters = (8, 16)
    w = synth.unet_weights(filters, 1, 2, seed=1)
    x = np.random.default_rng(0).standard_normal((1, 16, 16, 1)).astype(np.float32)
    a = unet_c.unet_forward(x, w, filters)['logits']
    monkeypatch.setenv('SQREF_THREADS', '1')
    b = unet_c.unet_forward(x, w, filters)['logits']
    np.testing.assert_array_equal(a, b)


def test_bf16_rounding():
    v = np.array([1.0, 1.00390625, 1.005859375, -3.1415927, 0.0, 65504.0], np.float32)
    import torch
    want = torch.from_numpy(v).to(torch.bfloat16).float().numpy()
    np.testing.assert_array_equal(unet_c.round_bf16(v), want)


def test_exact_tail_experiment_contract():
    """oracle-only switch used by scripts/exact_tail_experiment.py: the bf16 contract with the last block and the
    head in full fp32 sits between the bf16 and the fp32 contracts."""
    from oracle import unet_c
    from sequitr_b200 import synth
    filters = (16, 32)
    w = synth.unet_weights(filters, 1, 2, bridge='concat', seed=3)
    x = synth.frames(1, 48, 64, 1, seed=2, n_objects=3)
    f32 = unet_c.unet_forward(x, w, filters, 'concat', contract='fp32')['logits']
    b16 = unet_c.unet_forward(x, w, filters, 'concat', contract='bf16')['logits']
    mix = unet_c.unet_forward(x, w, filters, 'concat', contract='bf16', exact_tail=True)['logits']
    assert 0 < np.abs(mix - f32).mean() < np.abs(b16 - f32).mean()
