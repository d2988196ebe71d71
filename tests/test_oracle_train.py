"""Pins of oracle/train_oracle.py (CPU): the float64 autograd restatement against central finite differences of its
own loss, known-answer values of the dropout hash, the Adam rule against a hand computation."""
import numpy as np
import pytest

from oracle import train_oracle
from sequitr_b200 import synth


def _batch(ndim, n, size, cin, k, seed):
    rng = np.random.default_rng(seed)
    sp = (n,) + (size,) * ndim
    image = rng.standard_normal(sp + (cin,)).astype(np.float32)
    labels = rng.integers(0, k, sp).astype(np.uint8)
    wmap = (1.0 + 9.0 * rng.random(sp)).astype(np.float32)
    return image, labels, wmap


def test_mix64_known_answers():
    # splitmix64 reference stream for seed 0: the finaliser applied to 1, 2, 3 times the golden gamma
    g = 0x9E3779B97F4A7C15
    assert train_oracle._mix64_int(g) == 0xE220A8397B1DCDAF
    assert train_oracle._mix64_int(2 * g) == 0x6E789E6AA1B965F4
    assert train_oracle._mix64_int(3 * g) == 0x06C45D188009454F
    z = np.array([g, (2 * g) & train_oracle._M64], dtype=np.uint64)
    assert [int(v) for v in train_oracle._mix64_np(z)] == [0xE220A8397B1DCDAF, 0x6E789E6AA1B965F4]


def test_dropout_keep_rate_and_determinism():
    keep = train_oracle.dropout_keep(train_oracle.step_seed(7, 3), 2, 200000, 0.4)
    assert abs(keep.mean() - 0.6) < 5e-3
    again = train_oracle.dropout_keep(train_oracle.step_seed(7, 3), 2, 200000, 0.4)
    assert (keep == again).all()
    other = train_oracle.dropout_keep(train_oracle.step_seed(7, 4), 2, 200000, 0.4)
    assert 0.4 < (keep == other).mean() < 0.6            # 0.6^2 + 0.4^2 = 0.52 for independent masks
    assert train_oracle.dropout_keep(1, 0, 1000, 0.0).all()


@pytest.mark.parametrize('ndim,bridge', [(2, 'concat'), (2, 'eltwise_mul'), (3, 'concat')])
def test_autograd_against_finite_differences(ndim, bridge):
    filters = (3, 4)
    w = synth.unet_weights(filters, 2, 3, ndim=ndim, bridge=bridge, seed=5)
    image, labels, wmap = _batch(ndim, 2, 4, 2, 3, seed=ndim)
    loss, grads, _ = train_oracle.gradients(w, image, labels, wmap, filters, bridge, ndim, rate=0.3, seed=11, step=2)
    rng = np.random.default_rng(0)
    for name in ('UNet/down0/conv1/kernel', 'UNet/down1/conv2/bias', 'UNet/up0/upscale/kernel',
                 'UNet/up0/conv1/kernel', 'UNet/to_image/kernel', 'UNet/up0/upscale/bias'):
        for _ in range(3):
            idx = tuple(int(rng.integers(0, s)) for s in w[name].shape)
            h = 1e-5
            vals = []
            for sgn in (+1, -1):
                w2 = {k: v.astype(np.float64).copy() for k, v in w.items()}
                w2[name][idx] += sgn * h
                vals.append(train_oracle.gradients(w2, image, labels, wmap, filters, bridge, ndim, rate=0.3, seed=11,
                                                   step=2)[0])
            fd = (vals[0] - vals[1]) / (2 * h)
            assert abs(fd - grads[name][idx]) <= 1e-6 + 1e-5 * abs(fd), (name, idx, fd, grads[name][idx])


def test_frozen_affine_layers_against_finite_differences():
    """conv + bias + frozen per-channel affine + ReLU: gradients only for kernels and biases, checked by finite
    differences; scale / shift get none."""
    filters = (3, 4)
    w = synth.unet_weights(filters, 2, 3, ndim=2, bridge='concat', seed=5, affine=True)
    image, labels, wmap = _batch(2, 2, 4, 2, 3, seed=7)
    loss, grads, _ = train_oracle.gradients(w, image, labels, wmap, filters, 'concat', 2, rate=0.3, seed=1, step=0)
    assert any(n.endswith('/scale') for n in w) and all(n.endswith(('/kernel', '/bias')) for n in grads)
    rng = np.random.default_rng(1)
    for name in ('UNet/down0/conv1/bias', 'UNet/up0/conv1/kernel', 'UNet/down1/conv2/kernel'):
        idx = tuple(int(rng.integers(0, s)) for s in w[name].shape)
        h = 1e-5
        vals = []
        for sgn in (+1, -1):
            w2 = {k: v.astype(np.float64).copy() for k, v in w.items()}
            w2[name][idx] += sgn * h
            vals.append(train_oracle.gradients(w2, image, labels, wmap, filters, 'concat', 2, rate=0.3, seed=1,
                                               step=0)[0])
        fd = (vals[0] - vals[1]) / (2 * h)
        assert abs(fd - grads[name][idx]) <= 1e-6 + 1e-5 * abs(fd), (name, idx, fd, grads[name][idx])


def test_adam_rule_by_hand():
    opt = train_oracle.Adam({'a': np.array([1.0, -2.0])}, learning_rate=0.1, beta1=0.9, beta2=0.999, epsilon=1e-8)
    g = np.array([0.5, -0.25])
    opt.apply({'a': g})
    # first step of Adam: m_hat = g, v_hat = g^2 -> the step is lr * sign(g) (up to eps)
    np.testing.assert_allclose(opt.w['a'], np.array([1.0 - 0.1, -2.0 + 0.1]), rtol=0, atol=2e-7)
    opt.apply({'a': g})
    np.testing.assert_allclose(opt.w['a'], np.array([1.0 - 0.2, -2.0 + 0.2]), rtol=0, atol=5e-7)
    sgd = train_oracle.Adam({'a': np.array([1.0])}, learning_rate=0.5, optimizer='sgd')
    sgd.apply({'a': np.array([2.0])})
    assert sgd.w['a'][0] == 0.0
