"""Training step of the UNet on the device (BASELINE config 5; reference networks/unet.py:224-277 in TRAIN mode fed by
tr_augment :348-401) against the float64 autograd oracle (oracle/train_oracle.py): loss, every gradient, the dropout
mask, the Adam / SGD update, and inference with the updated weights.

Tolerance: the device computes in fp32 with fixed-order split sums, the oracle in float64 -- gradients agree to
2e-4 of the tensor's largest magnitude (plus 1e-7 absolute), the loss to 1e-5 relative."""
import numpy as np
import pytest

from oracle import train_oracle
from sequitr_b200 import synth

pytestmark = pytest.mark.gpu


def _batch(ndim, n, shape, cin, k, seed):
    rng = np.random.default_rng(seed)
    sp = (n,) + tuple(shape)
    image = rng.standard_normal(sp + (cin,)).astype(np.float32)
    labels = rng.integers(0, k, sp).astype(np.uint8)
    wmap = (1.0 + 9.0 * rng.random(sp)).astype(np.float32)
    return image, labels, wmap


def _net(ndim, filters, cin, k, bridge, shape, weights, dropout):
    from sequitr_b200.networks import UNet2D, UNet3D
    from sequitr_b200.networks.unet import ModeKeys
    cls = UNet2D if ndim == 2 else UNet3D
    net = cls({'filters': filters, 'shape': tuple(shape), 'bridge': bridge, 'num_inputs': cin, 'num_outputs': k,
               'compute': 'fp32', 'dropout': dropout}, mode=ModeKeys.TRAIN)
    net.load_weights(weights)
    return net


def _assert_grads(got, ref, tol=2e-4):
    assert set(got) == set(ref) and all(n.endswith(('/kernel', '/bias')) for n in got)
    for name in sorted(ref):
        scale = np.abs(ref[name]).max()
        err = np.abs(got[name].astype(np.float64) - ref[name]).max()
        assert err <= tol * scale + 1e-7, (name, err, scale)


@pytest.mark.parametrize('ndim,bridge,filters,shape,rate', [
    (2, 'concat', (4, 8, 16), (16, 24), 0.0),
    (2, 'concat', (4, 8, 16), (16, 24), 0.4),
    (2, 'eltwise_mul', (4, 8), (8, 12), 0.4),        # the reference's default bridge (networks/unet.py:138)
    (2, 'eltwise_add', (4, 8), (8, 12), 0.0),
    (2, 'eltwise_sub', (4, 8), (8, 12), 0.25),
    (2, None, (4, 8), (8, 12), 0.0),
    (2, 'concat', (16, 32, 64), (32, 32), 0.4),      # the first three widths of the default filters
    (2, 'concat', (5,), (6, 10), 0.4),               # a single level: no pooling, no up block
    (3, 'concat', (4, 8), (8, 8, 12), 0.4),
    (3, 'eltwise_mul', (3, 6, 12), (8, 8, 8), 0.0),
])
def test_loss_and_gradients_against_the_autograd_oracle(sq, ndim, bridge, filters, shape, rate):
    cin, k = (3, 3) if bridge == 'concat' else (1, 2)
    w = synth.unet_weights(filters, cin, k, ndim=ndim, bridge=bridge or 'none', seed=3)
    image, labels, wmap = _batch(ndim, 2, shape, cin, k, seed=len(filters) + ndim)
    net = _net(ndim, filters, cin, k, bridge, shape, w, rate)
    tr = net.trainer(seed=1234)
    loss = tr.step(image, labels, wmap, apply_update=False)
    ref_loss, ref_grads, _ = train_oracle.gradients(w, image, labels, wmap, filters, bridge, ndim, rate=rate,
                                                    seed=1234, step=0)
    assert abs(loss - ref_loss) <= 1e-5 * abs(ref_loss)
    _assert_grads(tr.gradients(), ref_grads)
    # gradients only: the weights are untouched, and the step is deterministic
    for name, arr in tr.weights().items():
        assert (arr == w[name]).all(), name
    assert tr.step(image, labels, wmap, apply_update=False) == loss
    tr.close()


@pytest.mark.parametrize('ndim,shape', [(2, (16, 24)), (3, (8, 8, 8))])
def test_layers_with_a_frozen_affine(sq, ndim, shape):
    """conv_layer = conv + bias + optional per-channel affine (folded BN) + ReLU: kernels and
    biases train, scale / shift stay frozen; the plan's folded epilogue shift follows the bias after every update."""
    filters, cin, k, rate = (4, 8), 2, 3, 0.25
    w = synth.unet_weights(filters, cin, k, ndim=ndim, bridge='concat', seed=6, affine=True)
    assert any(n.endswith('/scale') for n in w)
    image, labels, wmap = _batch(ndim, 2, shape, cin, k, seed=9)
    net = _net(ndim, filters, cin, k, 'concat', shape, w, rate)
    tr = net.trainer(learning_rate=0.05, optimizer='sgd', seed=3)
    loss = tr.step(image, labels, wmap, apply_update=False)
    ref_loss, ref_grads, _ = train_oracle.gradients(w, image, labels, wmap, filters, 'concat', ndim, rate=rate, seed=3,
                                                    step=0)
    assert abs(loss - ref_loss) <= 1e-5 * abs(ref_loss)
    _assert_grads(tr.gradients(), ref_grads)
    opt = train_oracle.Adam(w, learning_rate=0.05, optimizer='sgd')
    for step in range(2):
        loss = tr.step(image, labels, wmap)
        ref_loss, grads, _ = train_oracle.gradients(opt.w, image, labels, wmap, filters, 'concat', ndim, rate=rate,
                                                    seed=3, step=step)
        assert abs(loss - ref_loss) <= 2e-5 * abs(ref_loss), step
        opt.apply(grads)
    got = tr.weights()
    assert set(got) == set(w)
    for name in sorted(got):
        if name.endswith(('/scale', '/shift')):
            assert (got[name] == w[name]).all(), name                      # frozen
        else:
            np.testing.assert_allclose(got[name], opt.w[name], rtol=0, atol=2e-6, err_msg=name)
    # inference on the same plan uses the re-folded shift
    logits = net.predict(image, want=('logits',))['logits']
    _, _, ref_logits = train_oracle.gradients(opt.w, image, labels, wmap, filters, 'concat', ndim)
    np.testing.assert_allclose(logits, ref_logits, rtol=0, atol=2e-4)
    tr.close()


@pytest.mark.parametrize('optimizer', ['adam', 'sgd'])
def test_three_steps_follow_the_oracle_and_inference_sees_the_update(sq, optimizer):
    filters, shape, cin, k, rate = (4, 8, 16), (16, 16), 1, 2, 0.4
    w = synth.unet_weights(filters, cin, k, ndim=2, bridge='concat', seed=8)
    net = _net(2, filters, cin, k, 'concat', shape, w, rate)
    lr = 1e-2 if optimizer == 'adam' else 0.05
    tr = net.trainer(learning_rate=lr, optimizer=optimizer, seed=77)
    opt = train_oracle.Adam(w, learning_rate=lr, optimizer=optimizer)
    for step in range(3):
        image, labels, wmap = _batch(2, 2, shape, cin, k, seed=100 + step)
        loss = tr.step(image, labels, wmap)
        ref_loss, grads, _ = train_oracle.gradients(opt.w, image, labels, wmap, filters, 'concat', 2, rate=rate,
                                                    seed=77, step=step)
        assert abs(loss - ref_loss) <= 2e-5 * abs(ref_loss), step
        opt.apply(grads)
    got = tr.weights()
    for name in sorted(got):
        # Adam's first steps move every weight by ~lr whatever the gradient's size: elements whose gradient is at
        # rounding level may differ in the sign of the step, so the bound is relative to the step, not the weight
        np.testing.assert_allclose(got[name], opt.w[name], rtol=0, atol=(0.05 * lr if optimizer == 'adam' else 1e-6),
                                   err_msg=name)
    # the same plan serves inference with the trained weights (dropout off outside the step)
    image, _, _ = _batch(2, 2, shape, cin, k, seed=5)
    logits = net.predict(image, want=('logits',))['logits']
    _, _, ref_logits = train_oracle.gradients(opt.w, image, np.zeros((2,) + shape, np.uint8),
                                              np.ones((2,) + shape, np.float32), filters, 'concat', 2)
    np.testing.assert_allclose(logits, ref_logits, rtol=0, atol=(2e-2 if optimizer == 'adam' else 1e-4))


def test_loss_goes_down_on_a_fixed_batch_and_weights_load_into_a_bf16_network(sq):
    from sequitr_b200.networks import UNet2D
    filters, shape = (16, 32), (64, 64)
    scene = synth.frames(2, shape[0], shape[1], seed=3, n_objects=6)
    image = np.asarray(scene, dtype=np.float32).reshape((2,) + shape + (1,))
    labels = (image[..., 0] > image.mean()).astype(np.uint8)
    wmap = np.where(labels > 0, 2.0, 1.0).astype(np.float32)
    w = synth.unet_weights(filters, 1, 2, ndim=2, bridge='concat', seed=2)
    net = _net(2, filters, 1, 2, 'concat', shape, w, 0.0)
    tr = net.trainer(learning_rate=3e-3)
    losses = [tr.step(image, {'label': np.eye(2, dtype=np.uint8)[labels], 'weights': wmap[..., None]})
              for _ in range(40)]
    assert losses[-1] < 0.5 * losses[0], losses[::8]
    mask = net.predict(image, want=('mask',))['mask']
    assert (mask == labels).mean() > 0.9
    fast = UNet2D({'filters': filters, 'shape': shape, 'bridge': 'concat', 'num_inputs': 1, 'num_outputs': 2,
                   'compute': 'bf16'})
    fast.load_weights(tr.weights())
    assert (fast.predict(image, want=('mask',))['mask'] == mask).mean() > 0.98


def test_trainer_argument_errors(sq):
    from sequitr_b200.networks import UNet2D
    w = synth.unet_weights((4, 8), 1, 2, ndim=2, bridge='concat', seed=1)
    bf = UNet2D({'filters': (4, 8), 'shape': (8, 8), 'bridge': 'concat', 'compute': 'bf16'})
    bf.load_weights(w)
    with pytest.raises(ValueError):
        bf.trainer()
    net = _net(2, (4, 8), 1, 2, 'concat', (8, 8), w, 0.0)
    tr = net.trainer()
    image, labels, wmap = _batch(2, 1, (8, 8), 1, 2, seed=0)
    with pytest.raises(ValueError):
        tr.step(image, labels + 2, wmap)                 # class id beyond num_outputs
    with pytest.raises(ValueError):
        tr.step(image[:, :7], labels, wmap)              # weights of another shape
    with pytest.raises(ValueError):
        tr.step(image[:, :7], labels[:, :7], wmap[:, :7])    # 7 rows: not divisible by 2 (utils.py:234-240)
    net.load_weights(w)
    with pytest.raises(RuntimeError):
        tr.step(image, labels, wmap)                     # the plan the trainer was bound to is gone


def _dp_worker(rank, world, port, out_dir):
    import os
    import sys
    import torch
    import torch.distributed as dist
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    filters, shape, cin, k = (8, 16, 32), (32, 32), 1, 2
    w = synth.unet_weights(filters, cin, k, ndim=2, bridge='concat', seed=4)
    image, labels, wmap = _batch(2, 4, shape, cin, k, seed=21)
    net = _net(2, filters, cin, k, 'concat', shape, w, 0.0)
    tr = net.trainer(learning_rate=0.05, optimizer='sgd')
    share = slice(rank * 2, rank * 2 + 2)                      # this rank's half of the batch
    losses = [tr.step(image[share], labels[share], wmap[share]) for _ in range(2)]
    got = tr.weights()
    if rank == 0:
        # the same two steps on the whole batch in one process
        ref_net = _net(2, filters, cin, k, 'concat', shape, w, 0.0)
        ref = ref_net.trainer(learning_rate=0.05, optimizer='sgd', data_parallel=False)
        ref_losses = [ref.step(image, labels, wmap) for _ in range(2)]
        want = ref.weights()
        err = max(float(np.abs(got[n] - want[n]).max()) for n in want)
        np.save(os.path.join(out_dir, 'dp.npy'), np.array([err] + losses + ref_losses))
    dist.barrier()
    dist.destroy_process_group()


def test_data_parallel_step_equals_the_whole_batch_step(sq, tmp_path):
    """Two ranks, half a batch each, ONE all-reduce of the gradient arena per step (NCCL): the weights after two SGD
    steps equal those of a single process stepping on the whole batch (up to the fp32 order of the sums)."""
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip('needs two GPUs (gpurun --gpus 2)')
    mp.spawn(_dp_worker, args=(2, 29671, str(tmp_path)), nprocs=2, join=True)
    res = np.load(str(tmp_path / 'dp.npy'))
    assert res[0] <= 2e-6, res
    np.testing.assert_allclose(res[1:3], res[3:5], rtol=1e-6)
