"""Weighted cross-entropy step (config 5): UNet logits + GPU weight map -> loss and head gradient."""
import numpy as np
import pytest

from oracle import loss_oracle, weightmap_oracle
from sequitr_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('k', [2, 3, 5])
def test_weighted_ce_matches_oracle(sq, k):
    import torch
    from sequitr_b200 import ops
    rng = np.random.default_rng(k)
    lab = synth.instance_labels(96, 128, 9, seed=k, rmin=4, rmax=9)
    labels = np.where(lab > 0, 1 + (lab - 1) % (k - 1), 0).astype(np.uint8)
    logits = (rng.standard_normal((96, 128, k)) * 3).astype(np.float32)
    weights = ops.weightmap_unet(torch.from_numpy(lab.astype(np.int32))[None].cuda(), 10., 5.)[0]
    np.testing.assert_allclose(weights.cpu().numpy(),
                               weightmap_oracle.weightmap_w3(lab, 10., 5.).astype(np.float32), rtol=1.2e-7)
    loss, grad = ops.weighted_cross_entropy(torch.from_numpy(logits).cuda(), torch.from_numpy(labels).cuda(),
                                            weights.contiguous())
    ref_loss, ref_grad = loss_oracle.weighted_ce(logits, labels, weights.cpu().numpy())
    assert abs(float(loss) - ref_loss) <= 2e-6 * abs(ref_loss)
    np.testing.assert_allclose(grad.cpu().numpy(), ref_grad, atol=2e-6 * np.abs(ref_grad).max())
    # deterministic reduction
    loss2, _ = ops.weighted_cross_entropy(torch.from_numpy(logits).cuda(), torch.from_numpy(labels).cuda(),
                                          weights.contiguous(), want_grad=False)
    assert float(loss2) == float(loss)


def test_unet3d_training_step_pieces(sq):
    """Config 5 at a test size: UNet3D forward (fp32 exact) -> weighted CE with a weight volume."""
    import torch
    from oracle import unet_c
    from sequitr_b200 import ops
    from sequitr_b200.networks import UNet3D
    filters = (8, 16)
    w = synth.unet_weights(filters, 1, 2, ndim=3, bridge='concat', seed=9)
    x = synth.volumes(1, 8, 32, 32, 1)
    net = UNet3D({'filters': filters, 'shape': (32, 32, 8), 'bridge': 'concat', 'compute': 'fp32'})
    net.load_weights(w)
    out = net.predict(torch.from_numpy(x).cuda(), want=('logits',))
    ref = unet_c.unet_forward(x, w, filters, 'concat')
    np.testing.assert_array_equal(out['logits'].cpu().numpy(), ref['logits'])
    labels = (ref['logits'][..., 1] > ref['logits'][..., 0]).astype(np.uint8)
    wmap = np.stack([weightmap_oracle.weightmap_w1(labels[0, z], 10., 5.)[..., 0] for z in range(8)])[None]
    loss, grad = ops.weighted_cross_entropy(out['logits'], torch.from_numpy(labels).cuda(),
                                            torch.from_numpy(wmap.astype(np.float32)).cuda())
    ref_loss, ref_grad = loss_oracle.weighted_ce(ref['logits'], labels, wmap.astype(np.float32))
    assert abs(float(loss) - ref_loss) <= 2e-6 * abs(ref_loss)
    np.testing.assert_allclose(grad.cpu().numpy(), ref_grad, atol=2e-6 * np.abs(ref_grad).max())
