"""Pin the weighted cross-entropy oracle to torch's own implementation (CPU)."""
import numpy as np
import torch
import torch.nn.functional as F

from oracle import loss_oracle


def test_weighted_ce_oracle_matches_torch():
    rng = np.random.default_rng(0)
    logits = rng.standard_normal((3, 7, 5, 4)) * 2
    labels = rng.integers(0, 4, (3, 7, 5)).astype(np.uint8)
    weights = rng.uniform(1, 11, (3, 7, 5))
    loss, grad = loss_oracle.weighted_ce(logits, labels, weights)
    t = torch.tensor(logits, requires_grad=True)
    ce = F.cross_entropy(t.reshape(-1, 4), torch.tensor(labels.reshape(-1).astype(np.int64)), reduction='none')
    ref = (ce * torch.tensor(weights.reshape(-1))).sum() / ce.numel()
    ref.backward()
    assert abs(loss - float(ref)) < 1e-12 * abs(float(ref))
    np.testing.assert_allclose(grad, t.grad.numpy(), atol=1e-15)
