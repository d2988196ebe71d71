"""CPU oracle of the weighted softmax cross-entropy (TEST INFRASTRUCTURE ONLY).

The reference prepares one-hot labels and a per-pixel 'weights' map in ``tr_augment``
(/root/reference/sequitr/networks/unet.py:396-401) but does not ship the loss that consumes
them ("parity unpinned"); this is the definition in include/sequitr_b200.h in float64 NumPy."""
import numpy as np


def weighted_ce(logits, labels, weights):
    """Returns (loss, grad) with loss = mean_i w_i * CE_i and grad = w_i (softmax - onehot) / N."""
    l = np.asarray(logits, np.float64)
    k = l.shape[-1]
    l2 = l.reshape(-1, k)
    y = np.asarray(labels).reshape(-1).astype(np.int64)
    w = np.asarray(weights, np.float64).reshape(-1)
    m = l2.max(-1, keepdims=True)
    e = np.exp(l2 - m)
    s = e.sum(-1, keepdims=True)
    lse = (m + np.log(s))[:, 0]
    n = l2.shape[0]
    loss = float((w * (lse - l2[np.arange(n), y])).sum() / n)
    p = e / s
    p[np.arange(n), y] -= 1.0
    grad = (p * w[:, None] / n).reshape(l.shape)
    return loss, grad
