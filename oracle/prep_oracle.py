"""CPU restatement of the reference's pre-inference clean-up pipes (TEST INFRASTRUCTURE ONLY).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU baseline may import this module;
the product path (``sequitr_b200``) never does.

Each function follows the reference's ``pipeline.py`` line by line (NumPy / SciPy, the reference's own
dependencies) and is pinned by ``tests/golden/prep_ref.npz`` -- outputs of the REFERENCE classes
themselves (imported through ``oracle/ref_loader.py`` by ``scripts/make_golden.py``) -- and by a live
comparison with the reference where ``/root/reference`` exists.
"""
import numpy as np
from scipy.ndimage import median_filter


def promote(image):
    """ImagePipe.__call__ (reference pipeline.py:174-180): 2-D input becomes (H,W,1) float32."""
    image = np.asarray(image)
    if image.ndim == 2:
        image = image[..., np.newaxis].astype('float32')
    return image


def image_norm(image):
    """ImageNorm.pipe (pipeline.py:352-356): per channel (x - mean) / (1e-99 + std), in the image's own
    dtype (float32: the epsilon underflows to zero)."""
    image = promote(image).copy()
    for chnl in range(image.shape[-1]):
        image[..., chnl] = (image[..., chnl] - np.mean(image[..., chnl])) / \
            (1e-99 + np.std(image[..., chnl]))
    return image


def image_outliers(image, sigma=2, threshold=5.):
    """ImageOutliers.pipe (pipeline.py:287-296): pixels further than ``threshold`` from the
    ``sigma``-sized median take the median value."""
    image = promote(image).copy()
    for chnl in range(image.shape[-1]):
        filtered = image[..., chnl].copy()
        med = median_filter(filtered, sigma)
        differences = np.abs(image[..., chnl] - med) > threshold
        filtered[differences] = med[differences]
        image[..., chnl] = filtered
    return image


def image_bgsubtract(image):
    """ImageBGSubtract.pipe (pipeline.py:384-405): least-squares second-order surface over
    (u = column, v = row), subtracted; float64 result (H,W,1).  Solved with ``lstsq`` on centred, scaled
    coordinates -- the same surface the reference's ``inv(A.T*A)*A.T`` yields, without its conditioning
    (the two agree to ~1e-12 on 64^2..1024^2 images, tests/test_oracle_prep.py)."""
    image = promote(image)
    h, w = image.shape[0], image.shape[1]
    u, v = np.meshgrid(np.arange(0, w), np.arange(0, h))
    s = (u - 0.5 * (w - 1)) / (0.5 * w)
    t = (v - 0.5 * (h - 1)) / (0.5 * h)
    A = np.stack([np.ones_like(s), s, t, s * s, s * t, t * t], -1).reshape(-1, 6)
    k, *_ = np.linalg.lstsq(A, np.ravel(image).astype(np.float64), rcond=None)
    background = (A @ k).reshape(h, w)
    return image - background[..., np.newaxis]
