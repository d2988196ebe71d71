"""CPU oracle for the per-pixel loss-weight maps (TEST INFRASTRUCTURE ONLY).

Restates, in NumPy/SciPy, the arithmetic of
``/root/reference/sequitr/pipeline.py``:

* ``image_pipe_call``   <- ``ImagePipe.__call__``          pipeline.py:174-180
* ``weightmap_w1``      <- ``ImageWeightMap.pipe``         pipeline.py:470-479
* ``weightmap_w2``      <- ``ImageWeightMap2.pipe/edist``  pipeline.py:509-566
* ``weightmap_w3``      <- north-star formula (not in the reference):
  ``w_c + w0 * exp(-(d1+d2)^2 / (2 sigma^2))`` on instance labels, with the
  reference's envelope (pipeline.py:477-479): ``w0*(1-m)*exp(..) + 1 + m``.
* ``image_labels`` / ``weights_folder_name`` / ``weights_file_name``
                        <- ``weightmap.py:31-73, 188-199``

Pinned against the reference's own code by ``tests/test_oracle_weightmap.py``
(live import when /root/reference exists) and by the committed fixtures in
``tests/golden/`` (made by ``scripts/make_golden.py`` from the reference).
"""
import re
import numpy as np
from scipy.ndimage import (distance_transform_edt, binary_erosion,
                           binary_dilation, gaussian_filter)

INF_D2 = np.iinfo(np.int64).max


def image_pipe_call(image):
    """pipeline.py:174-180 -- 2-D input gains a channel axis and becomes float32."""
    image = np.asarray(image)
    if image.ndim < 3:
        image = image[..., np.newaxis].astype('float32')
    return image


def edt_squared(mask2d):
    """Exact integer squared distance from every pixel to the nearest pixel
    where ``mask2d`` is non-zero (0 on those pixels).  With no such pixel SciPy
    behaves as if one existed at (-1, 0); this oracle records that behaviour
    (distance_transform_edt call at pipeline.py:476)."""
    m = np.asarray(mask2d) != 0
    h, w = m.shape
    if not m.any():
        r = np.arange(h, dtype=np.int64)[:, None] + 1
        c = np.arange(w, dtype=np.int64)[None, :]
        return r * r + c * c
    idx = distance_transform_edt(~m, return_distances=False, return_indices=True)
    r = np.arange(h, dtype=np.int64)[:, None] - idx[0]
    c = np.arange(w, dtype=np.int64)[None, :] - idx[1]
    return r * r + c * c


def weightmap_w1(image, w0=10., sigma=5.):
    """ImageWeightMap (pipeline.py:470-479).  Returns (H,W,1) float64."""
    image = image_pipe_call(image)
    weight_map = distance_transform_edt(1. - image)
    weight_map = w0 * (1. - image) * np.exp(-(weight_map * weight_map) /
                                            (2. * sigma ** 2 + 1e-99))
    return weight_map + image + 1.


def _edist(tri, i):
    """pipeline.py:555-566 -- edge lengths of simplex i (1024 outside the hull)."""
    if i == -1:
        return [1024., 1024., 1024.]
    s = tri.simplices[i]
    p = np.zeros((4, 2))
    p[0:3, :] = tri.points[s]
    p[3, :] = p[0, :]
    d = np.diff(p, axis=0)
    return np.sqrt(d[:, 0] ** 2 + d[:, 1] ** 2)


def weightmap_w2(image, w0=10., sigma=5.):
    """ImageWeightMap2 (pipeline.py:514-553).  Returns (H,W,1) float64."""
    from scipy.spatial import Delaunay
    image = image_pipe_call(image)
    s = np.array([[0, 1, 0], [1, 1, 1], [0, 1, 0]])
    b = np.squeeze(image.astype('bool'))
    b_erode_outline = np.logical_xor(binary_erosion(b, iterations=1, structure=s), b)
    b_dilate = binary_dilation(b, iterations=3, structure=s)
    b_dilate_outline = np.logical_xor(
        binary_erosion(b_dilate, iterations=1, structure=s), b_dilate)
    b_erode = np.logical_xor(b_erode_outline, b_dilate_outline)
    x, y = np.where(b_erode)
    points = np.column_stack((x, y))
    tri = Delaunay(points)
    fx, fy = np.where(np.logical_not(b))
    free_space = np.column_stack((fx, fy))
    simplices = tri.find_simplex(free_space)
    # vectorised max edge length per simplex (pipeline.py:545 takes np.max(edist))
    pts = tri.points[tri.simplices]                      # (n_simplex, 3, 2)
    e = np.sqrt(((pts - np.roll(pts, -1, axis=1)) ** 2).sum(-1)).max(-1)
    vals = np.where(simplices >= 0, e[np.maximum(simplices, 0)], 1024.)
    weight_map = np.zeros(image.shape)
    weight_map[fx, fy, ...] = vals.reshape((-1, 1))
    mask = b[..., np.newaxis].astype('float32')
    weight_map = gaussian_filter(weight_map, 1.)
    weight_map = w0 * (1. - mask) * np.exp(-(weight_map * weight_map) /
                                           (2. * sigma ** 2 + 1e-99))
    return weight_map + 1. + mask


def two_nearest_instances_d2(labels):
    """For every pixel: exact integer squared distances to the nearest and to
    the second-nearest *distinct* instance (label > 0) -- brute force over
    instances with one exact EDT each.  INF_D2 where fewer than 1 / 2 exist."""
    labels = np.asarray(labels)
    ids = [int(i) for i in np.unique(labels) if i > 0]
    h, w = labels.shape
    d1 = np.full((h, w), INF_D2, dtype=np.int64)
    d2 = np.full((h, w), INF_D2, dtype=np.int64)
    for i in ids:
        d = edt_squared(labels == i)
        lt1 = d < d1
        d2 = np.where(lt1, d1, np.minimum(d2, d))
        d1 = np.where(lt1, d, d1)
    return d1, d2


def two_nearest_instances_d2_cropped(labels, margin):
    """The same two squared distances, exact wherever they are <= margin^2 and INF_D2 (or larger than margin^2)
    elsewhere, at a cost that allows BASELINE's 2048^2 frames with ~600 instances: the EDT of instance i is only
    taken on its bounding box grown by `margin` pixels (every pixel within `margin` of the instance lies in that
    crop, and its nearest pixel of the instance lies in the bounding box, so the cropped transform is exact there)."""
    from scipy.ndimage import find_objects
    labels = np.asarray(labels)
    h, w = labels.shape
    d1 = np.full((h, w), INF_D2, dtype=np.int64)
    d2 = np.full((h, w), INF_D2, dtype=np.int64)
    lim = int(margin) * int(margin)
    for i, sl in enumerate(find_objects(labels.astype(np.int64)), start=1):
        if sl is None:
            continue
        y0, y1 = max(sl[0].start - margin, 0), min(sl[0].stop + margin, h)
        x0, x1 = max(sl[1].start - margin, 0), min(sl[1].stop + margin, w)
        d = edt_squared(labels[y0:y1, x0:x1] == i)
        d = np.where(d <= lim, d, INF_D2)                       # beyond the margin the crop says nothing
        a, b = d1[y0:y1, x0:x1], d2[y0:y1, x0:x1]
        lt1 = d < a
        d2[y0:y1, x0:x1] = np.where(lt1, a, np.minimum(b, d))
        d1[y0:y1, x0:x1] = np.where(lt1, d, a)
    return d1, d2


def weightmap_w3(labels, w0=10., sigma=5., wc=None, dtype=np.float64, margin=None):
    """North-star U-Net weight map on an int instance-label image (H,W):
    ``w = w0*(1-m)*exp(-(d1+d2)^2/(2 sigma^2 + 1e-99)) + wc[m]`` with
    ``m = labels > 0`` and ``wc = (1, 2)`` by default (the reference's
    ``+ 1 + m`` class term, pipeline.py:479).  Returns (H,W) float64.
    ``margin``: use the per-instance cropped transforms (distances beyond `margin` pixels count as infinite:
    pick it so that w0*exp(-margin^2/(2 sigma^2)) is far below one ulp of the class weight)."""
    labels = np.asarray(labels)
    m = (labels > 0)
    d1sq, d2sq = two_nearest_instances_d2(labels) if margin is None else two_nearest_instances_d2_cropped(labels, margin)
    wc0, wc1 = (1., 2.) if wc is None else (float(wc[0]), float(wc[1]))
    with np.errstate(over='ignore', invalid='ignore'):
        d1 = np.where(d1sq == INF_D2, np.inf, np.sqrt(d1sq.astype(np.float64)))
        d2 = np.where(d2sq == INF_D2, np.inf, np.sqrt(d2sq.astype(np.float64)))
        s = d1 + d2
        term = w0 * np.exp(-(s * s) / (2. * sigma ** 2 + 1e-99))
    out = np.where(m, wc1, wc0 + term)
    return out.astype(dtype)


# ---------------------------------------------------------------- weightmap.py

def image_labels(raw, thresh_fn=lambda x: x > 0):
    """ImageLabels.__init__ (weightmap.py:36-63) on an in-memory array.
    Returns (labels uint8 (H,W), n_outputs)."""
    raw = np.asarray(raw)
    assert raw.ndim > 1 and raw.ndim < 4
    if raw.ndim == 3:
        l_data = np.zeros(raw.shape[1:], dtype='uint8')
        for l in range(raw.shape[0]):
            l_data[thresh_fn(raw[l, ...])] = l + 1
        outputs = raw.shape[0] + 1
    else:
        l_data = thresh_fn(raw).astype('uint8')
        outputs = 2
    if outputs > 5:
        raise ValueError('More that five output classes!')
    return l_data, outputs


def weights_folder_name(w0, sigma, name_weights_folder=True):
    """weightmap.py:186-190."""
    base = 'weights'
    if name_weights_folder:
        base += '_w0-{0:2.2f}_sigma-{1:2.2f}'.format(w0, sigma)
    return base


def weights_file_name(label_file):
    """weightmap.py:198-199."""
    return re.match('([a-zA-Z0-9()]+)_([a-zA-Z0-9()]+_)*', label_file).group(0) + 'weights.tif'
