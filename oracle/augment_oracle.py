"""CPU oracle for ``tr_augment`` (reference ``networks/unet.py:348-401``).  TEST INFRASTRUCTURE ONLY:
imported by ``tests/`` (and nothing in the product path).

The arithmetic of the reference lives in a third-party dependency that is absent here:
``tf.contrib.image.rotate`` / ``tf.contrib.image.transform`` of TensorFlow 1.x (the reference pins
no version -- README.md:25 says "Tensorflow"; its API usage implies 1.12 <= TF < 2.0).  This file
restates that published algorithm in NumPy float32:

* ``angles_to_projective_transforms`` (tensorflow/contrib/image/python/ops/image_ops.py):
  ``x_off = ((W-1) - (cos*(W-1) - sin*(H-1))) / 2``, ``y_off = ((H-1) - (sin*(W-1) + cos*(H-1))) / 2``,
  transform ``[cos, -sin, x_off, sin, cos, y_off, 0, 0]`` maps OUTPUT (x, y) to the INPUT sample point;
* ``ProjectiveGenerator`` (tensorflow/contrib/image/kernels/image_ops.h): NEAREST reads
  ``input[round(y), round(x)]`` (``std::round``, half away from zero), BILINEAR combines the four
  neighbours ``floor``/``floor+1`` each read with fill value 0 outside the image, in the order
  ``(x_ceil-x)*f(x_floor) + (x-x_floor)*f(x_ceil)`` per row, then the same over rows.

**Parity unpinned** by the reference (no TensorFlow, no fixtures); the restatement is cross-checked in
``tests/test_oracle_augment.py`` against two independent implementations of zero-filled bilinear /
nearest resampling (``scipy.ndimage.map_coordinates(mode='grid-constant')`` and
``torch.nn.functional.grid_sample(padding_mode='zeros', align_corners=True)``).

Reference quirks kept on purpose (``networks/unet.py``):
* ``:386`` ``if (ch,cw != height,width)`` is a non-empty tuple, i.e. always true: the crop always runs;
* ``:380-383`` the out-of-frame mask uses NEAREST while the weights use BILINEAR, so the one-pixel rim
  of the rotated frame carries partially faded weights without the +1;
* ``:396-398`` the label is expanded over ``range(5)`` and cut to ``num_outputs`` channels.
"""
import numpy as np

f32 = np.float32


def rotation_transform(theta, height, width):
    """angles_to_projective_transforms for one angle -> float32 (a0, a1, a2, b0, b1, b2)."""
    theta = f32(theta)
    c, s = np.cos(theta, dtype=f32), np.sin(theta, dtype=f32)
    wm1, hm1 = f32(width) - f32(1), f32(height) - f32(1)
    x_off = (wm1 - (c * wm1 - s * hm1)) / f32(2)
    y_off = (hm1 - (s * wm1 + c * hm1)) / f32(2)
    return np.array([c, -s, x_off, s, c, y_off], dtype=f32)


def sample_points(transform, ys, xs):
    """Input sample point of every output pixel: float32, products and sums rounded one by one."""
    t = np.asarray(transform, dtype=f32)
    x = xs.astype(f32)
    y = ys.astype(f32)
    ix = (t[0] * x + t[1] * y) + t[2]
    iy = (t[3] * x + t[4] * y) + t[5]
    return iy.astype(f32), ix.astype(f32)


def _read(img, y, x):
    h, w = img.shape[:2]
    ok = (y >= 0) & (y < h) & (x >= 0) & (x < w)
    yy = np.clip(y, 0, h - 1)
    xx = np.clip(x, 0, w - 1)
    v = img[yy, xx].astype(f32)
    return np.where(ok if v.ndim == ok.ndim else ok[..., None], v, f32(0))


def bilinear(img, iy, ix):
    yf, xf = np.floor(iy), np.floor(ix)
    yc, xc = yf + f32(1), xf + f32(1)
    y0 = np.clip(yf, -2, img.shape[0] + 1).astype(np.int64)
    x0 = np.clip(xf, -2, img.shape[1] + 1).astype(np.int64)
    ex = (lambda a: a[..., None]) if img.ndim == 3 else (lambda a: a)
    v_floor = ex(xc - ix) * _read(img, y0, x0) + ex(ix - xf) * _read(img, y0, x0 + 1)
    v_ceil = ex(xc - ix) * _read(img, y0 + 1, x0) + ex(ix - xf) * _read(img, y0 + 1, x0 + 1)
    return (ex(yc - iy) * v_floor + ex(iy - yf) * v_ceil).astype(f32)


def _round_half_away(v):
    return np.where(v >= 0, np.floor(v + f32(0.5)), np.ceil(v - f32(0.5)))


def nearest(img, iy, ix, want_inside=False):
    # std::round; v + 0.5 in float32 can itself round up for the largest float below .5, so use
    # the exact definition on float64 copies
    ry = _round_half_away(iy.astype(np.float64)).astype(np.int64)
    rx = _round_half_away(ix.astype(np.float64)).astype(np.int64)
    h, w = img.shape[:2]
    inside = (ry >= 0) & (ry < h) & (rx >= 0) & (rx < w)
    out = np.where(inside, img[np.clip(ry, 0, h - 1), np.clip(rx, 0, w - 1)], 0).astype(img.dtype)
    return (out, inside) if want_inside else out


def tr_augment(image, label, weights, theta, rh, rw, ch, cw, num_outputs=2):
    """One example.  image (H,W,C) float32, label (H,W) uint8, weights (H,W) float32 ->
    (image (ch,cw,C) float32, label (ch,cw,num_outputs) uint8 one-hot, weights (ch,cw) float32)."""
    image = np.asarray(image, dtype=f32)
    h, w = image.shape[:2]
    t = rotation_transform(theta, h, w)
    ys, xs = np.meshgrid(np.arange(rh, rh + ch), np.arange(rw, rw + cw), indexing='ij')
    iy, ix = sample_points(t, ys, xs)
    im = bilinear(image, iy, ix)
    lab, inside = nearest(np.asarray(label, dtype=np.uint8), iy, ix, want_inside=True)
    wg = bilinear(np.asarray(weights, dtype=f32), iy, ix) + np.where(inside, f32(0), f32(1))
    onehot = np.stack([(lab == k).astype(np.uint8) for k in range(5)], axis=-1)[..., :num_outputs]
    return im, onehot, wg.astype(f32)
