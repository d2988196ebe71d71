"""Cross-check the tr_augment oracle (TensorFlow 1.x projective resampling restated in NumPy float32)
against two independent implementations of zero-filled bilinear / nearest resampling (CPU)."""
import numpy as np
import pytest
import scipy.ndimage as ndi
import torch
import torch.nn.functional as F

from oracle import augment_oracle as ao


def _case(seed, h, w, c=2):
    rng = np.random.default_rng(seed)
    img = rng.standard_normal((h, w, c)).astype(np.float32)
    lab = rng.integers(0, 4, (h, w)).astype(np.uint8)
    wgt = rng.uniform(1, 11, (h, w)).astype(np.float32)
    return img, lab, wgt


@pytest.mark.parametrize('theta', [0.0, 0.3, np.pi / 2, 2.5, 4.0, 6.1])
def test_bilinear_matches_scipy_and_torch(theta):
    h, w = 37, 53
    img, lab, wgt = _case(1, h, w)
    t = ao.rotation_transform(theta, h, w)
    ys, xs = np.meshgrid(np.arange(h), np.arange(w), indexing='ij')
    iy, ix = ao.sample_points(t, ys, xs)
    got = ao.bilinear(img, iy, ix)
    for ch in range(img.shape[2]):
        ref = ndi.map_coordinates(img[..., ch].astype(np.float64), [iy.astype(np.float64), ix.astype(np.float64)],
                                  order=1, mode='grid-constant', cval=0.0)
        np.testing.assert_allclose(got[..., ch], ref, atol=2e-5)
    # torch: zeros padding, align_corners=True (pixel centres at integer coordinates)
    gx = torch.tensor(ix.astype(np.float64)) * 2 / (w - 1) - 1
    gy = torch.tensor(iy.astype(np.float64)) * 2 / (h - 1) - 1
    grid = torch.stack([gx, gy], -1)[None]
    tin = torch.tensor(img.astype(np.float64)).permute(2, 0, 1)[None]
    tref = F.grid_sample(tin, grid, mode='bilinear', padding_mode='zeros', align_corners=True)[0].permute(1, 2, 0)
    np.testing.assert_allclose(got, tref.numpy(), atol=2e-5)


def test_rotation_is_about_the_centre_and_identity_at_zero():
    h, w = 24, 40
    img, lab, wgt = _case(2, h, w)
    im, oh, wg = ao.tr_augment(img, lab, wgt, 0.0, 0, 0, h, w, num_outputs=4)
    np.testing.assert_array_equal(im, img)
    np.testing.assert_array_equal(wg, wgt)
    np.testing.assert_array_equal(oh.argmax(-1), lab)
    # the centre maps to itself for every angle
    for theta in (0.7, 3.0, 5.5):
        t = ao.rotation_transform(theta, h, w)
        cy, cx = np.array([(h - 1) / 2.]), np.array([(w - 1) / 2.])
        iy, ix = ao.sample_points(t, cy, cx)
        assert abs(iy[0] - cy[0]) < 1e-4 and abs(ix[0] - cx[0]) < 1e-4
    # a quarter turn of a square image is np.rot90 (nearest: exact up to the float32 sine of pi/2)
    sq = np.arange(25 * 25, dtype=np.uint8).reshape(25, 25)
    ys, xs = np.meshgrid(np.arange(25), np.arange(25), indexing='ij')
    iy, ix = ao.sample_points(ao.rotation_transform(np.pi / 2, 25, 25), ys, xs)
    r = ao.nearest(sq, iy, ix)
    assert np.array_equal(r, np.rot90(sq, 1)) or np.array_equal(r, np.rot90(sq, -1))


def test_nearest_rounds_half_away_from_zero_and_fills_zero():
    lab = np.arange(1, 13, dtype=np.uint8).reshape(3, 4)
    iy = np.array([0.5, 1.5, -0.5, -0.4999, 2.4999, 2.5], dtype=np.float32)
    ix = np.array([0.5, 2.5, 0.0, 0.0, 3.4999, 3.0], dtype=np.float32)
    out, inside = ao.nearest(lab, iy, ix, want_inside=True)
    #   (1,1)->6   (2,3)->12   (-1,0) outside   (0,0)->1   (2,3)->12   (3,3) outside
    np.testing.assert_array_equal(out, [6, 12, 0, 1, 12, 0])
    np.testing.assert_array_equal(inside, [True, True, False, True, True, False])


def test_weights_gain_one_outside_the_rotated_frame_and_labels_one_hot():
    h, w = 48, 48
    img, lab, wgt = _case(3, h, w, c=1)
    im, oh, wg = ao.tr_augment(img, lab, wgt, np.pi / 4, 4, 6, 40, 36, num_outputs=3)
    assert im.shape == (40, 36, 1) and oh.shape == (40, 36, 3) and wg.shape == (40, 36)
    assert oh.dtype == np.uint8 and set(np.unique(oh)) <= {0, 1}
    # the crop's corner lies outside the rotated frame: image 0, label class 0, weight exactly 1
    assert im[0, 0, 0] == 0 and oh[0, 0, 0] == 1 and wg[0, 0] == 1.0
    # class 3 pixels (not among the 3 outputs) have an all-zero one-hot row
    assert (oh.sum(-1) <= 1).all() and (oh.sum(-1) == 0).any()
    # inside the frame the weights stay within the input range
    centre = wg[15:25, 13:23]
    assert centre.min() >= 1.0 and centre.max() <= 11.0


def test_oracle_reproduces_the_frozen_vectors(golden_dir):
    import os
    g = np.load(os.path.join(golden_dir, 'augment_kat.npz'))
    for i in range(len(g['theta'])):
        im, oh, wg = ao.tr_augment(g['image'][i], g['label'][i], g['weights'][i], g['theta'][i], g['crop'][i, 0],
                                   g['crop'][i, 1], int(g['ch']), int(g['cw']), int(g['k']))
        assert np.array_equal(im, g['image_out'][i]) and np.array_equal(oh, g['label_out'][i])
        assert np.array_equal(wg, g['weights_out'][i])
