"""Pin the weight-map oracle to the reference's own code (CPU)."""
import os

import numpy as np
import pytest

from oracle import weightmap_oracle as wo
from oracle import ref_loader
from sequitr_b200 import synth


def _golden(golden_dir):
    return np.load(os.path.join(golden_dir, 'weightmap_ref.npz'))


def test_w1_matches_reference_golden(golden_dir):
    g = _golden(golden_dir)
    n = 0
    for key in g.files:
        if not key.startswith('w1_'):
            continue
        name, w0s, ss = key[3:].rsplit('_', 2)
        w0, sigma = float(w0s.split('-')[1]), float(ss.split('-')[1])
        got = wo.weightmap_w1(g['in_' + name], w0, sigma)
        ref = g[key]
        assert got.dtype == ref.dtype == np.float64 and got.shape == ref.shape
        np.testing.assert_array_equal(got, ref)
        n += 1
    assert n >= 14


def test_w2_matches_reference_golden(golden_dir):
    g = _golden(golden_dir)
    for name in ('discs64', 'discs96x80'):
        got = wo.weightmap_w2(g['in_' + name], 10., 5.)
        ref = g['w2_%s_w0-10_s-5' % name]
        assert got.shape == ref.shape
        np.testing.assert_allclose(got, ref, rtol=0, atol=1e-12)


@pytest.mark.skipif(not ref_loader.reference_available(), reason="reference tree not present")
def test_w1_w2_match_live_reference():
    ref = ref_loader.load_reference_pipeline()
    mask = synth.instance_labels(128, 160, 20, seed=21, rmin=4, rmax=10) > 0
    for (w0, s) in ((10., 5.), (30., 3.)):
        np.testing.assert_array_equal(wo.weightmap_w1(mask, w0, s),
                                      ref.ImageWeightMap(w0=w0, sigma=s)(mask.copy()))
    np.testing.assert_allclose(wo.weightmap_w2(mask, 10., 5.),
                               ref.ImageWeightMap2(w0=10., sigma=5.)(mask.copy()), atol=1e-12)


def test_edt_squared_is_exact_and_handles_empty(golden_dir):
    g = _golden(golden_dir)
    from scipy.ndimage import distance_transform_edt
    for name in ('discs64', 'border40x56', 'noise48', 'single33x31', 'allbg16x24', 'allfg16'):
        m = g['in_' + name]
        d2 = wo.edt_squared(m)
        d = distance_transform_edt(1. - m.astype('float32'))
        np.testing.assert_array_equal(np.sqrt(d2.astype(np.float64)), d)


def test_w3_reduces_to_w1_distance_and_is_symmetric_in_gap():
    lab = synth.instance_labels(96, 96, 10, seed=4, rmin=4, rmax=9)
    d1, d2 = wo.two_nearest_instances_d2(lab)
    np.testing.assert_array_equal(d1, wo.edt_squared(lab > 0))   # d1 == W1's distance
    assert (d2[lab == 0] >= d1[lab == 0]).all()
    w = wo.weightmap_w3(lab, 10., 5.)
    assert w.shape == lab.shape and (w[lab > 0] == 2.0).all()
    assert (w[lab == 0] >= 1.0).all() and w.max() <= 11.0
    # single instance: no second instance -> gap term vanishes
    one = np.zeros((32, 32), np.int32)
    one[10:14, 10:14] = 7
    w1 = wo.weightmap_w3(one, 10., 5.)
    assert (w1[one == 0] == 1.0).all()


def test_image_labels_and_names():
    raw = np.zeros((3, 8, 8), np.uint8)
    raw[0, :4] = 5
    raw[2, 2:6] = 1
    lab, n = wo.image_labels(raw)
    assert n == 4 and lab.dtype == np.uint8 and set(np.unique(lab)) == {0, 1, 3}
    assert lab[3, 0] == 3 and lab[0, 0] == 1
    lab2, n2 = wo.image_labels(np.arange(16).reshape(4, 4))
    assert n2 == 2 and lab2.sum() == 15
    with pytest.raises(ValueError):
        wo.image_labels(np.zeros((5, 4, 4)))
    assert wo.weights_folder_name(10., 3.) == 'weights_w0-10.00_sigma-3.00'
    assert wo.weights_folder_name(10., 3., False) == 'weights'
    assert wo.weights_file_name('pos1_GFP_0001.tif') == 'pos1_GFP_weights.tif'


def test_cropped_w3_oracle_equals_the_brute_force_one():
    """The per-instance cropped oracle used at BASELINE's 2048^2 size against the brute-force oracle: identical
    squared distances up to the margin and weight maps equal to the last bit of the exp term that can matter."""
    from sequitr_b200 import synth
    lab = synth.instance_labels(200, 260, 30, seed=4, rmin=4, rmax=10)
    d1, d2 = wo.two_nearest_instances_d2(lab)
    c1, c2 = wo.two_nearest_instances_d2_cropped(lab, 48)
    near1, near2 = d1 <= 48 * 48, d2 <= 48 * 48
    np.testing.assert_array_equal(c1[near1], d1[near1])
    np.testing.assert_array_equal(c2[near2], d2[near2])
    assert (c1[~near1] > 48 * 48).all() and (c2[~near2] > 48 * 48).all()
    np.testing.assert_allclose(wo.weightmap_w3(lab, 10., 5., margin=48), wo.weightmap_w3(lab, 10., 5.), rtol=1e-15, atol=0)
