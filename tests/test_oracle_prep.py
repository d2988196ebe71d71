"""Pin the pre-inference pipe oracle to the reference's own code (CPU)."""
import os

import numpy as np
import pytest

from oracle import prep_oracle as po
from oracle import ref_loader


def _golden(golden_dir):
    return np.load(os.path.join(golden_dir, 'prep_ref.npz'))


def _names(g):
    return [k[3:] for k in g.files if k.startswith('in_')]


def test_norm_and_outliers_match_reference_golden(golden_dir):
    g = _golden(golden_dir)
    assert len(_names(g)) >= 3
    for name in _names(g):
        img = g['in_' + name]
        ref = g['norm_' + name]
        got = po.image_norm(img)
        assert got.dtype == ref.dtype == np.float32 and got.shape == ref.shape
        np.testing.assert_array_equal(got, ref)
        np.testing.assert_array_equal(po.image_outliers(img), g['outl2_' + name])
        np.testing.assert_array_equal(po.image_outliers(img, sigma=3, threshold=2.), g['outl3_' + name])
        assert (g['outl2_' + name] != po.promote(img)).sum() >= 1          # the hot pixels were replaced


def test_bgsubtract_matches_reference_golden(golden_dir):
    g = _golden(golden_dir)
    n = 0
    for name in _names(g):
        if 'bg_' + name not in g.files:
            continue
        ref = g['bg_' + name]
        got = po.image_bgsubtract(g['in_' + name])
        assert got.dtype == ref.dtype == np.float64 and got.shape == ref.shape
        # the reference inverts the raw normal matrix: ~1e-12 of numerical noise on values ~1e2
        np.testing.assert_allclose(got, ref, rtol=0, atol=1e-9)
        n += 1
    assert n >= 2


@pytest.mark.skipif(not ref_loader.reference_available(), reason='reference tree not present')
def test_live_reference_larger_images():
    ref = ref_loader.load_reference_pipeline()
    rng = np.random.default_rng(5)
    yy, xx = np.mgrid[0:300, 0:260]
    img = (900 + 0.3 * xx - 0.2 * yy + 1e-3 * xx * xx + rng.standard_normal((300, 260)) * 12).astype(np.float32)
    img[rng.integers(0, 300, 40), rng.integers(0, 260, 40)] += 500
    np.testing.assert_array_equal(po.image_norm(img), ref.ImageNorm()(img.copy()))
    np.testing.assert_array_equal(po.image_outliers(img, 2, 5.), ref.ImageOutliers()(img.copy()))
    np.testing.assert_allclose(po.image_bgsubtract(img), ref.ImageBGSubtract()(img.copy()), rtol=0, atol=1e-8)
