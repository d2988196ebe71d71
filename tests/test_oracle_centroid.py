"""Pin the label-and-localise oracle (SciPy restatement of utils.CentroidWriter.write)."""
import os

import numpy as np

from oracle import centroid_oracle as co


def test_golden_tables(golden_dir):
    g = np.load(os.path.join(golden_dir, 'centroid_scipy.npz'))
    for name in ('discs2c', 'noise', 'empty', 'ushape', 'vol'):
        tabs = co.centroid_tables(g['in_' + name])
        counts = np.array([len(t) for t in tabs], np.int32)
        np.testing.assert_array_equal(counts, g['counts_' + name])
        flat = np.concatenate(tabs, 0) if counts.sum() else np.zeros((0, 5), np.float32)
        assert flat.dtype == np.float32
        np.testing.assert_array_equal(flat, g['table_' + name])


def test_scipy_facts_the_gpu_kernel_relies_on():
    # 4-connectivity, raster-order numbering even when provisional labels merge
    u = np.zeros((1, 6, 7), np.uint8)
    u[0, 1:5, 1] = 1
    u[0, 1:5, 5] = 1
    u[0, 4, 1:6] = 1
    u[0, 0, 3] = 1
    t = co.centroid_tables(u)[0]
    assert len(t) == 2
    assert tuple(t[0, 1:3]) == (0.0, 3.0)          # first in raster order = the lone pixel
    d = np.zeros((1, 4, 4), np.uint8)
    d[0, 0, 0] = d[0, 1, 1] = 1                    # diagonal contact does not connect
    assert len(co.centroid_tables(d)[0]) == 2
    # classes ascending, rows [frame, axis0, axis1, 0, class]
    m = np.zeros((2, 5, 5), np.uint8)
    m[1, 0, 0] = 3
    m[1, 4, 1:4] = 2
    t = co.centroid_tables(m)
    assert len(t[0]) == 0 and t[0].shape == (0, 5)
    np.testing.assert_array_equal(t[1], np.array([[1, 4, 2, 0, 2], [1, 0, 0, 0, 3]], np.float32))


def test_volumetric_axes_swap():
    v = np.zeros((1, 3, 4, 5), np.uint8)           # (N,Z,X,Y)
    v[0, 2, 1, 4] = 1
    t = co.centroid_tables(v)[0]
    # swapaxes(1,-1) -> (N,Y,X,Z): columns are (axis0=Y index, axis1=X index, axis2=Z index)
    np.testing.assert_array_equal(t, np.array([[0, 4, 1, 2, 1]], np.float32))
