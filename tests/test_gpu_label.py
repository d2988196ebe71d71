"""Parity of the CUDA label-and-localise path with the SciPy oracle (bit-exact)."""
import os

import numpy as np
import pytest

from oracle import centroid_oracle as co
from sequitr_b200 import synth

pytestmark = pytest.mark.gpu


def _check(ops, stack, swapped_for_gpu=None, max_rows=4096):
    want = co.centroid_tables(stack)
    gpu_in = stack if swapped_for_gpu is None else swapped_for_gpu
    got = ops.label_centroids_host(gpu_in, max_rows=max_rows)
    assert len(got) == len(want)
    for i, (a, b) in enumerate(zip(got, want)):
        assert a.dtype == np.float32 and a.shape == b.shape, (i, a.shape, b.shape)
        np.testing.assert_array_equal(a, b)


def test_golden_fixtures(sq, golden_dir):
    from sequitr_b200 import ops, utils
    g = np.load(os.path.join(golden_dir, 'centroid_scipy.npz'))
    for name in ('discs2c', 'noise', 'empty', 'ushape', 'vol'):
        st = g['in_' + name]
        tabs = utils.CentroidWriter.centroids(st)        # the reference-facing entry point
        counts = np.array([len(t) for t in tabs], np.int32)
        np.testing.assert_array_equal(counts, g['counts_' + name])
        flat = np.concatenate(tabs, 0) if counts.sum() else np.zeros((0, 5), np.float32)
        np.testing.assert_array_equal(flat, g['table_' + name])


@pytest.mark.parametrize('shape,density,classes', [((3, 37, 53), 0.5, 1), ((2, 64, 96), 0.6, 3),
                                                   ((1, 130, 70), 0.35, 5), ((2, 33, 31), 0.9, 2)])
def test_random_masks(sq, shape, density, classes):
    from sequitr_b200 import ops
    rng = np.random.default_rng(sum(shape))
    m = (rng.random(shape) < density).astype(np.uint8) * rng.integers(1, classes + 1, shape).astype(np.uint8)
    _check(ops, m, max_rows=8192)


def test_label_matrix_matches_scipy_numbering(sq):
    from sequitr_b200 import ops
    m = synth.class_mask(200, 264, 40, n_classes=3, seed=8, rmin=4, rmax=10)[None]
    m[0, 50, :] = 2                                   # a long thin bar crossing everything
    tabs, labels = ops.label_centroids_host(m, want_labels=True)
    np.testing.assert_array_equal(labels[0], co.label_matrix(m[0]))
    np.testing.assert_array_equal(tabs[0], co.centroid_tables(m)[0])


def test_edge_cases(sq):
    from sequitr_b200 import ops
    _check(ops, np.zeros((2, 16, 16), np.uint8))                       # empty frames
    _check(ops, np.full((1, 40, 72), 7, np.uint8))                     # one frame-filling object
    cb = (np.indices((1, 32, 32)).sum(0) % 2).astype(np.uint8)         # checkerboard: 512 singletons
    _check(ops, cb)
    tabs = ops.label_centroids_host(cb, max_rows=16)                   # overflow -> retried larger
    assert len(tabs[0]) == 512
    sp = np.zeros((1, 64, 64), np.uint8)                               # spiral / nested U shapes
    for k in range(2, 30, 4):
        sp[0, k, k:64 - k] = 1
        sp[0, k:64 - k, 64 - k - 1] = 1
        sp[0, 64 - k - 1, k:64 - k] = 1
        sp[0, k + 4:64 - k, k] = 1
    _check(ops, sp)
    one = np.zeros((1, 1, 1), np.uint8)
    one[0, 0, 0] = 255
    _check(ops, one)
    with pytest.raises(ValueError):
        ops.label_centroids_host(np.zeros((4, 4), np.uint8))


def test_volumetric_6_connectivity(sq):
    from sequitr_b200 import ops, utils
    rng = np.random.default_rng(5)
    vol = (rng.random((2, 9, 20, 24)) > 0.6).astype(np.uint8) * rng.integers(1, 3, (2, 9, 20, 24)).astype(np.uint8)
    want = co.centroid_tables(vol)                       # oracle swaps (N,Z,X,Y) -> (N,Y,X,Z)
    got = utils.CentroidWriter.centroids(vol)
    for a, b in zip(got, want):
        np.testing.assert_array_equal(a, b)


def test_full_size_frame_2048(sq):
    from sequitr_b200 import ops
    m = np.stack([synth.class_mask(2048, 2048, 600, n_classes=3, seed=s) for s in (1, 2)])
    _check(ops, m)
    # size-independent property: translating the frame translates every centroid
    sh = np.zeros_like(m)
    sh[:, 3:, 5:] = m[:, :-3, :-5]
    a = ops.label_centroids_host(m[:, :-3, :-5].copy())
    b = ops.label_centroids_host(sh)
    for x, y in zip(a, b):
        # (rows are float32: the shifted centroid may round differently by one ulp)
        np.testing.assert_allclose(x[:, 1].astype(np.float64) + 3, y[:, 1], rtol=1.2e-7)
        np.testing.assert_allclose(x[:, 2].astype(np.float64) + 5, y[:, 2], rtol=1.2e-7)
        np.testing.assert_array_equal(x[:, 4], y[:, 4])


@pytest.mark.parametrize('shape,n', [((4096, 4096), 1), ((1500, 3000), 3), ((33, 5000), 4), ((700, 40), 9)])
def test_look_back_scan_over_many_and_few_blocks(sq, shape, n):
    """The single-pass front end places runs in raster order with a decoupled look-back scan over blocks of 128 row
    segments: 256 blocks per frame at 4096^2 (eight look-back rounds), ragged last blocks, rows shorter than one
    segment, several frames with their own tickets -- all bit-identical to SciPy."""
    from sequitr_b200 import ops
    m = np.stack([synth.class_mask(shape[0], shape[1], max(4, shape[0] * shape[1] // 7000), n_classes=2, seed=10 + s,
                                   rmin=3, rmax=12) for s in range(n)])
    _check(ops, m)
    # worst case for the run lists: a checkerboard of single-pixel runs in the top-left corner of every frame
    cb = np.zeros_like(m)
    cb[:, :32:2, :64:2] = 1
    cb[:, 1:32:2, 1:64:2] = 2
    _check(ops, cb)


def test_device_api_equals_host_api(sq):
    import torch
    from sequitr_b200 import ops
    m = synth.class_mask(256, 320, 50, n_classes=2, seed=3, rmin=4, rmax=10)[None].repeat(3, 0)
    m[1] = np.roll(m[1], 17, axis=1)
    host = ops.label_centroids_host(m, frame0=10)
    table, counts = ops.label_centroids(torch.from_numpy(m).cuda(), max_rows=256, frame0=10)
    torch.cuda.synchronize()
    for i in range(3):
        np.testing.assert_array_equal(table[i, :int(counts[i])].cpu().numpy(), host[i])
        assert host[i][0, 0] == 10 + i


def test_centroid_writer_file(sq, tmp_path):
    """CentroidWriter.write (utils.py:505-578) end to end: GPU label-and-localise -> a real HDF5 file with the
    reference's ``frames/frame_<i>/coords`` layout (float32 (n,5) rows; shape (0,) for a frame without objects),
    read back by the package's byte-level reader (h5py is not in this image) and compared with the oracle."""
    from sequitr_b200 import hdf5min, utils
    m = np.stack([synth.class_mask(96, 96, 8, n_classes=2, seed=1, rmin=4, rmax=8),
                  np.zeros((96, 96), np.uint8),                                  # nothing to find
                  synth.class_mask(96, 96, 5, n_classes=3, seed=2, rmin=3, rmax=6)])
    w = utils.CentroidWriter(str(tmp_path / 'cells.hdf5'))
    w.write(m)
    w.close()
    want = co.centroid_tables(m)
    raw = open(str(tmp_path / 'cells.hdf5'), 'rb').read()
    assert raw[:8] == b'\x89HDF\r\n\x1a\n'
    with hdf5min.File(str(tmp_path / 'cells.hdf5'), 'r') as r:
        assert r['frames'].keys() == ['frame_0', 'frame_1', 'frame_2']
        for i in range(3):
            got = r['frames']['frame_%d' % i]['coords']
            assert got.dtype == np.float32
            if len(want[i]):
                np.testing.assert_array_equal(got[...], want[i])
            else:
                assert got.shape == (0,)
    assert len(want[0]) > 0 and len(want[1]) == 0 and len(want[2]) > 0
