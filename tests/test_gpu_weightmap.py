"""Parity of the CUDA weight maps with the reference (golden vectors) and the oracle."""
import os

import numpy as np
import pytest

from oracle import weightmap_oracle as wo
from sequitr_b200 import synth

pytestmark = pytest.mark.gpu

# fp64 arithmetic identical to the reference up to exp()'s last-bit rounding (CUDA libdevice vs
# NumPy: both <= 1 ulp): tolerance in units of the weight map = 4 ulp of fp64.
RTOL64 = 1e-15 * 4


def test_w1_matches_reference_golden(sq, golden_dir):
    from sequitr_b200 import pipeline
    g = np.load(os.path.join(golden_dir, 'weightmap_ref.npz'))
    n = 0
    for key in g.files:
        if not key.startswith('w1_'):
            continue
        name, w0s, ss = key[3:].rsplit('_', 2)
        w0, sigma = float(w0s.split('-')[1]), float(ss.split('-')[1])
        got = pipeline.ImageWeightMap(w0=w0, sigma=sigma)(g['in_' + name])   # reference-facing call
        ref = g[key]
        assert got.dtype == np.float64 and got.shape == ref.shape
        np.testing.assert_allclose(got, ref, rtol=RTOL64, atol=0)
        n += 1
    assert n >= 14


def test_squared_distances_are_bit_exact(sq, golden_dir):
    from sequitr_b200 import ops
    g = np.load(os.path.join(golden_dir, 'weightmap_ref.npz'))
    for name in ('discs64', 'discs96x80', 'border40x56', 'noise48', 'allfg16', 'allbg16x24', 'single33x31'):
        m = g['in_' + name]
        _, d2 = ops.weightmap_edt_host(m, 10., 5., want_d2=True)
        np.testing.assert_array_equal(d2.astype(np.int64), wo.edt_squared(m))
    rng = np.random.default_rng(2)
    for shape, p in (((301, 517), 0.002), ((64, 1030), 0.3), ((1, 77), 0.1), ((90, 1), 0.1)):
        m = rng.random(shape) < p
        _, d2 = ops.weightmap_edt_host(m, 3., 2., want_d2=True)
        np.testing.assert_array_equal(d2.astype(np.int64), wo.edt_squared(m))


def test_w1_float32_output_and_cutoff(sq):
    from sequitr_b200 import ops
    m = synth.instance_labels(300, 420, 12, seed=4) > 0          # large empty regions -> cut-off path
    for (w0, s) in ((10., 5.), (30., 3.), (10., 0.5), (0.1, 40.)):
        ref = wo.weightmap_w1(m, w0, s)[..., 0]
        got64 = ops.weightmap_edt_host(m, w0, s, out_dtype='float64')
        np.testing.assert_allclose(got64, ref, rtol=RTOL64, atol=0)
        got32 = ops.weightmap_edt_host(m, w0, s, out_dtype='float32')
        assert got32.dtype == np.float32
        # what weightmap.py:205 saves: the fp64 map rounded to fp32 (1 ulp for the exp difference)
        np.testing.assert_allclose(got32, ref.astype(np.float32), rtol=1.2e-7, atol=0)


@pytest.mark.parametrize('seed', range(8))
def test_w1_random_shapes_and_radii(sq, seed):
    """Bounded (cut-off radius) kernels on random shapes: widths that are / are not multiples of 16 and
    32 (the two row-pass kernels), sparse to dense masks, radii from 2 to the 64-pixel limit and beyond
    (general kernels), several frames per call."""
    from sequitr_b200 import ops
    rng = np.random.default_rng(50 + seed)
    h = int(rng.integers(1, 220))
    w = int(rng.choice([16, 48, 80, 112, 176, 272])) if seed % 2 == 0 else int(rng.integers(1, 300))
    n = int(rng.integers(1, 4))
    p = float(rng.choice([0.0005, 0.005, 0.05, 0.4]))
    m = rng.random((n, h, w)) < p
    if seed == 3:
        m[1 % n] = False                                         # a seedless frame among seeded ones
    for (w0, s) in ((10., 5.), (10., float(rng.choice([0.4, 2., 9., 14.]))), (1e-3, 3.)):
        ref = np.stack([wo.weightmap_w1(f, w0, s)[..., 0] for f in m])
        got64 = ops.weightmap_edt_host(m, w0, s, out_dtype='float64')
        np.testing.assert_allclose(got64, ref, rtol=RTOL64, atol=0)
        got32 = ops.weightmap_edt_host(m, w0, s, out_dtype='float32')
        np.testing.assert_allclose(got32, ref.astype(np.float32), rtol=1.2e-7, atol=0)


def test_w1_full_size_2048(sq):
    from sequitr_b200 import ops
    lab = synth.instance_labels(2048, 2048, 600, seed=1)
    m = lab > 0
    got, d2 = ops.weightmap_edt_host(m, 10., 5., out_dtype='float64', want_d2=True)
    np.testing.assert_array_equal(d2.astype(np.int64), wo.edt_squared(m))
    np.testing.assert_allclose(got, wo.weightmap_w1(m, 10., 5.)[..., 0], rtol=RTOL64, atol=0)
    # size-independent property: the transform commutes with flips and transposition
    f = ops.weightmap_edt_host(m[::-1, ::-1].copy(), 10., 5., out_dtype='float64')
    np.testing.assert_array_equal(f[::-1, ::-1], got)
    t = ops.weightmap_edt_host(m.T.copy(), 10., 5., out_dtype='float64')
    np.testing.assert_array_equal(t.T, got)


@pytest.mark.parametrize('shape,n,seed', [((96, 96), 10, 4), ((150, 260), 30, 9), ((64, 300), 12, 2)])
def test_w3_matches_oracle(sq, shape, n, seed):
    from sequitr_b200 import ops
    lab = synth.instance_labels(shape[0], shape[1], n, seed=seed, rmin=4, rmax=9)
    lab[5:8, 5:60] = lab.max() + 1                     # thin sliver: second-nearest behind the nearest
    lab[9:11, 5:60] = lab.max() + 1
    for (w0, s, wc) in ((10., 5., None), (30., 3., (0.5, 3.0)), (10., 25., None)):
        ref = wo.weightmap_w3(lab, w0, s, wc)
        got = ops.weightmap_unet_host(lab, w0, s, wc, out_dtype='float64')
        np.testing.assert_allclose(got, ref, rtol=RTOL64, atol=0)
        got32 = ops.weightmap_unet_host(lab, w0, s, wc, out_dtype='float32')
        np.testing.assert_allclose(got32, ref.astype(np.float32), rtol=1.2e-7, atol=0)


@pytest.mark.parametrize('seed', range(6))
def test_w3_random_label_images(sq, seed):
    """Random instance-label images: noise labels (touching instances, single pixels, many labels per
    row), discs, widths around the 32-pixel tile and word sizes, several radii (bounded and general)."""
    from sequitr_b200 import ops
    rng = np.random.default_rng(70 + seed)
    h, w = int(rng.integers(1, 150)), int(rng.integers(1, 200))
    if seed % 2 == 0:
        lab = np.where(rng.random((h, w)) < float(rng.choice([0.01, 0.1, 0.5])),
                       rng.integers(1, 6, size=(h, w)), 0).astype(np.int32)
    else:
        lab = synth.instance_labels(max(h, 24), max(w, 24), 8, seed=seed, rmin=2, rmax=6).astype(np.int32)
    for (w0, s, wc) in ((10., 5., None), (10., float(rng.choice([0.5, 2., 12.])), (0.25, 1.5))):
        ref = wo.weightmap_w3(lab, w0, s, wc)
        got = ops.weightmap_unet_host(lab, w0, s, wc, out_dtype='float64')
        np.testing.assert_allclose(got, ref, rtol=RTOL64, atol=0)
        got32 = ops.weightmap_unet_host(lab, w0, s, wc, out_dtype='float32')
        np.testing.assert_allclose(got32, ref.astype(np.float32), rtol=1.2e-7, atol=0)


def test_w3_wide_labels_take_the_scan_kernel(sq):
    """The DPX column pass packs (squared distance, label) into one 32-bit key, labels < 2^18; frames
    holding wider labels are flagged by the row pass and go through the scan kernel.  Same results, also
    when one call mixes both kinds of frame."""
    import torch
    from sequitr_b200 import ops
    lab = synth.instance_labels(120, 200, 20, seed=11, rmin=3, rmax=8).astype(np.int32)
    wide = np.where(lab > 0, lab + 300000, 0).astype(np.int32)          # same geometry, labels >= 2^18
    huge = np.where(lab > 0, lab * 7 + (1 << 30), 0).astype(np.int32)
    ref = wo.weightmap_w3(lab, 10., 5.)
    for variant in (wide, huge):
        np.testing.assert_allclose(ops.weightmap_unet_host(variant, 10., 5., out_dtype='float64'), ref,
                                   rtol=RTOL64, atol=0)
    stack = np.stack([lab, wide, lab[::-1].copy(), huge])
    got = ops.weightmap_unet(torch.from_numpy(stack).cuda(), 10., 5., None, 'float64').cpu().numpy()
    for i, r in enumerate((ref, ref, ref[::-1], ref)):
        np.testing.assert_allclose(got[i], r, rtol=RTOL64, atol=0)
    # labels just below / at the limit
    edge = np.where(lab > 0, lab + (1 << 18) - 1 - lab.max(), 0).astype(np.int32)      # max label = 2^18 - 1
    np.testing.assert_allclose(ops.weightmap_unet_host(edge, 10., 5., out_dtype='float64'), ref, rtol=RTOL64, atol=0)
    np.testing.assert_allclose(ops.weightmap_unet_host(edge + (edge > 0), 10., 5., out_dtype='float64'), ref,
                               rtol=RTOL64, atol=0)


def test_w3_edge_cases_and_properties(sq):
    from sequitr_b200 import ops, pipeline
    empty = np.zeros((40, 50), np.int32)
    np.testing.assert_array_equal(ops.weightmap_unet_host(empty), np.ones((40, 50)))
    one = empty.copy()
    one[10:20, 10:20] = 3
    w = ops.weightmap_unet_host(one)
    assert (w[one == 0] == 1.0).all() and (w[one > 0] == 2.0).all()
    lab = synth.instance_labels(2048, 2048, 600, seed=3)
    w3 = ops.weightmap_unet_host(lab, 10., 5.)
    w1 = ops.weightmap_edt_host(lab > 0, 10., 5.)
    assert (w3 <= w1 + 1e-12).all()                    # d1 + d2 >= d1
    assert w3.max() > 1.5                              # narrow gaps do get boosted
    f = ops.weightmap_unet_host(lab[::-1].copy(), 10., 5.)
    np.testing.assert_array_equal(f[::-1], w3)
    # the pipe splits a bool mask into instances with the GPU component labeller
    sub = lab[:256, :256]
    from scipy.ndimage import label
    via_pipe = pipeline.ImageWeightMapUNet(10., 5.)(sub > 0)
    np.testing.assert_allclose(via_pipe[..., 0], wo.weightmap_w3(label(sub > 0)[0], 10., 5.),
                               rtol=RTOL64, atol=0)


def test_w3_full_size_2048_against_the_cropped_oracle(sq):
    """BASELINE configs[1] size: 2048^2 instance labels, ~600 instances, w0 = 10, sigma = 5, against the exact
    per-instance oracle (cropped transforms; tests/test_oracle_weightmap.py pins it to the brute-force one)."""
    from sequitr_b200 import ops
    lab = synth.instance_labels(2048, 2048, 600, seed=3)
    ref = wo.weightmap_w3(lab, 10., 5., margin=48)             # 10 exp(-48^2 / 50) ~ 1e-19: far below one ulp of 1.0
    got = ops.weightmap_unet_host(lab, 10., 5., out_dtype='float64')
    np.testing.assert_allclose(got, ref, rtol=RTOL64, atol=0)
    assert (ref > 1.5).sum() > 1000                            # the gap term is exercised
    got32 = ops.weightmap_unet_host(lab, 10., 5., out_dtype='float32')
    np.testing.assert_allclose(got32, ref.astype(np.float32), rtol=1.2e-7, atol=0)      # the fp64 map rounded once


def test_device_api(sq):
    import torch
    from sequitr_b200 import ops
    lab = np.stack([synth.instance_labels(128, 192, 14, seed=s, rmin=4, rmax=9) for s in (1, 2, 3)])
    m = torch.from_numpy((lab > 0).astype(np.uint8)).cuda()
    out = ops.weightmap_edt(m, 10., 5., out_dtype='float32')
    out3 = ops.weightmap_unet(torch.from_numpy(lab.astype(np.int32)).cuda(), 10., 5.)
    torch.cuda.synchronize()
    for i in range(3):
        np.testing.assert_allclose(out[i].cpu().numpy(),
                                   wo.weightmap_w1(lab[i] > 0, 10., 5.)[..., 0].astype(np.float32),
                                   rtol=1.2e-7)
        np.testing.assert_allclose(out3[i].cpu().numpy(),
                                   wo.weightmap_w3(lab[i], 10., 5.).astype(np.float32), rtol=1.2e-7)


def test_create_weightmaps_gpu_methods(sq, tmp_path):
    import cv2
    from sequitr_b200 import weightmap
    lab_dir = tmp_path / 'setA' / 'label'
    lab_dir.mkdir(parents=True)
    mask = (synth.instance_labels(96, 128, 9, seed=1, rmin=4, rmax=9) > 0)
    cv2.imwrite(str(lab_dir / 'pos1_0001.tif'), mask.astype(np.uint8) * 255)
    out = weightmap.create_weightmaps(str(tmp_path), ['setA'], w0=10., sigma=5., method='edt')
    w = cv2.imread(out[0], cv2.IMREAD_UNCHANGED)
    np.testing.assert_allclose(w, wo.weightmap_w1(mask, 10., 5.)[..., 0].astype(np.float32), rtol=1.2e-7)
    out = weightmap.create_weightmaps(str(tmp_path), ['setA'], w0=10., sigma=5., method='unet')
    from scipy.ndimage import label
    w = cv2.imread(out[0], cv2.IMREAD_UNCHANGED)
    np.testing.assert_allclose(w, wo.weightmap_w3(label(mask)[0], 10., 5.).astype(np.float32), rtol=1.2e-7)
    # the default stays the reference's own pipe (weightmap.py:181: ImageWeightMap2, host-side Delaunay)
    out = weightmap.create_weightmaps(str(tmp_path), ['setA'], w0=10., sigma=5.)
    w = cv2.imread(out[0], cv2.IMREAD_UNCHANGED)
    np.testing.assert_allclose(w, wo.weightmap_w2(mask, 10., 5.)[..., 0].astype(np.float32), rtol=1e-6)
