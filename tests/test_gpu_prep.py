"""Pre-inference pipes on the GPU against the reference's outputs.

ImageOutliers is comparison/selection only -> bit-exact.  ImageNorm: the GPU takes the moments in
fp64 and rounds them to float32, NumPy sums in float32 pairwise; the two float32 means / stds can
differ in the last ulp, which moves a normalised value by <= ~1e-6 relative (tolerance 4e-6 + 4e-6 |x|).
ImageBGSubtract: fp64 fit, |difference| <= 1e-8 on backgrounds ~1e2..1e3.
"""
import os

import numpy as np
import pytest

from oracle import prep_oracle as po

pytestmark = pytest.mark.gpu


def _golden(golden_dir):
    return np.load(os.path.join(golden_dir, 'prep_ref.npz'))


def _close_norm(got, ref):
    assert got.dtype == ref.dtype == np.float32 and got.shape == ref.shape
    np.testing.assert_allclose(got, ref, rtol=4e-6, atol=4e-6)


def test_pipes_match_reference_golden(sq, golden_dir):
    from sequitr_b200 import pipeline
    g = _golden(golden_dir)
    for name in [k[3:] for k in g.files if k.startswith('in_')]:
        img = g['in_' + name]
        _close_norm(pipeline.ImageNorm()(img.copy()), g['norm_' + name])
        got = pipeline.ImageOutliers()(img.copy())
        assert got.dtype == np.float32
        np.testing.assert_array_equal(got, g['outl2_' + name])
        np.testing.assert_array_equal(pipeline.ImageOutliers(sigma=3, threshold=2.)(img.copy()), g['outl3_' + name])
        if 'bg_' + name in g.files:
            bg = pipeline.ImageBGSubtract()(img.copy())
            assert bg.dtype == np.float64 and bg.shape == g['bg_' + name].shape
            np.testing.assert_allclose(bg, g['bg_' + name], rtol=0, atol=1e-8)


def test_in_place_semantics_and_pipeline(sq):
    """ImageNorm / ImageOutliers update the caller's array like the reference; ImageBGSubtract does not."""
    from sequitr_b200 import pipeline
    rng = np.random.default_rng(3)
    img = (rng.standard_normal((48, 40, 1)) * 4 + 20).astype(np.float32)
    img[5, 5, 0] = 500.
    keep = img.copy()
    out = pipeline.ImageOutliers()(img)
    assert out is img and img[5, 5, 0] != 500.
    np.testing.assert_array_equal(img, po.image_outliers(keep))
    out = pipeline.ImageNorm()(img)
    assert out is img and abs(float(img.mean())) < 1e-5 and abs(float(img.std()) - 1) < 1e-5
    before = img.copy()
    bg = pipeline.ImageBGSubtract()(img)
    assert bg is not img
    np.testing.assert_array_equal(img, before)
    with pytest.raises(ValueError):
        pipeline.ImageBGSubtract()(np.zeros((8, 8, 3), np.float32))


@pytest.mark.parametrize('size', [1, 2, 3, 4, 5])
def test_outliers_every_median_size_and_tiny_images(sq, size):
    from sequitr_b200 import ops
    rng = np.random.default_rng(size)
    for shape in ((1, 1, 1), (2, 3, 1), (5, 4, 2), (37, 53, 3)):
        img = (rng.standard_normal(shape) * 3).astype(np.float32)
        got = ops.image_pipe_host('outliers', img, size=size, threshold=1.5)
        np.testing.assert_array_equal(got, po.image_outliers(img, size, 1.5))


def test_full_size_stack_on_device(sq):
    """BASELINE's frame size, batch of 4, device-resident tensors."""
    import torch
    from sequitr_b200 import ops, synth
    x = synth.frames(4, 2048, 2048, 1, seed=99) * 37.0 + 500.0
    yy, xx = np.mgrid[0:2048, 0:2048]
    x[..., 0] += (0.02 * xx + 0.01 * yy + 3e-6 * xx * yy).astype(np.float32)
    x = x.astype(np.float32)
    xd = torch.from_numpy(x).cuda()
    nrm = ops.image_norm(xd).cpu().numpy()
    for i in range(4):
        _close_norm(nrm[i], po.image_norm(x[i]))
    out = ops.image_outliers(xd, 2, 5.).cpu().numpy()
    np.testing.assert_array_equal(out[1], po.image_outliers(x[1], 2, 5.))
    bg = ops.image_bgsubtract(xd, out_dtype='float64').cpu().numpy()
    np.testing.assert_allclose(bg[2], po.image_bgsubtract(x[2]), rtol=0, atol=1e-7)
    bg32 = ops.image_bgsubtract(xd, out_dtype='float32').cpu().numpy()
    np.testing.assert_allclose(bg32[2], po.image_bgsubtract(x[2]).astype(np.float32), rtol=0, atol=1e-4)
    # in-place normalisation (in == out) is allowed
    ops.image_norm(xd, out=xd)
    np.testing.assert_array_equal(xd.cpu().numpy(), nrm)


def test_raw_camera_frames_through_the_hot_path(sq, tmp_path):
    """uint16 / uint8 frames (dataio.OctopusData.frames_raw) cross PCIe as stored and are widened -- and
    optionally normalised -- on the device: identical tables and masks to the float32 route."""
    import torch
    from sequitr_b200 import ops, synth
    from sequitr_b200.dataio import OctopusData, write_octopus_stream
    from sequitr_b200.networks import UNet2D
    filters = (16, 32, 64)
    net = UNet2D({'filters': filters, 'shape': (96, 128), 'bridge': 'concat', 'compute': 'bf16'})
    net.load_weights(synth.blob_detector_weights(filters, 1, 2, seed=1))
    x = synth.frames(10, 96, 128, 1, seed=3, n_objects=5)[..., 0]
    raw = np.clip(x * 400.0 + 3000.0, 0, 65535).astype(np.uint16)
    stem = str(tmp_path / 'BF_pos0_')
    write_octopus_stream(stem, raw, frames_per_file=4)
    batch = OctopusData(stem).frames_raw(0, 10)
    np.testing.assert_array_equal(batch, raw)
    # no normalisation: same as handing the float32 copy of the integers
    t_raw, m_raw = net.segment_and_localise(batch, return_mask=True)
    t_f32, m_f32 = net.segment_and_localise(batch.astype(np.float32), return_mask=True)
    np.testing.assert_array_equal(m_raw, m_f32)
    for a, b in zip(t_raw, t_f32):
        np.testing.assert_array_equal(a, b)
    # with ImageNorm on the device == image_norm then the float route (same kernels, same bits)
    t_n, m_n = net.segment_and_localise(batch, return_mask=True, normalise=True)
    xn = ops.image_norm(torch.from_numpy(batch.astype(np.float32)[..., None]).cuda()).cpu().numpy()
    t_ref, m_ref = net.segment_and_localise(xn, return_mask=True)
    np.testing.assert_array_equal(m_n, m_ref)
    for a, b in zip(t_n, t_ref):
        np.testing.assert_array_equal(a, b)
    assert sum(len(t) for t in t_n) >= 10                  # the discs are found on normalised frames
    # 8-bit stream
    raw8 = (raw >> 6).astype(np.uint8)
    t8, m8 = net.segment_and_localise(raw8, return_mask=True, normalise=True)
    x8 = ops.image_norm(torch.from_numpy(raw8.astype(np.float32)[..., None]).cuda()).cpu().numpy()
    np.testing.assert_array_equal(m8, net.segment_and_localise(x8, return_mask=True)[1])


@pytest.mark.parametrize('shape', [(70, 120), (256, 264)])
def test_image_norm_feeds_the_fused_first_pair_in_16_bits(sq, monkeypatch, shape):
    """uint16 frames + ImageNorm on the host-call route: by default the frames are widened, normalised
    ((x - mean) / std in float32, the arithmetic of sq_image_norm) and rounded to bf16 ONCE (2 B/px) and
    conv_qf_kernel's builders load that as is; SQ_QNORM=2: only the moments are computed up front and the builders
    normalise the raw values themselves; SQ_QNORM=0: the float32 stage of round 1.  Same bits on all three."""
    from sequitr_b200 import synth
    from sequitr_b200.networks import UNet2D
    filters = (16, 32)
    net = UNet2D({'filters': filters, 'shape': shape, 'bridge': 'concat', 'compute': 'bf16'})
    net.load_weights(synth.blob_detector_weights(filters, 1, 2, seed=1))
    x = synth.frames(9, shape[0], shape[1], 1, seed=4, n_objects=6)[..., 0]
    raw = np.clip(x * 400.0 + 3000.0, 0, 65535).astype(np.uint16)
    t_bf16, m_bf16 = net.segment_and_localise(raw, return_mask=True, normalise=True)
    for mode in ('0', '2'):
        monkeypatch.setenv('SQ_QNORM', mode)
        t_other, m_other = net.segment_and_localise(raw, return_mask=True, normalise=True)
        np.testing.assert_array_equal(m_bf16, m_other)
        for a, b in zip(t_bf16, t_other):
            np.testing.assert_array_equal(a, b)
    assert sum(len(t) for t in t_bf16) >= 9


def test_overlapped_stack_calls_give_the_same_tables(sq):
    """shard.segment_stack(overlap=True): two host threads, the second through net.twin() (same weights, its own
    library handle, streams and device arena), alternate the calls -- tables identical to the sequential walk."""
    from sequitr_b200 import shard, synth
    from sequitr_b200.networks import UNet2D
    filters = (16, 32, 64)
    net = UNet2D({'filters': filters, 'shape': (128, 160), 'bridge': 'concat', 'compute': 'bf16'})
    net.load_weights(synth.blob_detector_weights(filters, 1, 2, seed=1))
    x = synth.frames(37, 128, 160, 1, seed=6, n_objects=5)[..., 0]
    raw = np.clip(x * 400.0 + 3000.0, 0, 65535).astype(np.uint16)
    seq = shard.segment_stack(net, raw, frame0=100, frames_per_call=8, max_rows=256)
    for _ in range(2):
        over = shard.segment_stack(net, raw, frame0=100, frames_per_call=8, max_rows=256, overlap=True)
        assert len(over) == 37 and shard.tables_digest(over) == shard.tables_digest(seq)
    assert sum(len(t) for t in seq) >= 37 and seq[5][0, 0] == 105.0
