"""Host-side logic and the C-ABI surface, runnable without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

import sequitr_b200
from sequitr_b200 import _lib, ops, pipeline, utils, weightmap, synth
from sequitr_b200.networks import unet as U


def test_library_loads_and_exports_every_declared_symbol():
    lib = _lib.load()
    declared = _lib.declared_symbols()
    assert len(declared) >= 22
    for name in declared:
        assert hasattr(lib, name), 'missing export %s' % name
        assert name in _lib._SIGNATURES, 'no ctypes signature for %s' % name
    assert b'sm_100a' in lib.sq_version()


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    with pytest.raises(_lib.SequitrError):
        ops.weightmap_edt_host(np.zeros((8, 8), np.uint8))
    with pytest.raises(_lib.SequitrError):
        pipeline.ImageWeightMap()(np.zeros((8, 8), bool))
    with pytest.raises(_lib.SequitrError):
        utils.CentroidWriter.centroids(np.zeros((1, 8, 8), np.uint8))
    h = ctypes.c_void_p()
    st = _lib.load().sq_create(0, ctypes.byref(h))
    assert st != 0 and len(_lib.load().sq_last_error()) > 0


def test_product_never_imports_the_oracle():
    root = os.path.dirname(os.path.abspath(sequitr_b200.__file__))
    for dirpath, _, files in os.walk(root):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle', text, re.M), f
                assert 'oracle/_build' not in text and 'libsqref' not in text, f


def test_unet_constructor_defaults_and_errors():
    net = U.UNet({}, U.ModeKeys.PREDICT)
    assert net.filters == (16, 32, 64, 128, 256) and net.dropout == 0.4
    assert net.n_inputs == 1 and net.n_outputs == 2 and net.shape == (1024, 1024)
    assert net.bridge_type == 'eltwise_mul' and net.kernel == (3, 3)
    assert net.width == 1024 and net.height == 1024 and net.slices == 0 and net.ndim == 2
    assert not net.training and U.UNet({}, U.ModeKeys.TRAIN).training
    with pytest.raises(ValueError, match='Bridge type not recognized'):
        U.UNet({'bridge': 'bogus'})
    with pytest.raises(DeprecationWarning):
        net.btype
    x = U._Sym((1, 8, 8, 1))
    for fn, args in ((net.conv_layer, (x, 4)), (net.conv_layer_1x1, (x, 2)),
                     (net.conv_transpose_layer, (x, 4)), (net.max_pool_layer, (x,))):
        with pytest.raises(NotImplementedError):
            fn(*args)
    with pytest.raises(ValueError):
        U.UNet2D({'shape': (8, 8, 8)})
    with pytest.raises(ValueError):
        U.UNet3D({'shape': (8, 8)})


@pytest.mark.parametrize('bridge', U.BRIDGE_TYPES)
def test_unet_topology_trace_matches_reference_walk(bridge, monkeypatch):
    net = U.UNet2D({'filters': (4, 8, 16), 'shape': (16, 24), 'bridge': bridge, 'num_inputs': 3,
                    'num_outputs': 3})
    seen = {}

    def fake_execute(self, input_layer, want=('logits',)):
        seen['shape'] = tuple(input_layer.shape)
        return {'logits': np.zeros(input_layer.shape[:-1] + (3,), np.float32)}
    monkeypatch.setattr(U.UNet, '_execute', fake_execute)
    feats = np.zeros((2, 16 * 24 * 3), np.float32)
    logits = net.build(feats)
    assert seen['shape'] == (2, 16, 24, 3) and logits.shape == (2, 16, 24, 3)
    assert net.logits() is logits
    scopes = [t[1] for t in net._trace if t[0] != 'pool']
    assert scopes == synth.unet_layer_names((4, 8, 16))
    merged = [t for t in net._trace if t[1] == 'UNet/up0/conv1'][0]
    assert merged[2] == (8 if bridge == 'concat' else 4)
    assert len(net._net) == 3 + 2 + 1      # down layers, up layers, logits


def test_unet3d_reshape_and_trace(monkeypatch):
    net = U.UNet3D({'filters': (4, 8), 'shape': (16, 12, 8), 'bridge': 'concat'})
    assert net.slices == 8 and net.kernel == (3, 3, 3)
    monkeypatch.setattr(U.UNet, '_execute',
                        lambda self, x, want=('logits',): {'logits': np.zeros(x.shape[:-1] + (2,))})
    out = net.build(np.zeros((1, 8 * 16 * 12), np.float32))
    assert out.shape == (1, 8, 16, 12, 2)


def test_image_pipe_contract():
    class Double(pipeline.ImagePipe):
        def pipe(self, image):
            return image * 2
    out = Double()(np.ones((4, 5)))
    assert out.shape == (4, 5, 1) and out.dtype == np.float32
    with pytest.raises(NotImplementedError):
        pipeline.ImagePipe()(np.ones((2, 2)))
    with pytest.raises(TypeError):
        pipeline.ImagePipeline([Double(), 3])
    p = pipeline.ImagePipeline([Double(), Double()])
    assert len(p) == 1 and p(np.ones((2, 2))).max() == 4
    d = Double()
    d.update()
    assert d.iter == 0
    with pytest.raises(ValueError):
        pipeline._binary_plane(np.full((4, 4, 1), 0.5, np.float32), 'x')


def test_weightmap2_host_pipe_matches_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, 'weightmap_ref.npz'))
    got = pipeline.ImageWeightMap2(10., 5.)(g['in_discs64'])
    np.testing.assert_allclose(got, g['w2_discs64_w0-10_s-5'], atol=1e-12)


def test_create_weightmaps_files_and_names(tmp_path):
    import cv2
    lab_dir = tmp_path / 'setA' / 'label'
    lab_dir.mkdir(parents=True)
    mask = (synth.instance_labels(64, 64, 6, seed=1, rmin=4, rmax=8) > 0).astype(np.uint8) * 255
    cv2.imwrite(str(lab_dir / 'pos1_GFP_0001.tif'), mask)
    (lab_dir / 'notes.txt').write_text('x')
    out = weightmap.create_weightmaps(str(tmp_path), ['setA'], w0=10., sigma=5., method='delaunay')
    assert out == [str(tmp_path / 'setA' / 'weights_w0-10.00_sigma-5.00' / 'pos1_GFP_weights.tif')]
    w = cv2.imread(out[0], cv2.IMREAD_UNCHANGED)
    assert w.dtype == np.float32 and w.shape == (64, 64)
    want = pipeline.ImageWeightMap2(10., 5.)(mask > 0)[..., 0].astype(np.float32)
    np.testing.assert_array_equal(w, want)
    with pytest.raises(ValueError):
        weightmap.create_weightmaps(str(tmp_path), ['setA'], method='nope')


def test_image_labels():
    raw = np.zeros((3, 8, 8), np.uint8)
    raw[0, :4] = 5
    raw[2, 2:6] = 1
    il = weightmap.ImageLabels(raw)
    assert il.outputs == 4 and il.labels().dtype == np.uint8 and il.labels()[3, 0] == 3
    assert weightmap.ImageLabels(np.eye(4)).outputs == 2
    with pytest.raises(ValueError):
        weightmap.ImageLabels(np.zeros((5, 4, 4)))


def test_utils_helpers(tmp_path):
    assert utils.power_of_two(1024) and not utils.power_of_two(1200)
    assert utils.divisible_by_two_n_times(1200, 4) and not utils.divisible_by_two_n_times(1200, 5)
    with pytest.raises(IOError):
        utils.HDF5FileHandler('/nonexistent_dir_xyz/out.hdf5')
    with pytest.raises(TypeError):
        utils.HDF5FileHandler(None)
    h = utils.HDF5FileHandler(str(tmp_path / 'out.h5'))
    assert h.filename == str(tmp_path / 'out.hdf5')
    h.close()


def test_synthetic_generators_are_seeded():
    a = synth.frames(2, 64, 64, 1, seed=5, n_objects=4)
    b = synth.frames(2, 64, 64, 1, seed=5, n_objects=4)
    assert a.dtype == np.float32 and a.shape == (2, 64, 64, 1) and np.array_equal(a, b)
    lab = synth.instance_labels(128, 128, 12, seed=2, rmin=4, rmax=8)
    assert lab.max() >= 8
    from scipy.ndimage import label
    assert label(lab > 0)[1] == lab.max()          # >= 2 px gaps: binary and instance views agree
    w = synth.blob_detector_weights((8, 16), 1, 2)
    assert set(k.rsplit('/', 1)[0] for k in w) == set(synth.unet_layer_names((8, 16)))


def test_octopus_stream_reader(tmp_path):
    """dataio.OctopusData on a synthetic stream: file ranges, contiguous mode, headers, float frames
    (reference dataio/octopus.py:171-247) and raw batches across file boundaries."""
    from sequitr_b200.dataio import OctopusData, write_octopus_stream
    rng = np.random.default_rng(0)
    frames = rng.integers(0, 4096, size=(7, 12, 20), dtype=np.uint16)
    stem = str(tmp_path / 'BF_pos0_')
    assert write_octopus_stream(stem, frames, frames_per_file=3, extra={'X': np.arange(7) * 0.5}) == 3
    write_octopus_stream(stem, frames[:2], first_file=5)          # a later, non-contiguous file
    o = OctopusData(stem)
    assert len(o) == 7 and o.bit_depth == 16 and o.framesize == (12, 20) and o.filelist == [0, 1, 2]
    assert o.header_keys[:4] == ['N', 'H', 'W', 'Bit_Depth'] and o.header(0)['X'] == '0.0'
    for i in (0, 2, 3, 6):
        assert o[i].dtype == np.float64
        np.testing.assert_array_equal(o[i], frames[i].astype(float))
    assert o.info(4)['N'] == 4 and o.info(4)['X'] == '2.0'
    np.testing.assert_array_equal(o.frames_raw(1, 5), frames[1:6])
    assert o.frames_raw(5, 10).shape == (2, 12, 20) and o.frames_raw(0, 7).dtype == np.uint16
    assert len(OctopusData(stem, contiguous=False)) == 9
    with pytest.raises(IndexError):
        o[7]
    with pytest.raises(IOError):
        OctopusData(str(tmp_path / 'missing_'))
    write_octopus_stream(stem, frames[:1], first_file=3, age=0)   # still being written: ignored
    assert o.refresh() is False
    frames8 = rng.integers(0, 255, size=(2, 6, 6), dtype=np.uint8)
    write_octopus_stream(str(tmp_path / 'GFP_'), frames8)
    o8 = OctopusData(str(tmp_path / 'GFP_'))
    assert o8.bit_depth == 8
    np.testing.assert_array_equal(o8.to_array(), frames8)


def test_micromanager_position_reader(tmp_path):
    """dataio.MicromanagerReader on a synthetic Micro-Manager 2.0 position folder (reference
    dataio/micromanager.py:30-262) and the baseline TIFF reader it sits on."""
    from sequitr_b200.dataio import (MicromanagerReader, read_tiff, write_tiff, write_micromanager_position)
    rng = np.random.default_rng(1)
    frames = rng.integers(0, 65535, size=(5, 14, 22), dtype=np.uint16)
    gfp = rng.integers(0, 255, size=(5, 14, 22), dtype=np.uint8)
    pos = str(tmp_path / 'Pos0')
    write_micromanager_position(pos, frames, channel_index=0, n_channels=2)
    write_micromanager_position(pos, gfp, channel_index=1, n_channels=2)
    r = MicromanagerReader(pos, channel=0)
    assert len(r) == 5 and r.dtype == np.uint16 and (r.width, r.height) == (22, 14)
    for i in range(5):
        np.testing.assert_array_equal(r[i], frames[i])
    np.testing.assert_array_equal(r.frames_raw(1, 3), frames[1:4])
    assert r.frames_raw(3, 10).shape == (2, 14, 22)
    m = r.get_metadata(2)
    assert m['filename'].endswith('time000000002_z000.tif') and m['x_position'] == 20.0
    assert abs(m['timestamp'] - r.metadata.start_time - 3.275) < 1e-6
    assert r.metadata.shape == (5, 1, 22, 14)
    g = MicromanagerReader(pos, channel=1)
    assert g.dtype == np.uint8
    np.testing.assert_array_equal(g[4], gfp[4])
    assert len(MicromanagerReader(pos)) == 10                      # no channel filter: both channels
    with pytest.raises(IndexError):
        r[5]
    w = rng.standard_normal((9, 7)).astype(np.float32)             # float32 weight maps (weightmap.py:205)
    write_tiff(str(tmp_path / 'w.tif'), w)
    np.testing.assert_array_equal(read_tiff(str(tmp_path / 'w.tif')), w)
    big = struct_be_tiff(tmp_path, frames[0])
    np.testing.assert_array_equal(read_tiff(big), frames[0])


def struct_be_tiff(tmp_path, img):
    """A big-endian, two-strip TIFF written by hand (the reader must not depend on its own writer)."""
    import struct
    h, w = img.shape
    rows0 = h // 2
    s0 = img[:rows0].astype('>u2').tobytes()
    s1 = img[rows0:].astype('>u2').tobytes()
    off0, off1 = 8, 8 + len(s0)
    extra = off1 + len(s1)                       # strip offsets / counts arrays live after the pixel data
    ifd = extra + 16
    entries = [(256, 3, 1, w << 16), (257, 3, 1, h << 16), (258, 3, 1, 16 << 16), (259, 3, 1, 1 << 16),
               (273, 4, 2, extra), (277, 3, 1, 1 << 16), (278, 3, 1, rows0 << 16), (279, 4, 2, extra + 8)]
    path = str(tmp_path / 'be.tif')
    with open(path, 'wb') as fh:
        fh.write(b'MM' + struct.pack('>HI', 42, ifd) + s0 + s1)
        fh.write(struct.pack('>IIII', off0, off1, len(s0), len(s1)))
        fh.write(struct.pack('>H', len(entries)))
        for tag, typ, cnt, val in entries:
            fh.write(struct.pack('>HHII', tag, typ, cnt, val))
        fh.write(struct.pack('>I', 0))
    return path


def test_pinned_array_and_reader_destination(tmp_path):
    from sequitr_b200.dataio import OctopusData, write_octopus_stream
    import torch
    frames = np.arange(5 * 6 * 8, dtype=np.uint16).reshape(5, 6, 8)
    stem = str(tmp_path / 'S_')
    write_octopus_stream(stem, frames, frames_per_file=2)
    buf = np.zeros((4, 6, 8), np.uint16)
    if torch.cuda.is_available():                      # page-locking needs a CUDA driver
        buf = utils.pinned_array((4, 6, 8), np.uint16)
    got = OctopusData(stem).frames_raw(1, 3, out=buf)
    np.testing.assert_array_equal(got, frames[1:4])
    assert np.shares_memory(got, buf)
    with pytest.raises(ValueError):
        OctopusData(stem).frames_raw(0, 2, out=np.zeros((2, 6, 8), np.uint8))


def test_image_pipeline_json_round_trip(tmp_path):
    """pipeline.py:104-154: constructor arguments are read back from attributes of the same name, the
    file maps class names to them, loading rebuilds the chain.  No compute call is made here."""
    import json
    p = pipeline.ImagePipeline([pipeline.ImageOutliers(sigma=3, threshold=4.5), pipeline.ImageNorm(),
                                pipeline.ImageWeightMap(w0=8., sigma=2.5), pipeline.ImageSample(samples=4, ROI_size=(64, 64)),
                                pipeline.ImageFlip()])
    fn = str(tmp_path / 'pipe')
    p.save(fn)
    with open(fn + '.json') as f:
        d = json.load(f)
    assert list(d['ImagePipeline']) == ['ImageOutliers', 'ImageNorm', 'ImageWeightMap', 'ImageSample', 'ImageFlip']
    assert d['ImagePipeline']['ImageWeightMap'] == {'w0': 8.0, 'sigma': 2.5}
    q = pipeline.ImagePipeline.load(fn + '.json')
    assert [type(a) for a in q.pipeline] == [type(a) for a in p.pipeline]
    assert q.pipeline[0].sigma == 3 and q.pipeline[0].threshold == 4.5 and q.pipeline[2].w0 == 8.0
    assert q.pipeline[3].samples == 4 and tuple(q.pipeline[3].ROI_size) == (64, 64)
    assert len(q) == 1 * 1 * 1 * 4 * 4
    with pytest.raises(TypeError):
        pipeline.save_image_pipeline(fn, [1, 2])
    # the index-only pipes: flips cycle with update(), samples reuse their coordinates
    img = np.arange(12, dtype=np.float32).reshape(3, 4)
    flip = pipeline.ImageFlip()
    outs = []
    for _ in range(4):
        outs.append(flip(img)[..., 0].copy())
        flip.update()
    assert np.array_equal(outs[0], img) and np.array_equal(outs[1], img[:, ::-1])
    assert np.array_equal(outs[2], img[::-1]) and np.array_equal(outs[3], img[::-1, ::-1])
    big = np.random.default_rng(0).random((100, 120)).astype(np.float32)
    smp = pipeline.ImageSample(samples=3, ROI_size=(16, 16))
    a = smp(big)
    b = smp(big * 2)
    assert a.shape == (3, 16, 16, 1) and np.allclose(b, 2 * a)
    (x, y) = smp.coords[0]
    assert np.allclose(a[0, ..., 0], big[x - 8:x + 8, y - 8:y + 8])


def test_parity_helpers_on_the_two_oracle_contracts():
    """sequitr_b200.parity on real data without a GPU: the oracle's bf16 contract stands in for the
    tensor-core path, its fp32 contract is the reference arithmetic.  Mask flips are near-ties only and
    the centroid rows pair up one to one."""
    from oracle import unet_c, centroid_oracle
    from sequitr_b200 import parity
    filters = (16, 32, 64)
    x = synth.frames(2, 96, 128, 1, seed=3, n_objects=6)
    w = synth.blob_detector_weights(filters, 1, 2, seed=1)
    a = unet_c.unet_forward(x, w, filters, 'concat', contract='bf16')
    b = unet_c.unet_forward(x, w, filters, 'concat', contract='fp32')
    mp = parity.mask_parity(a['mask'], b['mask'], b['logits'])
    assert mp['pixels'] == 2 * 96 * 128 and mp['mismatch'] == int((a['mask'] != b['mask']).sum())
    assert mp['mismatch_frac'] <= 1e-3 and mp['max_margin_of_mismatch'] <= 0.01 * mp['logit_range']
    same = parity.mask_parity(b['mask'], b['mask'], b['logits'])
    assert same['mismatch'] == 0 and same['max_margin_of_mismatch'] == 0.0
    ta, tb = centroid_oracle.centroid_tables(a['mask']), centroid_oracle.centroid_tables(b['mask'])
    cd = parity.centroid_set_diff(ta, tb)
    assert cd['rows'] == cd['ref_rows'] == sum(len(t) for t in tb) and cd['unmatched'] == cd['ref_unmatched'] == 0
    assert cd['identical'] + cd['moved'] == cd['rows'] and cd['max_shift_px'] <= 0.5
    assert parity.centroid_set_diff(tb, tb)['rows_changed'] == 0
    # an object that vanishes and one that moves by more than the tolerance
    broken = [t.copy() for t in tb]
    broken[0] = broken[0][1:]
    broken[1][0, 1] += 3.0
    cd = parity.centroid_set_diff(broken, tb)
    assert cd['ref_unmatched'] == 2 and cd['unmatched'] == 1 and cd['rows_changed'] == 3
    with pytest.raises(ValueError):
        parity.mask_parity(a['mask'][:1], b['mask'], b['logits'])
    # K > 2 margins
    lg = np.array([[[0.0, 2.0, 1.5]]], np.float32)
    assert abs(parity.top2_margin(lg)[0, 0] - 0.5) < 1e-7


def test_camera_stack_is_a_function_of_the_global_frame_index():
    """synth.camera_stack: forked renderers write one shared mapping; any shard of the stack equals the
    same frames rendered alone (what makes the per-rank slices of bench.py's 2000-frame stack consistent)."""
    whole = synth.camera_stack(0, 11, 48, 64, seed=5, workers=3, chunk=2)
    assert whole.shape == (11, 48, 64) and whole.dtype == np.uint16
    part = synth.camera_stack(4, 9, 48, 64, seed=5, workers=1)
    np.testing.assert_array_equal(part, whole[4:9])
    one = synth.to_camera_counts(synth.frames(1, 48, 64, 1, seed=5, first_frame=7)[..., 0])
    np.testing.assert_array_equal(one[0], whole[7])
    assert len(synth.camera_stack(3, 3, 48, 64)) == 0


def test_segment_stack_and_digest_with_a_stub_network():
    """shard.segment_stack walks a rank's range call by call with GLOBAL frame indices; tables_digest is
    independent of how the stack was split into calls / ranks."""
    from sequitr_b200 import shard

    class Stub(object):
        def __init__(self):
            self.calls = []

        def segment_and_localise(self, frames, frame0=0, max_rows=4096, normalise=False):
            self.calls.append((len(frames), frame0, normalise))
            return [np.full((int(f[0, 0]) % 3, 5), frame0 + i, np.float32) for i, f in enumerate(frames)]

    stack = np.arange(23, dtype=np.uint16)[:, None, None] * np.ones((1, 2, 2), np.uint16)
    s1 = Stub()
    whole = shard.segment_stack(s1, stack, frame0=0, frames_per_call=10)
    assert s1.calls == [(10, 0, True), (10, 10, True), (3, 20, True)]
    parts = []
    for r in range(4):
        lo, hi = shard.frame_range(r, 4, 23)
        parts.append(shard.segment_stack(Stub(), stack[lo:hi], frame0=lo, frames_per_call=4))
    merged = shard.merge_tables(parts)
    assert len(merged) == 23 and all((t[:, 0] == i).all() for i, t in enumerate(merged))
    assert shard.tables_digest(merged) == shard.tables_digest(whole)
    merged[5] = merged[5] + 1
    assert shard.tables_digest(merged) != shard.tables_digest(whole)
    # overlap=True: two host threads alternate the calls (the odd ones through net.twin()); same tables, same order
    s2, s3 = Stub(), Stub()
    s2.twin = lambda: s3
    over = shard.segment_stack(s2, stack, frame0=0, frames_per_call=4, overlap=True)
    assert shard.tables_digest(over) == shard.tables_digest(whole)
    assert [c[1] for c in s2.calls] == [0, 8, 16] and [c[1] for c in s3.calls] == [4, 12, 20]


def test_trainer_label_layouts():
    """`Trainer._class_ids`: what tr_augment yields (one-hot uint8, weights with a trailing axis) and plain class ids
    reach the C ABI as (class ids uint8, weights float32); an all-zero one-hot row clears its weight."""
    import numpy as np
    import pytest
    import torch
    from sequitr_b200.networks.unet import Trainer
    sp = (2, 4, 6)
    rng = np.random.default_rng(0)
    ids = rng.integers(0, 3, sp).astype(np.uint8)
    wgt = (1 + rng.random(sp)).astype(np.float32)
    onehot = np.eye(3, dtype=np.uint8)[ids]
    for lab_in, w_in in ((ids, wgt), (ids[..., None], wgt[..., None]), (onehot, wgt[..., None])):
        lab, w = Trainer._class_ids(torch.from_numpy(lab_in.copy()), torch.from_numpy(w_in.copy()), sp, 3)
        assert lab.dtype == torch.uint8 and tuple(lab.shape) == sp and lab.is_contiguous()
        assert w.dtype == torch.float32 and tuple(w.shape) == sp
        assert np.array_equal(lab.numpy(), ids) and np.array_equal(w.numpy(), wgt)
    # a pixel of a class beyond the K outputs: tr_augment's one-hot row is all zero (networks/unet.py:396-398)
    onehot2 = onehot[..., :2].copy()
    lab, w = Trainer._class_ids(torch.from_numpy(onehot2), torch.from_numpy(wgt.copy()), sp, 2)
    beyond = ids == 2
    assert beyond.any() and (w.numpy()[beyond] == 0).all() and np.array_equal(w.numpy()[~beyond], wgt[~beyond])
    assert (lab.numpy()[~beyond] == ids[~beyond]).all() and int(lab.max()) < 2
    with pytest.raises(ValueError):
        Trainer._class_ids(torch.from_numpy(ids), torch.from_numpy(wgt), sp, 2)          # class id 2 >= 2 outputs
    with pytest.raises(ValueError):
        Trainer._class_ids(torch.from_numpy(ids[:, :3]), torch.from_numpy(wgt), sp, 3)   # label shape
    with pytest.raises(ValueError):
        Trainer._class_ids(torch.from_numpy(ids), torch.from_numpy(wgt[:, :3]), sp, 3)   # weight shape
