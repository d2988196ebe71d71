"""tr_augment on the GPU (sq_tr_augment through the C ABI) against the oracle: bit-exact --
the coordinate arithmetic is float32 without fused multiply-add on both sides."""
import os

import numpy as np
import pytest

from oracle import augment_oracle as ao

pytestmark = pytest.mark.gpu


def _stack(seed, n, h, w, c):
    rng = np.random.default_rng(seed)
    img = rng.standard_normal((n, h, w, c)).astype(np.float32)
    lab = rng.integers(0, 5, (n, h, w)).astype(np.uint8)
    wgt = rng.uniform(1, 11, (n, h, w)).astype(np.float32)
    return img, lab, wgt


def _run(sq, img, lab, wgt, thetas, crops, ch, cw, k):
    import torch
    from sequitr_b200 import ops
    n, h, w, c = img.shape
    tr = np.stack([ops.rotation_transform(t, h, w) for t in thetas])
    got = ops.tr_augment(torch.from_numpy(img).cuda(), torch.from_numpy(lab).cuda(), torch.from_numpy(wgt).cuda(),
                         tr, crops, ch, cw, k)
    return [g.cpu().numpy() for g in got]


@pytest.mark.parametrize('h,w,c,ch,cw,k', [(64, 64, 1, 32, 32, 2), (75, 100, 3, 48, 64, 3), (33, 41, 2, 33, 41, 5),
                                          (128, 96, 1, 1, 1, 1)])
def test_augment_matches_oracle_bit_exact(sq, h, w, c, ch, cw, k):
    n = 5
    img, lab, wgt = _stack(h * w + c, n, h, w, c)
    rng = np.random.default_rng(9)
    thetas = np.concatenate([[0.0, np.pi / 2], 2 * np.pi * rng.uniform(size=n - 2)]).astype(np.float32)
    crops = np.stack([rng.integers(0, h - ch + 1, n), rng.integers(0, w - cw + 1, n)], 1).astype(np.int32)
    gi, gl, gw = _run(sq, img, lab, wgt, thetas, crops, ch, cw, k)
    for i in range(n):
        ri, rl, rw_ = ao.tr_augment(img[i], lab[i], wgt[i], thetas[i], crops[i, 0], crops[i, 1], ch, cw, k)
        assert np.array_equal(gi[i], ri)
        assert np.array_equal(gl[i], rl)
        assert np.array_equal(gw[i], rw_)


def test_augment_more_frames_than_one_launch(sq):
    n, h, w = 70, 40, 40           # the kernel takes 32 frames' parameters per launch
    img, lab, wgt = _stack(5, n, h, w, 1)
    rng = np.random.default_rng(10)
    thetas = (2 * np.pi * rng.uniform(size=n)).astype(np.float32)
    crops = rng.integers(0, 9, (n, 2)).astype(np.int32)
    gi, gl, gw = _run(sq, img, lab, wgt, thetas, crops, 32, 32, 2)
    for i in (0, 31, 32, 33, 63, 64, 69):
        ri, rl, rw_ = ao.tr_augment(img[i], lab[i], wgt[i], thetas[i], crops[i, 0], crops[i, 1], 32, 32, 2)
        assert np.array_equal(gi[i], ri) and np.array_equal(gl[i], rl) and np.array_equal(gw[i], rw_)


def test_augment_golden(sq, golden_dir):
    g = np.load(os.path.join(golden_dir, 'augment_kat.npz'))
    gi, gl, gw = _run(sq, g['image'], g['label'], g['weights'], g['theta'], g['crop'], int(g['ch']), int(g['cw']),
                      int(g['k']))
    assert np.array_equal(gi, g['image_out']) and np.array_equal(gl, g['label_out']) and np.array_equal(gw, g['weights_out'])


def test_augment_full_size_properties(sq):
    """2048^2 -> 1024^2 crops (the size class of the BASELINE configs): identity angle = plain crop,
    quarter turn = rot90 of the labels, one class per pixel at an arbitrary angle."""
    import torch
    from sequitr_b200 import ops
    n, h, w, ch, cw = 2, 2048, 2048, 1024, 1024
    img, lab, wgt = _stack(77, n, h, w, 1)
    crops = np.array([[100, 900], [1024, 0]], dtype=np.int32)
    gi, gl, gw = _run(sq, img, lab, wgt, np.zeros(n, np.float32), crops, ch, cw, 5)
    for i in range(n):
        r0, c0 = crops[i]
        assert np.array_equal(gi[i], img[i, r0:r0 + ch, c0:c0 + cw])
        assert np.array_equal(gw[i], wgt[i, r0:r0 + ch, c0:c0 + cw])
        assert np.array_equal(gl[i].argmax(-1), lab[i, r0:r0 + ch, c0:c0 + cw])
    gi, gl, gw = _run(sq, img, lab, wgt, np.full(n, np.pi / 2, np.float32), crops, ch, cw, 5)
    for i in range(n):
        r0, c0 = crops[i]
        a, b = np.rot90(lab[i], 1), np.rot90(lab[i], -1)
        got = gl[i].argmax(-1)
        assert np.array_equal(got, a[r0:r0 + ch, c0:c0 + cw]) or np.array_equal(got, b[r0:r0 + ch, c0:c0 + cw])
    gi, gl, gw = _run(sq, img, lab, wgt, np.full(n, 0.6, np.float32), crops, ch, cw, 5)
    # every pixel carries exactly one class; weights stay in [0, 12] (input range 1..11, +1 outside, and a
    # bilinear fade towards the zero fill on the one-pixel rim of the rotated frame, see the oracle's notes)
    assert (gl.sum(-1) == 1).all() and gw.min() >= 0.0 and gw.max() <= 12.0
    assert (gw[gl[..., 0] == 1] >= 0).all()


def test_reference_entry_point(sq):
    """networks.unet.tr_augment keeps the reference's call signature and return layout."""
    from sequitr_b200.networks import unet
    n, h, w = 3, 96, 80
    img, lab, wgt = _stack(21, n, h, w, 1)
    feats = {'image': img, 'label': lab[..., None], 'weights': wgt[..., None], 'shape': (n, h, w, 1)}
    out, d = unet.tr_augment(feats, {'shape': (64, 48), 'num_outputs': 2}, rng=np.random.RandomState(4))
    assert tuple(out.shape) == (n, 64, 48, 1) and tuple(d['label'].shape) == (n, 64, 48, 2)
    assert tuple(d['weights'].shape) == (n, 64, 48, 1)
    rs = np.random.RandomState(4)
    thetas = [np.float32(2.) * np.float32(rs.uniform()) * np.float32(np.pi) for _ in range(n)]
    crops = [(rs.randint(0, h - 64), rs.randint(0, w - 48)) for _ in range(n)]
    for i in range(n):
        ri, rl, rw_ = ao.tr_augment(img[i], lab[i], wgt[i], thetas[i], crops[i][0], crops[i][1], 64, 48, 2)
        assert np.array_equal(out[i].cpu().numpy(), ri)
        assert np.array_equal(d['label'][i].cpu().numpy(), rl)
        assert np.array_equal(d['weights'][i, ..., 0].cpu().numpy(), rw_)
    with pytest.raises(ValueError):
        unet.tr_augment(feats, {'shape': (128, 48)})
    assert unet.preprocess_norm(feats) is feats


def test_augment_rejects_bad_arguments(sq):
    img, lab, wgt = _stack(1, 1, 16, 16, 1)
    with pytest.raises(ValueError):
        _run(sq, img, lab, wgt, [0.0], np.array([[8, 0]], np.int32), 12, 12, 2)     # crop leaves the image
    with pytest.raises(ValueError):
        _run(sq, img, lab, wgt, [0.0], np.array([[0, 0]], np.int32), 8, 8, 6)       # > 5 label channels
