#!/usr/bin/env python
"""Freeze golden vectors under tests/golden/ (run in the build container only).

* ``weightmap_ref.npz``  -- inputs + outputs of the REFERENCE's own
  ``pipeline.ImageWeightMap`` / ``ImageWeightMap2`` (imported from
  /root/reference by oracle/ref_loader.py).  These pin the oracle and, through
  it, the CUDA kernels to the real reference code.
* ``centroid_scipy.npz`` -- masks + per-frame centroid tables from the SciPy
  restatement of ``utils.CentroidWriter.write`` (the reference file is py2-only
  and cannot run; SciPy is the reference's own dependency for this path).
* ``unet_kat.npz``       -- a small known-answer test of the plain-C UNet oracle
  (seeded input/weights -> logits/mask), to detect drift of the numeric contract.

* ``prep_ref.npz``       -- inputs + outputs of the REFERENCE's ``ImageNorm`` / ``ImageOutliers`` /
  ``ImageBGSubtract`` pipes (pre-inference clean-up).

* ``augment_kat.npz``    -- a known-answer test of the ``tr_augment`` oracle (TensorFlow's projective
  resampling restated; the reference's dependency is absent), to detect drift of the contract.

The /root/reference tree does not travel to the GPU box; these files do.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_loader, centroid_oracle, unet_c   # noqa: E402
from sequitr_b200 import synth                           # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden')


def weightmap_cases():
    rng = np.random.default_rng(7)
    cases = {}
    cases['discs64'] = synth.instance_labels(64, 64, 6, seed=1, rmin=4, rmax=8) > 0
    cases['discs96x80'] = synth.instance_labels(96, 80, 10, seed=2, rmin=4, rmax=9) > 0
    m = np.zeros((40, 56), bool)
    m[3, 5] = True
    m[30:36, 40:50] = True
    m[0, :] = True                      # object touching the frame border
    cases['border40x56'] = m
    cases['noise48'] = rng.random((48, 48)) > 0.9
    cases['allfg16'] = np.ones((16, 16), bool)
    cases['allbg16x24'] = np.zeros((16, 24), bool)   # SciPy's "virtual seed at (-1,0)" case
    one = np.zeros((33, 31), bool)
    one[16, 15] = True
    cases['single33x31'] = one
    return cases


def prep_cases():
    rng = np.random.default_rng(11)
    cases = {}
    yy, xx = np.mgrid[0:72, 0:96]
    cam = 120 + 0.05 * xx + 0.11 * yy + 4e-4 * xx * yy - 3e-4 * yy * yy + rng.standard_normal((72, 96)) * 3
    cam[10, 20] += 400.0                 # hot pixels, one on the border, one in a corner
    cam[0, 50] += 250.0
    cam[71, 95] -= 300.0
    cases['camera72x96'] = cam.astype(np.float32)
    cases['noise33x31'] = (rng.standard_normal((33, 31)) * 7 + 50).astype(np.float32)
    rgb = (rng.standard_normal((40, 48, 3)) * np.array([1., 5., 20.]) + np.array([0., 10., -30.])).astype(np.float32)
    rgb[7, 9, 1] = 300.0
    cases['rgb40x48x3'] = rgb
    return cases


def augment_kat():
    from oracle import augment_oracle as ao
    rng = np.random.default_rng(31)
    n, h, w, c, ch, cw, k = 4, 40, 52, 2, 24, 32, 3
    image = rng.standard_normal((n, h, w, c)).astype(np.float32)
    label = rng.integers(0, 4, (n, h, w)).astype(np.uint8)
    weights = rng.uniform(1, 11, (n, h, w)).astype(np.float32)
    theta = np.array([0.0, 0.4, 2.2, 5.0], dtype=np.float32)
    crop = np.array([[0, 0], [16, 20], [3, 7], [10, 1]], dtype=np.int32)
    outs = [ao.tr_augment(image[i], label[i], weights[i], theta[i], crop[i, 0], crop[i, 1], ch, cw, k) for i in range(n)]
    np.savez_compressed(os.path.join(OUT, 'augment_kat.npz'), image=image, label=label, weights=weights, theta=theta,
                        crop=crop, ch=ch, cw=cw, k=k, image_out=np.stack([o[0] for o in outs]),
                        label_out=np.stack([o[1] for o in outs]), weights_out=np.stack([o[2] for o in outs]))
    print('augment_kat.npz')


def main():
    if sys.argv[1:] == ['augment']:
        return augment_kat()
    os.makedirs(OUT, exist_ok=True)
    ref = ref_loader.load_reference_pipeline()

    pr = {}
    for name, img in prep_cases().items():
        pr['in_' + name] = img
        pr['norm_' + name] = ref.ImageNorm()(img.copy())
        pr['outl2_' + name] = ref.ImageOutliers()(img.copy())
        pr['outl3_' + name] = ref.ImageOutliers(sigma=3, threshold=2.)(img.copy())
        if img.ndim == 2:
            pr['bg_' + name] = ref.ImageBGSubtract()(img.copy())
    np.savez_compressed(os.path.join(OUT, 'prep_ref.npz'), **pr)
    print('prep_ref.npz', len(pr), 'arrays')

    wm = {}
    for name, mask in weightmap_cases().items():
        wm['in_' + name] = mask
        for (w0, sigma) in ((10., 5.), (30., 3.)):
            key = '%s_w0-%g_s-%g' % (name, w0, sigma)
            wm['w1_' + key] = ref.ImageWeightMap(w0=w0, sigma=sigma)(mask.copy())
    # ImageWeightMap2 needs a non-degenerate triangulation: discs only
    for name in ('discs64', 'discs96x80'):
        mask = wm['in_' + name]
        wm['w2_%s_w0-10_s-5' % name] = ref.ImageWeightMap2(w0=10., sigma=5.)(mask.copy())
    np.savez_compressed(os.path.join(OUT, 'weightmap_ref.npz'), **wm)

    ce = {}
    stacks = {
        'discs2c': np.stack([synth.class_mask(96, 128, 14, n_classes=3, seed=s, rmin=4, rmax=9)
                             for s in (3, 4, 5)]),
        'noise': (np.random.default_rng(11).random((2, 40, 40)) > 0.6).astype(np.uint8) *
                 np.random.default_rng(12).integers(1, 4, (2, 40, 40)).astype(np.uint8),
        'empty': np.zeros((2, 16, 16), np.uint8),
    }
    u = np.zeros((1, 12, 12), np.uint8)           # U-shape: provisional labels merge
    u[0, 2:10, 2] = 1
    u[0, 2:10, 8] = 1
    u[0, 9, 2:9] = 1
    u[0, 0, 0] = 1                                # single pixel, diagonal to nothing
    u[0, 3, 4] = 2
    u[0, 4, 5] = 2                                # diagonal contact must NOT connect
    stacks['ushape'] = u
    vol = (np.random.default_rng(13).random((2, 6, 10, 12)) > 0.7).astype(np.uint8)
    stacks['vol'] = vol
    for name, st in stacks.items():
        ce['in_' + name] = st
        tabs = centroid_oracle.centroid_tables(st)
        ce['counts_' + name] = np.array([len(t) for t in tabs], np.int32)
        ce['table_' + name] = (np.concatenate(tabs, 0) if sum(len(t) for t in tabs)
                               else np.zeros((0, 5), np.float32))
    np.savez_compressed(os.path.join(OUT, 'centroid_scipy.npz'), **ce)

    filters = (8, 16, 32)
    w = synth.unet_weights(filters, 1, 2, ndim=2, bridge='concat', seed=5)
    x = synth.frames(1, 32, 48, 1, seed=9, n_objects=3)
    r = unet_c.unet_forward(x, w, filters, 'concat')
    rb = unet_c.unet_forward(x, w, filters, 'concat', contract='bf16')
    np.savez_compressed(os.path.join(OUT, 'unet_kat.npz'), x=x, logits=r['logits'],
                        mask=r['mask'], logits_bf16=rb['logits'], filters=np.array(filters))
    augment_kat()
    print('golden vectors written to', OUT)


if __name__ == '__main__':
    main()
