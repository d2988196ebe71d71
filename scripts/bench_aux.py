"""Memory-bound kernels of the hot path at BASELINE's sizes: achieved algorithmic GB/s vs the
measured HBM peak (CUDA events, inputs resident in HBM, 16 masks of 2048^2 per launch so that
every launch moves more than the 126 MB L2 can hold for the float outputs)."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sequitr_b200 import synth, ops          # noqa: E402

peaks = {}
try:
    peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                        'MEASURED_PEAKS.json')))
except Exception:
    pass
PEAK = peaks.get('hbm_gbs', 6650.0)
N, H, W = int(os.environ.get('N', 16)), 2048, 2048
labs = np.stack([synth.instance_labels(H, W, 600, seed=s) for s in range(4)])
labs = np.concatenate([labs] * (N // 4)).astype(np.int32)
lab_d = torch.from_numpy(labs).cuda()
mask_d = (lab_d > 0).to(torch.uint8).contiguous()
cls_d = torch.from_numpy(np.where(labs > 0, 1 + (labs - 1) % 2, 0).astype(np.uint8)).cuda()


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


frames_d = torch.from_numpy(synth.frames(4, H, W, 1, seed=5)).cuda().repeat(N // 4, 1, 1, 1).contiguous()
px = N * H * W
rows = []
for name, fn, bpp in (
        ('weightmap_edt  W1 (u8 -> f32)', lambda: ops.weightmap_edt(mask_d, 10., 5., 'float32'), 5),
        ('weightmap_edt  W1 (u8 -> f64)', lambda: ops.weightmap_edt(mask_d, 10., 5., 'float64'), 9),
        ('weightmap_unet W3 (i32 -> f32)', lambda: ops.weightmap_unet(lab_d, 10., 5., None, 'float32'), 8),
        ('ImageNorm      (f32 -> f32)', lambda: ops.image_norm(frames_d), 8),
        ('ImageOutliers  (f32 -> f32, 2x2 median)', lambda: ops.image_outliers(frames_d, 2, 5.), 8),
        ('ImageBGSubtract (f32 -> f32)', lambda: ops.image_bgsubtract(frames_d, 'float32'), 8),
        ('label_centroids L1 (u8 -> rows)', lambda: ops.label_centroids(cls_d, max_rows=2048), 1),
        ('label_centroids L1 (+ label matrix)', lambda: ops.label_centroids(cls_d, max_rows=2048, want_labels=True), 5)):
    ms = timeit(fn)
    gbs = px * bpp / (ms * 1e-3) / 1e9
    rows.append({'kernel': name, 'ms_per_launch': ms, 'us_per_frame': 1e3 * ms / N, 'bytes_per_px': bpp,
                 'achieved_GBs': gbs, 'frac_of_hbm_peak': gbs / PEAK})
    print('%-38s %8.3f ms  %7.1f us/frame  %2d B/px  %8.1f GB/s  %5.1f%% of %.0f GB/s'
          % (name, ms, 1e3 * ms / N, bpp, gbs, 100 * gbs / PEAK, PEAK))
# tr_augment: 2048^2 frames -> 1024^2 rotated crops, bytes counted per OUTPUT pixel (9 gathered in, 10 out)
lab8_d = (lab_d % 5).to(torch.uint8).contiguous()
wgt_d = ops.weightmap_edt(mask_d, 10., 5., 'float32')
rs = np.random.RandomState(0)
tr = np.stack([ops.rotation_transform(2 * np.pi * rs.uniform(), H, W) for _ in range(N)])
cr = rs.randint(0, 1024, (N, 2)).astype(np.int32)
ms = timeit(lambda: ops.tr_augment(frames_d, lab8_d, wgt_d, tr, cr, 1024, 1024, 2))
gbs = N * 1024 * 1024 * 19 / (ms * 1e-3) / 1e9
rows.append({'kernel': 'tr_augment 2048^2 -> 1024^2 (rotate+crop+one-hot)', 'ms_per_launch': ms, 'us_per_frame': 1e3 * ms / N,
             'bytes_per_px': 19, 'achieved_GBs': gbs, 'frac_of_hbm_peak': gbs / PEAK})
print('%-38s %8.3f ms  %7.1f us/frame  %2d B/px  %8.1f GB/s  %5.1f%% of %.0f GB/s'
      % ('tr_augment (rotate+crop+one-hot)', ms, 1e3 * ms / N, 19, gbs, 100 * gbs / PEAK, PEAK))
print(json.dumps({'frames_per_launch': N, 'hbm_peak_GBs': PEAK, 'kernels': rows}))
