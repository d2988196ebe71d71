"""One pass over the memory-bound side kernels (4 frames of 2048^2) for ncu captures."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sequitr_b200 import synth, ops          # noqa: E402

N, H, W = 4, 2048, 2048
labs = np.stack([synth.instance_labels(H, W, 600, seed=s) for s in range(N)]).astype(np.int32)
lab_d = torch.from_numpy(labs).cuda()
mask_d = (lab_d > 0).to(torch.uint8).contiguous()
cls_d = torch.from_numpy(np.where(labs > 0, 1 + (labs - 1) % 2, 0).astype(np.uint8)).cuda()
frames_d = torch.from_numpy(synth.frames(N, H, W, 1, seed=5)).cuda()
lab8_d = (lab_d % 5).to(torch.uint8).contiguous()
rs = np.random.RandomState(0)
tr = np.stack([ops.rotation_transform(2 * np.pi * rs.uniform(), H, W) for _ in range(N)])
cr = rs.randint(0, 1024, (N, 2)).astype(np.int32)
for rep in range(int(os.environ.get('REPS', '2'))):
    wgt_d = ops.weightmap_edt(mask_d, 10., 5., 'float32')
    ops.weightmap_unet(lab_d, 10., 5., None, 'float32')
    ops.image_norm(frames_d)
    ops.image_outliers(frames_d, 2, 5.)
    ops.image_bgsubtract(frames_d, 'float32')
    ops.label_centroids(cls_d, max_rows=2048)
    ops.tr_augment(frames_d, lab8_d, wgt_d, tr, cr, 1024, 1024, 2)
torch.cuda.synchronize()
print('aux_run ok')
