"""Tiny driver for ncu: W1 and W3 weight maps on 4 label frames of 2048^2."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sequitr_b200 import synth, ops
labs = np.stack([synth.instance_labels(2048, 2048, 600, seed=s) for s in range(4)]).astype(np.int32)
lab_d = torch.from_numpy(labs).cuda()
mask_d = (lab_d > 0).to(torch.uint8).contiguous()
for _ in range(3):
    ops.weightmap_edt(mask_d, 10., 5., 'float32')
    ops.weightmap_unet(lab_d, 10., 5., None, 'float32')
torch.cuda.synchronize()
