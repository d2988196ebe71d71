"""Tiny driver for ncu captures: a few UNet + label-and-localise passes on 2 frames of 2048^2."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sequitr_b200 import synth, ops
from sequitr_b200.networks import UNet2D
filters = (16, 32, 64, 128, 256)
n = int(os.environ.get('N', 2))
net = UNet2D({'filters': filters, 'shape': (2048, 2048), 'bridge': 'concat', 'compute': 'bf16'})
net.load_weights(synth.blob_detector_weights(filters, 1, 2, seed=1))
x = torch.from_numpy(synth.frames(n, 2048, 2048, 1, seed=1234)).cuda()
for _ in range(3):
    mask = net.predict(x, want=('mask',))['mask']
    table, counts = ops.label_centroids(mask, max_rows=2048)
torch.cuda.synchronize()
print('objects', counts.tolist())
