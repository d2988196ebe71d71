"""Diagnostics: per-role wait/work clocks of the x-combined conv kernels (SQ_XC_PHASE=1)."""
import os, sys
os.environ.setdefault('SQ_XC_PHASE', '1')
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sequitr_b200 import synth
from sequitr_b200.networks import UNet2D, UNet3D
filters = (16, 32, 64, 128, 256)
if os.environ.get('VOL', '1') == '1':
    net = UNet3D({'filters': filters, 'shape': (1024, 1024, 32), 'bridge': 'concat', 'compute': 'bf16'})
    net.load_weights(synth.unet_weights(filters, 1, 2, ndim=3, bridge='concat', seed=1))
    x = torch.randn((1, 32, 1024, 1024, 1), device='cuda')
else:
    net = UNet2D({'filters': filters, 'shape': (2048, 2048), 'bridge': 'concat', 'compute': 'bf16'})
    net.load_weights(synth.blob_detector_weights(filters, 1, 2, seed=1))
    x = torch.randn((4, 2048, 2048, 1), device='cuda')
for i in range(2):
    print('pass', i, file=sys.stderr)
    net.predict(x, want=('mask',))
torch.cuda.synchronize()
