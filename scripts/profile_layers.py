"""Per-layer device times of the bf16 tensor-core UNet at BASELINE's frame size."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sequitr_b200 import synth                      # noqa: E402
from sequitr_b200.networks import UNet2D            # noqa: E402

n, h, w = int(os.environ.get('N', 4)), int(os.environ.get('H', 2048)), int(os.environ.get('W', 2048))
filters = (16, 32, 64, 128, 256)
peaks = {}
try:
    peaks = json.load(open('MEASURED_PEAKS.json'))
except Exception:
    pass
peak = peaks.get('bf16_tflops', 1590.0)
net = UNet2D({'filters': filters, 'shape': (h, w), 'bridge': 'concat', 'compute': 'bf16'})
net.load_weights(synth.blob_detector_weights(filters, 1, 2, seed=1))
x = torch.randn((n, h, w, 1), device='cuda')
for _ in range(3):
    net.predict(x, want=('mask',))
torch.cuda.synchronize()
rows = net.profile(x)
rows = net.profile(x)
tot = sum(r[1] for r in rows)
print('%-22s %9s %9s %8s' % ('layer', 'ms', 'TFLOP/s', '%peak'))
for name, ms, fl in rows:
    tf = fl / (ms * 1e-3) / 1e12 if ms > 0 else 0
    print('%-22s %9.3f %9.1f %8.1f' % (name, ms, tf, 100 * tf / peak))
conv_ms = sum(r[1] for r in rows if r[2] > 0)
conv_fl = sum(r[2] for r in rows)
print('total %.3f ms for %d frames -> %.1f frames/s; dense layers %.3f ms, %.1f TFLOP/s = %.1f%% of %.0f'
      % (tot, n, n / (tot * 1e-3), conv_ms, conv_fl / (conv_ms * 1e-3) / 1e12,
         100 * conv_fl / (conv_ms * 1e-3) / 1e12 / peak, peak))
print('launches', net.launches())
