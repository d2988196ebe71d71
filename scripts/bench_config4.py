"""BASELINE configs[4]: UNet3D on a 64x1024x1024 z-stack (bf16 tensor-core mode) + one weighted
cross-entropy step using the GPU weight map.  Per-layer device times (CUDA events) and totals."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sequitr_b200 import synth, ops                 # noqa: E402
from sequitr_b200.networks import UNet3D            # noqa: E402

d, h, w = int(os.environ.get('D', 64)), int(os.environ.get('H', 1024)), int(os.environ.get('W', 1024))
filters = (16, 32, 64, 128, 256)
peaks = {}
try:
    peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                        'MEASURED_PEAKS.json')))
except Exception:
    pass
peak = peaks.get('bf16_tflops_sustained', peaks.get('bf16_tflops', 1382.1))
net = UNet3D({'filters': filters, 'shape': (h, w, d), 'bridge': 'concat', 'compute': 'bf16'})
net.load_weights(synth.unet_weights(filters, 1, 2, ndim=3, bridge='concat', seed=1))
x = torch.randn((1, d, h, w, 1), device='cuda')


def timeit(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


fwd_ms = timeit(lambda: net.predict(x, want=('mask',)))
rows = net.profile(x)
rows = net.profile(x)
print('%-22s %9s %9s %8s' % ('layer', 'ms', 'TFLOP/s', '%peak'))
for name, ms, fl in rows:
    tf = fl / (ms * 1e-3) / 1e12 if ms > 0 else 0
    print('%-22s %9.3f %9.1f %8.1f' % (name, ms, tf, 100 * tf / peak))
conv_ms = sum(r[1] for r in rows if r[2] > 0)
conv_fl = sum(r[2] for r in rows)
vox = d * h * w
print('UNet3D forward %.2f ms per %dx%dx%d stack = %.2f Gvoxel/s; dense layers %.1f TFLOP/s = %.1f%% of %.0f; %d launches'
      % (fwd_ms, d, h, w, vox / fwd_ms / 1e6, conv_fl / (conv_ms * 1e-3) / 1e12,
         100 * conv_fl / (conv_ms * 1e-3) / 1e12 / peak, peak, net.launches()))

# training-step loss on the same stack: weight map of the label slices (W1, per slice as weightmap.py
# does for z-stacks) -> weighted softmax cross-entropy + gradient w.r.t. the logits
out = net.predict(x, want=('logits',))
labels = (torch.rand((d, h, w), device='cuda') > 0.8).to(torch.uint8).contiguous()
wm_ms = timeit(lambda: ops.weightmap_edt(labels, 10., 5., 'float32'))
wmap = ops.weightmap_edt(labels, 10., 5., 'float32')
ce_ms = timeit(lambda: ops.weighted_cross_entropy(out['logits'], labels, wmap, want_grad=True))
loss, grad = ops.weighted_cross_entropy(out['logits'], labels, wmap, want_grad=True)
print('weight map W1 %.3f ms, weighted CE + grad %.3f ms (loss %.6f) for %d voxels'
      % (wm_ms, ce_ms, float(loss), vox))
print(json.dumps({'config': 'UNet3D %dx%dx%d bf16' % (d, h, w), 'forward_ms': fwd_ms,
                  'gvoxel_per_s': vox / fwd_ms / 1e6, 'dense_tflops': conv_fl / (conv_ms * 1e-3) / 1e12,
                  'frac_of_peak': conv_fl / (conv_ms * 1e-3) / 1e12 / peak, 'weightmap_ms': wm_ms,
                  'weighted_ce_ms': ce_ms, 'layers': rows}))
