"""The other BASELINE configs on one GPU (device-resident inputs, CUDA events): configs[0] (16 frames of
1024^2 x 1ch), configs[2] at batch 8 (the bench workload) and configs[3] (1600x1200 x 3ch, 3 classes)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sequitr_b200 import synth, ops                 # noqa: E402
from sequitr_b200.networks import UNet2D            # noqa: E402

filters = (16, 32, 64, 128, 256)
rows = []
for name, n, shape, cin, k in (('configs[0] 16 x 1024^2 x 1ch, 2 classes', 16, (1024, 1024), 1, 2),
                               ('configs[2] 8 x 2048^2 x 1ch, 2 classes', 8, (2048, 2048), 1, 2),
                               ('configs[3] 8 x 1600x1200 x 3ch, 3 classes', 8, (1600, 1200), 3, 3)):
    net = UNet2D({'filters': filters, 'shape': shape, 'bridge': 'concat', 'num_inputs': cin, 'num_outputs': k,
                  'compute': 'bf16'})
    net.load_weights(synth.blob_detector_weights(filters, cin, k, seed=1))
    x = torch.from_numpy(synth.frames(min(n, 4), shape[0], shape[1], cin, seed=7)).cuda()
    x = x.repeat((n + x.shape[0] - 1) // x.shape[0], 1, 1, 1)[:n].contiguous()

    def step():
        mask = net.predict(x, want=('mask',))['mask']
        return ops.label_centroids(mask, max_rows=2048)

    for _ in range(3):
        table, counts = step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for _ in range(reps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    flops = sum(r[2] for r in net.profile(x))
    rows.append({'config': name, 'ms_per_batch': ms, 'frames_per_s': n / ms * 1e3, 'objects': int(counts.sum()),
                 'unet_tflops_per_s': flops / (ms * 1e-3) / 1e12})
    print('%-46s %8.3f ms/batch  %8.1f frames/s  %6.1f TFLOP/s  %d objects' %
          (name, ms, n / ms * 1e3, flops / (ms * 1e-3) / 1e12, int(counts.sum())))
print(json.dumps(rows))
