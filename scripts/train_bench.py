"""Time the training step (BASELINE config 5) on the device: UNet2D with the default filters on 512x512 crops (what
tr_augment yields by default, reference networks/unet.py:366) and a small UNet3D.  Prints one JSON line per case.
    python scripts/train_bench.py [steps]
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import sequitr_b200
    from sequitr_b200 import synth
    from sequitr_b200.networks import UNet2D, UNet3D
    from sequitr_b200.networks.unet import ModeKeys
    sequitr_b200.require_gpu()
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    cases = [
        ('UNet2D 16-256 concat, batch 4 of 512x512x1', UNet2D, (16, 32, 64, 128, 256), (4, 512, 512), 2),
        ('UNet2D 16-256 concat, batch 1 of 1024x1024x1', UNet2D, (16, 32, 64, 128, 256), (1, 1024, 1024), 2),
        ('UNet3D 16-64 concat, batch 1 of 32x256x256x1', UNet3D, (16, 32, 64), (1, 32, 256, 256), 3),
    ]
    if os.environ.get('BIG'):
        # BASELINE configs[4] at full size: one 64x1024x1024 z-stack (fp32 activations + gradients: ~75 GB)
        cases.append(('UNet3D 16-64 concat, batch 1 of 64x1024x1024x1', UNet3D, (16, 32, 64), (1, 64, 1024, 1024), 3))
    only = os.environ.get('CASE')
    for i, (name, cls, filters, sp, ndim) in enumerate(cases):
        if only is not None and int(only) != i:
            continue
        rng = np.random.default_rng(i)
        net = cls({'filters': filters, 'shape': sp[1:], 'bridge': 'concat', 'num_inputs': 1, 'num_outputs': 2,
                   'compute': 'fp32', 'dropout': 0.4}, mode=ModeKeys.TRAIN)
        net.load_weights(synth.unet_weights(filters, 1, 2, ndim=ndim, bridge='concat', seed=1))
        tr = net.trainer(learning_rate=1e-3, seed=1)
        image = torch.from_numpy(rng.standard_normal(sp + (1,)).astype(np.float32)).cuda()
        labels = torch.from_numpy(rng.integers(0, 2, sp).astype(np.uint8)).cuda()
        wmap = torch.from_numpy((1 + 9 * rng.random(sp)).astype(np.float32)).cuda()
        l0 = tr.step(image, labels, wmap)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            loss = tr.step(image, labels, wmap)          # .item() of the loss synchronises every step
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) / steps * 1e3
        t0 = time.perf_counter()
        for _ in range(steps):
            net.predict(image, want=('mask',))
        torch.cuda.synchronize()
        fwd_ms = (time.perf_counter() - t0) / steps * 1e3
        px = int(np.prod(sp))
        print(json.dumps({'case': name, 'ms_per_step': round(ms, 2), 'fp32_inference_ms': round(fwd_ms, 2),
                          'pixels_per_s': round(px / ms * 1e3), 'first_loss': round(l0, 4), 'last_loss': round(loss, 4),
                          'arithmetic': 'fp32 CUDA cores'}))
        tr.close()


if __name__ == '__main__':
    main()
