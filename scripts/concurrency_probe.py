"""Experiment: do two UNet forward passes on two streams overlap usefully (HBM-bound level-0/1 layers of one
batch under the tensor-bound level-2..4 layers of the other)?  Serial = one stream, 2 x NB frames back to back;
concurrent = two streams with NB frames each.  SQ_GRID_DIV=2 halves every persistent grid so that CTAs of both
streams fit on an SM together."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sequitr_b200 import synth                      # noqa: E402
from sequitr_b200.networks import UNet2D            # noqa: E402

NB = int(os.environ.get('NB', 4))
filters = (16, 32, 64, 128, 256)
w = synth.blob_detector_weights(filters, 1, 2, seed=1)
nets = []
for i in range(2):
    n = UNet2D({'filters': filters, 'shape': (2048, 2048), 'bridge': 'concat', 'compute': 'bf16'})
    n.load_weights(w)
    nets.append(n)
xs = [torch.randn((NB, 2048, 2048, 1), device='cuda') for _ in range(2)]
streams = [torch.cuda.Stream(), torch.cuda.Stream()]


def run(concurrent, reps=6):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for r in range(reps):
        for i in range(2):
            with torch.cuda.stream(streams[i if concurrent else 0]):
                nets[i].predict(xs[i], want=('mask',))
    torch.cuda.synchronize()
    return 2 * NB * reps / (time.perf_counter() - t0)


for mode in (False, True):
    run(mode, 2)
    print('SQ_GRID_DIV=%s NB=%d %s: %.1f frames/s' % (os.environ.get('SQ_GRID_DIV', '1'), NB,
                                                       'two streams' if mode else 'one stream', run(mode)), flush=True)
