"""Sequential vs overlapped (two host threads, net.twin()) walk of a uint16 host stack through the host call."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sequitr_b200 import shard, synth                 # noqa: E402
from sequitr_b200.networks import UNet2D              # noqa: E402

n = int(os.environ.get('N', 512))
per_call = int(os.environ.get('CALL', 64))
filters = (16, 32, 64, 128, 256)
stack = synth.camera_stack(0, n, 2048, 2048, seed=1234, workers=8)   # forked renderers: BEFORE CUDA is initialised
t = torch.from_numpy(stack)
torch.cuda.cudart().cudaHostRegister(t.data_ptr(), t.numel() * 2, 0)
net = UNet2D({'filters': filters, 'shape': (2048, 2048), 'bridge': 'concat', 'compute': 'bf16'})
net.load_weights(synth.blob_detector_weights(filters, 1, 2, seed=1))
for overlap in (False, True, False, True):
    shard.segment_stack(net, stack[:2 * per_call], frames_per_call=per_call, max_rows=2048, overlap=overlap)   # warm-up
    t0 = time.perf_counter()
    tables = shard.segment_stack(net, stack, frames_per_call=per_call, max_rows=2048, overlap=overlap)
    dt = time.perf_counter() - t0
    print('overlap=%-5s %d frames in %.3f s -> %.0f frames/s  digest %s' % (overlap, n, dt, n / dt, shard.tables_digest(tables)[:12]))
