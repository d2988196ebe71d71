"""Level-0 fusion (SQ_FUSE_FIRST=1: down0/conv1 inside down0/conv2's producer) against the default two-launch
path: the logits must be bit-identical (same mma.sync fragments, same epilogue, same bf16 rounding of the
intermediate).  Opt-in: parity-exact but slower on B200."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sequitr_b200 import synth                       # noqa: E402
from sequitr_b200.networks import UNet2D             # noqa: E402

ok = True
for (h, w, n, filters) in ((2048, 2048, 2, (16, 32, 64, 128, 256)), (208, 144, 3, (16, 32, 64)), (64, 64, 1, (16, 32)),
                           (1600, 1200, 1, (16, 32, 64, 128, 256))):
    x = synth.frames(n, h, w, 1, seed=h + w)
    net = UNet2D({'filters': filters, 'shape': (h, w), 'bridge': 'concat', 'compute': 'bf16'})
    net.load_weights(synth.blob_detector_weights(filters, 1, 2, seed=1))
    os.environ['SQ_FUSE_FIRST'] = '0'
    a = net.predict(x, want=('logits', 'mask'))
    la = net.launches()
    os.environ['SQ_FUSE_FIRST'] = '1'
    b = net.predict(x, want=('logits', 'mask'))
    lb = net.launches()
    same = np.array_equal(a['logits'], b['logits']) and np.array_equal(a['mask'], b['mask'])
    ok &= same
    print('%4dx%-4d n=%d  launches %d -> %d  logits identical: %s  (max |diff| %.3g, fg %.3f)'
          % (h, w, n, la, lb, same, float(np.abs(a['logits'] - b['logits']).max()), float(b['mask'].mean())))
print('FUSE CHECK', 'OK' if ok else 'FAILED')
sys.exit(0 if ok else 1)
