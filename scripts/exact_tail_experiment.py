"""Would higher precision in the last two level-0 layers + head (tcgen05 kind::tf32, or a 3 x bf16 split) make the
benchmarked path's masks identical to the fp32 evaluation?  CPU experiment with the C oracle on the bench's own
weights and frames: the bf16 contract, and the bf16 contract with the last block and the head in FULL fp32 (the
limit of any such variant), both against the fp32 contract.  No GPU needed.

    python scripts/exact_tail_experiment.py [size] [frames]
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import unet_c                      # noqa: E402
from sequitr_b200 import parity, synth         # noqa: E402

size = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 1
filters = (16, 32, 64, 128, 256)
w = synth.blob_detector_weights(filters, 1, 2, seed=1)
x = synth.frames(frames, size, size, 1, seed=1234)
ref = unet_c.unet_forward(x, w, filters, 'concat', contract='fp32')
for name, kw in (('bf16 contract', {}), ('bf16 contract, last block + head in fp32', {'exact_tail': True})):
    out = unet_c.unet_forward(x, w, filters, 'concat', contract='bf16', **kw)
    mp = parity.mask_parity(out['mask'], ref['mask'], ref['logits'])
    err = np.abs(out['logits'] - ref['logits'])
    print('%-44s mask pixels that differ from fp32: %6d of %d (%.2e), largest fp32 margin at a flip %.4f; '
          'logit error max %.4f mean %.5f' % (name, mp['mismatch'], mp['pixels'], mp['mismatch_frac'],
                                               mp['max_margin_of_mismatch'], err.max(), err.mean()))
