"""Small end-to-end pass of every CUDA path for compute-sanitizer (memcheck / racecheck)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sequitr_b200 import synth, ops
from sequitr_b200.networks import UNet2D, UNet3D
filters = (16, 32, 64)
x = synth.frames(2, 64, 72, 1, seed=1, n_objects=3)
w = synth.blob_detector_weights(filters, 1, 2, seed=1)
for compute in ('bf16', 'fp32'):
    net = UNet2D({'filters': filters, 'shape': (64, 72), 'bridge': 'concat', 'compute': compute})
    net.load_weights(w)
    tables, mask = net.segment_and_localise(x, return_mask=True)
    out = net.predict(x)
    print(compute, [len(t) for t in tables], out['probs'].shape)
net = UNet2D({'filters': (16, 32), 'shape': (32, 40), 'bridge': 'eltwise_mul', 'num_inputs': 3, 'num_outputs': 3, 'compute': 'bf16'})
print(net.predict(synth.frames(1, 32, 40, 3, seed=2, n_objects=2))['mask'].shape)
n3 = UNet3D({'filters': (8, 16), 'shape': (16, 16, 8), 'bridge': 'concat', 'compute': 'fp32'})
print(n3.predict(synth.volumes(1, 8, 16, 16, 1))['mask'].shape)
lab = synth.instance_labels(96, 130, 9, seed=3, rmin=4, rmax=9)
print(ops.weightmap_edt_host(lab > 0, 10., 5.).shape, ops.weightmap_edt_host(lab > 0, 10., 5., want_d2=True)[1].max(),
      ops.weightmap_unet_host(lab, 10., 5.).shape, ops.weightmap_unet_host(lab, 10., 40.).shape)
m = (np.random.default_rng(0).random((2, 5, 33, 47)) > 0.6).astype(np.uint8)
t, l = ops.label_centroids_host(m, want_labels=True)
print([len(a) for a in t], l.max())
