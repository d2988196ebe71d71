import sys, time, numpy as np
sys.path.insert(0, '/root/repo')
from sequitr_b200 import synth
from sequitr_b200.networks import UNet2D
filters = (16, 32, 64, 128, 256)
net = UNet2D({'filters': filters, 'shape': (2048, 2048), 'bridge': 'concat', 'compute': 'bf16'})
net.load_weights(synth.blob_detector_weights(filters, 1, 2, seed=1))
base = synth.frames(4, 2048, 2048, 1, seed=3)
x = np.concatenate([base] * 50)                     # 200 frames, 3.4 GB
raw = np.clip(x[..., 0] * 400 + 3000, 0, 65535).astype(np.uint16)
for arr, kw in ((x, {}), (raw, {'normalise': True})):
    net.segment_and_localise(arr[:8], **kw)
    t = time.perf_counter()
    tables = net.segment_and_localise(arr, frame0=1000, **kw)
    dt = time.perf_counter() - t
    n = [len(tb) for tb in tables]
    assert len(tables) == 200 and all(np.array_equal(tables[i][:, 1:], tables[i % 4][:, 1:]) for i in range(200))
    assert tables[7][0, 0] == 1007.0
    print(arr.dtype, '200 frames in %.3f s -> %.0f frames/s (unpinned host array), objects/frame %s' % (dt, 200 / dt, n[:4]))
