"""Seeded synthetic microscopy-like inputs and UNet weights (host side, NumPy).

The reference ships no data, checkpoints or benchmarks, so
tests and ``bench.py`` use these generators: frames are N(0,1) camera noise
plus soft-edged discs ("cells", radius 8-20 px, amplitude 2-4 sigma) that drift
slowly with the frame index; label images are the same discs with >= 2 px
background gaps so the binary and the instance views agree.

Weights use the TF variable-scope names of the reference graph
(``UNet/down{i}/conv{1,2}``, ``UNet/up{i}/upscale``, ``UNet/up{i}/conv{1,2}``,
``UNet/to_image`` -- scopes from reference networks/unet.py:234,252,268-271,
294,312-315) and TF kernel layouts (HWIO for convolutions, (kh,kw,out,in) for
``conv2d_transpose``).
"""
import numpy as np

DEFAULT_FILTERS = (16, 32, 64, 128, 256)


# --------------------------------------------------------------------- scenes

def disc_scene(h, w, n_objects, seed=0, rmin=8.0, rmax=20.0, gap=2.0, max_tries=40):
    """Place up to n_objects non-overlapping discs. Returns (cy, cx, r) float64 arrays."""
    rng = np.random.default_rng(seed)
    cys, cxs, rs = [], [], []
    rmax = max(1.0, min(rmax, (min(h, w) - 6) / 2.0 - 1.0))
    rmin = min(rmin, rmax)
    cell = int(2 * rmax + gap) + 1
    gh, gw = h // cell + 2, w // cell + 2
    grid = [[[] for _ in range(gw)] for _ in range(gh)]
    for _ in range(n_objects):
        for _t in range(max_tries):
            r = rng.uniform(rmin, rmax)
            cy = rng.uniform(r + 1, h - r - 2)
            cx = rng.uniform(r + 1, w - r - 2)
            gy, gx = int(cy) // cell, int(cx) // cell
            ok = True
            for yy in range(max(gy - 1, 0), min(gy + 2, gh)):
                for xx in range(max(gx - 1, 0), min(gx + 2, gw)):
                    for j in grid[yy][xx]:
                        if (cys[j] - cy) ** 2 + (cxs[j] - cx) ** 2 < (rs[j] + r + gap + 1.5) ** 2:
                            ok = False
                            break
                    if not ok:
                        break
                if not ok:
                    break
            if ok:
                grid[gy][gx].append(len(rs))
                cys.append(cy)
                cxs.append(cx)
                rs.append(r)
                break
    return np.array(cys), np.array(cxs), np.array(rs)


def instance_labels(h, w, n_objects, seed=0, **kw):
    """int32 (H,W) instance-label image, 0 = background, instances 1..n."""
    cy, cx, r = disc_scene(h, w, n_objects, seed, **kw)
    lab = np.zeros((h, w), dtype=np.int32)
    for i in range(len(r)):
        y0, y1 = max(int(cy[i] - r[i]) - 1, 0), min(int(cy[i] + r[i]) + 2, h)
        x0, x1 = max(int(cx[i] - r[i]) - 1, 0), min(int(cx[i] + r[i]) + 2, w)
        yy, xx = np.mgrid[y0:y1, x0:x1]
        inside = (yy - cy[i]) ** 2 + (xx - cx[i]) ** 2 <= r[i] ** 2
        lab[y0:y1, x0:x1][inside] = i + 1
    return lab


def class_mask(h, w, n_objects, n_classes=2, seed=0, **kw):
    """uint8 (H,W) segmentation-like mask: class of instance i is 1 + i % (n_classes-1)."""
    lab = instance_labels(h, w, n_objects, seed, **kw)
    k = max(n_classes - 1, 1)
    return np.where(lab > 0, 1 + (lab - 1) % k, 0).astype(np.uint8)


def default_object_count(h, w):
    """~600 objects at 2048^2, scaled by area."""
    return max(1, int(round(600.0 * h * w / (2048.0 * 2048.0))))


def frames(n, h, w, cin=1, seed=1234, n_objects=None, first_frame=0):
    """float32 (N,H,W,Cin) synthetic time-lapse; frame i uses seed + first_frame + i
    for the noise and a slow drift of the shared disc scene."""
    if n_objects is None:
        n_objects = default_object_count(h, w)
    cy, cx, r = disc_scene(h, w, n_objects, seed)
    rng0 = np.random.default_rng(seed + 7)
    vy = rng0.uniform(-0.5, 0.5, size=len(r))
    vx = rng0.uniform(-0.5, 0.5, size=len(r))
    amp = rng0.uniform(2.0, 4.0, size=(len(r), cin))
    out = np.empty((n, h, w, cin), dtype=np.float32)
    for f in range(n):
        t = first_frame + f
        rng = np.random.default_rng(seed + 1000003 * (t + 1))
        img = rng.standard_normal((h, w, cin), dtype=np.float32)
        for i in range(len(r)):
            yc, xc = cy[i] + vy[i] * t, cx[i] + vx[i] * t
            y0, y1 = max(int(yc - r[i]) - 2, 0), min(int(yc + r[i]) + 3, h)
            x0, x1 = max(int(xc - r[i]) - 2, 0), min(int(xc + r[i]) + 3, w)
            if y0 >= y1 or x0 >= x1:
                continue
            yy, xx = np.mgrid[y0:y1, x0:x1]
            d = np.sqrt((yy - yc) ** 2 + (xx - xc) ** 2)
            soft = np.clip(r[i] - d + 0.5, 0.0, 1.0).astype(np.float32)
            img[y0:y1, x0:x1, :] += soft[..., None] * amp[i][None, None, :].astype(np.float32)
        out[f] = img
    return out


def to_camera_counts(frames_f32, gain=400.0, offset=3000.0):
    """float32 synthetic frames -> uint16 camera counts (what dataio.OctopusData.frames_raw delivers)."""
    return np.clip(np.asarray(frames_f32, dtype=np.float32) * np.float32(gain) + np.float32(offset),
                   0, 65535).astype(np.uint16)


def _camera_chunk(args):
    buf, lo, t0, n, h, w, seed = args
    out = np.frombuffer(buf, dtype=np.uint16).reshape(-1, h, w)
    out[t0 - lo:t0 - lo + n] = to_camera_counts(frames(n, h, w, 1, seed=seed, first_frame=t0)[..., 0])
    return n


def camera_stack(lo, hi, h, w, seed=1234, workers=None, chunk=4):
    """uint16 (hi-lo, H, W) time-lapse of GLOBAL frames [lo, hi): frame t depends on (seed, t) only, so
    every rank of a sharded run generates exactly its own slice of the same stack.  Rendered by a pool
    of forked worker processes into one anonymous shared mapping (call this BEFORE initialising CUDA:
    the children must not inherit a CUDA context).  Returns an ndarray backed by that mapping."""
    import mmap
    import os
    n = hi - lo
    nbytes = max(n * h * w * 2, mmap.PAGESIZE)
    buf = mmap.mmap(-1, nbytes)                       # MAP_SHARED | MAP_ANONYMOUS: visible to the children
    workers = max(1, min(workers or (os.cpu_count() or 1), (n + chunk - 1) // chunk or 1))
    jobs = [(buf, lo, t0, min(chunk, hi - t0), h, w, seed) for t0 in range(lo, hi, chunk)]
    if workers == 1 or len(jobs) <= 1:
        for j in jobs:
            _camera_chunk(j)
    else:
        procs = []
        for k in range(workers):                      # static round-robin split, no pickling of the buffer
            pid = os.fork()
            if pid == 0:
                code = 0
                try:
                    for j in jobs[k::workers]:
                        _camera_chunk(j)
                except BaseException:
                    code = 1
                os._exit(code)
            procs.append(pid)
        for pid in procs:
            _, status = os.waitpid(pid, 0)
            if status != 0:
                raise RuntimeError('camera_stack: a worker process failed (status %d)' % status)
    return np.frombuffer(buf, dtype=np.uint16)[:n * h * w].reshape(n, h, w)


def volumes(n, d, h, w, cin=1, seed=4321):
    """float32 (N,D,H,W,Cin) synthetic z-stacks: noise + a few soft balls."""
    rng = np.random.default_rng(seed)
    out = rng.standard_normal((n, d, h, w, cin), dtype=np.float32)
    nballs = max(1, (d * h * w) // 60000)
    zz, yy, xx = np.mgrid[0:d, 0:h, 0:w]
    for v in range(n):
        for _ in range(nballs):
            r = rng.uniform(min(3.0, d / 4.0), max(3.0, min(8.0, d / 2.0)))
            c = rng.uniform([0, 0, 0], [d, h, w])
            dist = np.sqrt((zz - c[0]) ** 2 + (yy - c[1]) ** 2 + (xx - c[2]) ** 2)
            out[v] += (np.clip(r - dist + 0.5, 0, 1) * 3.0).astype(np.float32)[..., None]
    return out


# -------------------------------------------------------------------- weights

def unet_layer_names(filters=DEFAULT_FILTERS):
    """Conv scopes in execution order (reference networks/unet.py:238-253)."""
    names = []
    for i in range(len(filters)):
        names += ['UNet/down%d/conv1' % i, 'UNet/down%d/conv2' % i]
    for i in reversed(range(len(filters) - 1)):
        names += ['UNet/up%d/upscale' % i, 'UNet/up%d/conv1' % i, 'UNet/up%d/conv2' % i]
    names.append('UNet/to_image')
    return names


def unet_weights(filters=DEFAULT_FILTERS, num_inputs=1, num_outputs=2, ndim=2,
                 bridge='concat', kernel=3, seed=42, scale=2.0, bias_std=0.05,
                 affine=False):
    """Random variance-scaling (fan-in, normal) weights, TF names and layouts.

    Returns dict name -> float32 array:
      '<scope>/kernel'  conv: (k,k[,k],Cin,Cout); upscale: (2,2[,2],Cout,Cin);
                        to_image: (1,1[,1],Cin,K)
      '<scope>/bias'    (Cout,)
      '<scope>/scale', '<scope>/shift'  optional folded-BN affine (affine=True)
    """
    rng = np.random.default_rng(seed)
    w = {}
    ks = (kernel,) * ndim

    def conv(name, cin, cout, k):
        fan_in = cin * int(np.prod(k))
        w[name + '/kernel'] = (rng.standard_normal(k + (cin, cout)) *
                               np.sqrt(scale / fan_in)).astype(np.float32)
        w[name + '/bias'] = (rng.standard_normal(cout) * bias_std).astype(np.float32)
        if affine and not name.endswith('to_image'):
            w[name + '/scale'] = rng.uniform(0.8, 1.2, cout).astype(np.float32)
            w[name + '/shift'] = (rng.standard_normal(cout) * bias_std).astype(np.float32)

    cin = num_inputs
    for i, f in enumerate(filters):
        conv('UNet/down%d/conv1' % i, cin, f, ks)
        conv('UNet/down%d/conv2' % i, f, f, ks)
        cin = f
    for i in reversed(range(len(filters) - 1)):
        f = filters[i]
        fan_in = cin * 2 ** ndim / 2 ** ndim   # each output pixel sees Cin inputs
        w['UNet/up%d/upscale/kernel' % i] = (rng.standard_normal((2,) * ndim + (f, cin)) *
                                             np.sqrt(1.0 / fan_in)).astype(np.float32)
        w['UNet/up%d/upscale/bias' % i] = (rng.standard_normal(f) * bias_std).astype(np.float32)
        merged = 2 * f if bridge == 'concat' else f
        conv('UNet/up%d/conv1' % i, merged, f, ks)
        conv('UNet/up%d/conv2' % i, f, f, ks)
        cin = f
    conv('UNet/to_image', cin, num_outputs, (1,) * ndim)
    return w


def blob_detector_weights(filters=DEFAULT_FILTERS, num_inputs=1, num_outputs=2,
                          seed=42, threshold=1.2, gain=8.0):
    """Random weights with one hand-made pathway so that a 2-D concat-bridge UNet
    segments the synthetic discs: channel 0 of down0/conv1, down0/conv2,
    (skip ->) up0/conv1, up0/conv2 is a chain of 3x3 box blurs of the mean input
    and the head thresholds it (class k>0 logit = gain*(blur - threshold)).
    All other channels keep their random weights, so every layer still does
    its full dense work."""
    w = unet_weights(filters, num_inputs, num_outputs, ndim=2, bridge='concat', seed=seed)
    f0 = filters[0]

    def box(name, src_channels, cin):
        k = w[name + '/kernel']
        k[:, :, :, 0] = 0.0
        for c in src_channels:
            k[:, :, c, 0] = 1.0 / (9.0 * len(src_channels))
        w[name + '/bias'][0] = 0.0

    box('UNet/down0/conv1', list(range(num_inputs)), num_inputs)
    box('UNet/down0/conv2', [0], f0)
    box('UNet/up0/conv1', [f0], 2 * f0)          # concat = [upsampled(f0), skip(f0)]
    box('UNet/up0/conv2', [0], f0)
    k = w['UNet/to_image/kernel']
    k[...] = 0.0
    b = w['UNet/to_image/bias']
    b[...] = 0.0
    for c in range(1, num_outputs):
        k[0, 0, 0, c] = gain
        b[c] = -gain * threshold * (1.0 + 0.5 * (c - 1))
    return w
