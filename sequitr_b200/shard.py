"""Frame sharding for the multi-GPU runs: frames (and whole z-stacks) are independent
units, so each rank owns one contiguous range and there is NO collective on the data path
(the reference's analogue is one worker process per GPU job,
reference core.py:41-42).  Per-rank centroid tables already carry global frame indices."""


def frame_range(rank, world_size, n_frames):
    """Contiguous range [lo, hi) of rank `rank` out of `world_size` for `n_frames` frames."""
    if not (0 <= rank < world_size):
        raise ValueError('rank %d outside world of %d' % (rank, world_size))
    return (rank * n_frames) // world_size, ((rank + 1) * n_frames) // world_size


def merge_tables(per_rank_tables):
    """Host-side concatenation of the ranks' per-frame tables in frame order."""
    out = []
    for tables in per_rank_tables:
        out.extend(tables)
    return out


def tables_digest(tables):
    """sha256 over the per-frame centroid tables in frame order (row count + float32 bytes of every
    frame): two runs of the same stack -- whatever the number of ranks, the shard boundaries or the
    chunking -- must produce the same digest."""
    import hashlib
    import numpy as np
    h = hashlib.sha256()
    for t in tables:
        t = np.ascontiguousarray(t, dtype=np.float32)
        h.update(np.int64(len(t)).tobytes())
        h.update(t.tobytes())
    return h.hexdigest()


def segment_stack(net, frames, frame0=0, frames_per_call=250, max_rows=4096, normalise=True, overlap=False):
    """Run this rank's contiguous frame range ``frames`` (host array (n,H,W[,C]); camera-native
    uint8 / uint16 or float32) through ``net.segment_and_localise`` call by call and return the list of
    per-frame tables, rows carrying GLOBAL frame indices (``frame0`` = global index of ``frames[0]``):
    the loop a Sequitr job runs over a time-lapse (reference utils.py:531-578), one rank's share of it.

    ``overlap``: two host threads alternate the calls, the second through ``net.twin()`` (same weights, its own
    library handle and streams): the one-frame copy ramp at the start of a call and the label-and-localise /
    read-back tail at its end then hide under the other call's convolutions.  Same tables, same order."""
    n = len(frames)
    starts = list(range(0, n, frames_per_call))

    def run(worker, s):
        return worker.segment_and_localise(frames[s:s + frames_per_call], frame0=frame0 + s, max_rows=max_rows,
                                           normalise=normalise)

    if not overlap or len(starts) < 2:
        out = []
        for s in starts:
            out.extend(run(net, s))
        return out
    from concurrent.futures import ThreadPoolExecutor
    twin = getattr(net, '_overlap_twin', None)
    if twin is None:
        twin = net._overlap_twin = net.twin()
    results = [None] * len(starts)

    def lane(worker, idx):
        for i in idx:
            results[i] = run(worker, starts[i])

    with ThreadPoolExecutor(2) as pool:
        futs = [pool.submit(lane, net, range(0, len(starts), 2)), pool.submit(lane, twin, range(1, len(starts), 2))]
        for f in futs:
            f.result()
    out = []
    for r in results:
        out.extend(r)
    return out


def average_gradients_(grad, loss=None, group=None):
    """The one exchange step of data-parallel training: every rank computed the gradient of the mean
    loss over ITS share of the batch into ``grad`` (the trainer's contiguous gradient arena, a float32 tensor); one
    all-reduce sums them over the ranks and the result is divided by the world size in place -- the gradient of the
    mean loss over the whole batch when the shares are equally large.  ``loss`` (a float64 tensor of one element) is
    averaged the same way.  NCCL over NVLink / NVSwitch for cuda tensors, gloo for the CPU tests.  Returns the world
    size (1 without an initialised process group: nothing to do)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return 1
    world = dist.get_world_size(group)
    if world == 1:
        return 1
    dist.all_reduce(grad, op=dist.ReduceOp.SUM, group=group)
    grad.div_(world)
    if loss is not None:
        dist.all_reduce(loss, op=dist.ReduceOp.SUM, group=group)
        loss.div_(world)
    return world


def bind_to_gpu_numa(device_index):
    """Pin the calling worker process to the CPUs NVML reports as closest to GPU ``device_index`` (one
    worker per GPU, the reference's model: core.py:41-42), so that the pinned frame buffers it allocates
    afterwards are first-touched on that GPU's NUMA node and eight workers do not share one socket's
    memory bandwidth.  Returns the CPU set, or None when NVML / affinity control is unavailable."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(device_index))
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * i + b for i, w in enumerate(mask) for b in range(64) if (int(w) >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:
        return None
