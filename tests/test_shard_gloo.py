"""Multi-rank host logic on CPU: world_size-2 gloo run of the frame sharding (no GPU)."""
import os
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n_frames, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from sequitr_b200 import shard
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    lo, hi = shard.frame_range(rank, world, n_frames)
    # stand-in for the per-rank hot path: one "table" per owned frame carrying its global index
    tables = [np.full((1 + f % 3, 5), f, np.float32) for f in range(lo, hi)]
    elapsed = torch.tensor([0.1 * (rank + 1)], dtype=torch.float64)
    dist.barrier()
    dist.all_reduce(elapsed, op=dist.ReduceOp.MAX)          # timing rule: max over ranks
    counts = torch.tensor([hi - lo], dtype=torch.int64)
    dist.all_reduce(counts)                                  # bookkeeping only, not data path
    gathered = [None] * world
    dist.all_gather_object(gathered, tables)
    if rank == 0:
        merged = shard.merge_tables(gathered)
        np.save(os.path.join(out_dir, 'res.npy'),
                np.array([float(elapsed), float(counts), len(merged)] + [t[0, 0] for t in merged]))
    dist.destroy_process_group()


def test_two_rank_sharding(tmp_path):
    n_frames = 11
    mp.spawn(_worker, args=(2, 29613, n_frames, str(tmp_path)), nprocs=2, join=True)
    res = np.load(str(tmp_path / 'res.npy'))
    assert abs(res[0] - 0.2) < 1e-12 and res[1] == n_frames and res[2] == n_frames
    np.testing.assert_array_equal(res[3:], np.arange(n_frames))


def test_frame_ranges_partition():
    from sequitr_b200 import shard
    for n in (0, 1, 7, 2000):
        for g in (1, 2, 4, 8):
            r = [shard.frame_range(i, g, n) for i in range(g)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(g - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


def _grad_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from sequitr_b200 import shard
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    # without a process group the exchange step is a no-op
    g0 = torch.arange(4, dtype=torch.float32)
    assert shard.average_gradients_(g0) == 1 and g0.tolist() == [0., 1., 2., 3.]
    dist.init_process_group('gloo', rank=rank, world_size=world)
    # stand-in for the per-rank gradient arena and loss of a data-parallel training step
    grad = torch.arange(6, dtype=torch.float32) * (rank + 1)
    loss = torch.tensor([1.0 + rank], dtype=torch.float64)
    assert shard.average_gradients_(grad, loss) == world
    if rank == 0:
        np.save(os.path.join(out_dir, 'grad.npy'), np.concatenate([grad.numpy().astype(np.float64), loss.numpy()]))
    dist.destroy_process_group()


def test_two_rank_gradient_average(tmp_path):
    mp.spawn(_grad_worker, args=(2, 29617, str(tmp_path)), nprocs=2, join=True)
    res = np.load(str(tmp_path / 'grad.npy'))
    np.testing.assert_allclose(res[:6], np.arange(6) * 1.5)       # mean of 1x and 2x
    assert res[6] == 1.5
