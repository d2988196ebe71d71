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
