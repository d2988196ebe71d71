"""Parity bookkeeping between two runs of the hot path (host side, NumPy only).

The consumer contract of the reference is the uint8 class mask (reference utils.py:492) and
the per-frame centroid rows ``[frame, x, y, z, class]`` it is turned into (utils.py:540-564).
The bf16 tensor-core mode cannot promise bit-identical masks against an fp32 evaluation of the
same network -- a pixel whose two best logits are closer than the accumulated rounding error can
flip -- so every place that reports a speed for that mode also reports, with these helpers,
  * how many mask pixels differ and how decided the fp32 evaluation was at those pixels
    (top-2 logit margin: a flip at a large margin would be a bug, a flip at a near-tie is rounding),
  * how many centroid rows of the consumer's tables change.
``tests/`` asserts the bounds; ``bench.py`` prints the numbers next to the throughput.
"""
import numpy as np


def top2_margin(logits):
    """Difference between the largest and the second largest logit of every pixel (..., K) -> (...)."""
    logits = np.asarray(logits)
    if logits.shape[-1] == 2:
        return np.abs(logits[..., 0] - logits[..., 1])
    part = np.partition(logits, -2, axis=-1)
    return part[..., -1] - part[..., -2]


def mask_parity(mask, ref_mask, ref_logits):
    """Compare a class mask with the reference evaluation of the same network.

    Returns a dict: ``pixels``, ``mismatch`` (count), ``mismatch_frac``, ``max_margin_of_mismatch``
    (largest reference top-2 logit margin at a differing pixel; 0.0 when the masks are identical),
    ``logit_range`` (max - min of the reference logits, the scale the margins are read against)."""
    mask, ref_mask = np.asarray(mask), np.asarray(ref_mask)
    if mask.shape != ref_mask.shape:
        raise ValueError('mask shapes differ: %s vs %s' % (mask.shape, ref_mask.shape))
    differ = mask != ref_mask
    n_diff = int(differ.sum())
    margin = top2_margin(ref_logits)
    return {
        'pixels': int(mask.size),
        'mismatch': n_diff,
        'mismatch_frac': n_diff / float(max(mask.size, 1)),
        'max_margin_of_mismatch': float(margin[differ].max()) if n_diff else 0.0,
        'logit_range': float(np.max(ref_logits) - np.min(ref_logits)),
    }


def centroid_set_diff(tables, ref_tables, tol_px=0.5):
    """Compare two lists of per-frame centroid tables (rows ``[frame, x, y, z, class]`` float32).

    Rows are matched frame by frame and class by class to their nearest neighbour in the other table
    (greedy on sorted distances, one-to-one).  Returns a dict:
      ``rows`` / ``ref_rows``        total row counts,
      ``identical``                  matched rows whose five float32 values are bit-equal,
      ``moved``                      matched within ``tol_px`` but not bit-equal,
      ``max_shift_px``               largest displacement among the matched rows,
      ``unmatched`` / ``ref_unmatched``  rows with no partner within ``tol_px`` (objects that appear,
                                     vanish, merge or split),
      ``rows_changed``               moved + unmatched + ref_unmatched (0 = identical centroid sets)."""
    if len(tables) != len(ref_tables):
        raise ValueError('frame counts differ: %d vs %d' % (len(tables), len(ref_tables)))
    out = {'rows': 0, 'ref_rows': 0, 'identical': 0, 'moved': 0, 'max_shift_px': 0.0,
           'unmatched': 0, 'ref_unmatched': 0}
    for a, b in zip(tables, ref_tables):
        a = np.asarray(a, dtype=np.float32).reshape(-1, 5)
        b = np.asarray(b, dtype=np.float32).reshape(-1, 5)
        out['rows'] += len(a)
        out['ref_rows'] += len(b)
        for c in np.union1d(a[:, 4], b[:, 4]):
            ia, ib = np.flatnonzero(a[:, 4] == c), np.flatnonzero(b[:, 4] == c)
            if len(ia) == 0 or len(ib) == 0:
                out['unmatched'] += len(ia)
                out['ref_unmatched'] += len(ib)
                continue
            pa, pb = a[ia, 1:4].astype(np.float64), b[ib, 1:4].astype(np.float64)
            used_a, used_b = np.zeros(len(ia), bool), np.zeros(len(ib), bool)
            # candidate pairs within tol_px through a KD-tree; greedy one-to-one by distance
            from scipy.spatial import cKDTree
            pairs = cKDTree(pa).sparse_distance_matrix(cKDTree(pb), tol_px, output_type='ndarray')
            for k in np.argsort(pairs['v'], kind='stable'):
                i, j, d = int(pairs['i'][k]), int(pairs['j'][k]), float(pairs['v'][k])
                if used_a[i] or used_b[j]:
                    continue
                used_a[i] = used_b[j] = True
                if np.array_equal(a[ia[i]], b[ib[j]]):
                    out['identical'] += 1
                else:
                    out['moved'] += 1
                out['max_shift_px'] = max(out['max_shift_px'], d)
            out['unmatched'] += int((~used_a).sum())
            out['ref_unmatched'] += int((~used_b).sum())
    out['rows_changed'] = out['moved'] + out['unmatched'] + out['ref_unmatched']
    return out
