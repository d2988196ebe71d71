"""CPU oracle for the label-and-localise post-process (TEST INFRASTRUCTURE ONLY).

Restates ``CentroidWriter.write`` (``/root/reference/sequitr/utils.py:505-578``)
around the *same* SciPy calls (``scipy.ndimage.label`` utils.py:547,
``center_of_mass`` utils.py:550), returning arrays instead of writing HDF5
(h5py is not installed here).  ``utils.py`` itself is Python-2 only and
imports TensorFlow at module scope, so it cannot be imported; the reference
ships no fixtures for this path ("parity unpinned" by reference tests; pinned
to SciPy, the reference's own dependency).
"""
import numpy as np
from scipy.ndimage import label, center_of_mass


def frame_centroids(out, frame_index):
    """Body of the per-frame loop, utils.py:539-566.  ``out`` is one frame
    (H,W) or one swapped volume (Y,X,Z).  Returns (n,5) float32 (n may be 0)
    and the list of per-class label matrices (for label-matrix parity)."""
    classes = [x for x in np.unique(out) if x > 0]
    this_frame = []
    matrices = []
    for c in classes:
        matrix, n_labels = label(out == c)
        labels = [l for l in np.unique(matrix) if l > 0]
        coords = center_of_mass(out, matrix, labels)
        matrices.append((int(c), matrix))
        if len(coords) < 1:
            continue
        if out.ndim == 3:
            x, y, z = zip(*coords)
        else:
            x, y = zip(*coords)
            z = [0.0] * len(x)
        this_class = np.zeros((len(x), 5), dtype='float32')
        this_class[:, 0] = frame_index
        this_class[:, 1] = x
        this_class[:, 2] = y
        this_class[:, 3] = z
        this_class[:, 4] = c
        this_frame.append(this_class)
    if this_frame:
        table = np.concatenate(this_frame, axis=0)
    else:
        table = np.zeros((0, 5), dtype='float32')
    return table, matrices


def centroid_tables(segmented):
    """utils.py:505-578 on the whole stack.  ``segmented``: (N,H,W) or
    (N,Z,X,Y) (swapped to (N,Y,X,Z) like utils.py:519).  Returns a list of
    per-frame (n_i,5) float32 tables == the ``frames/frame_<i>/coords`` datasets."""
    segmented = np.asarray(segmented)
    if segmented.ndim == 4:
        segmented = np.swapaxes(segmented, 1, -1)
    elif segmented.ndim != 3:
        raise ValueError("Incorrect image data shape.")
    return [frame_centroids(segmented[i, ...], i)[0] for i in range(segmented.shape[0])]


def label_matrix(out):
    """Combined label matrix for one frame: pixel -> SciPy label number within
    its class (0 on background); used to check the GPU label matrix."""
    res = np.zeros(out.shape, dtype=np.int32)
    for c, matrix in frame_centroids(out, 0)[1]:
        res[out == c] = matrix[out == c]
    return res
