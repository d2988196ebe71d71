"""Drop-in for the hot-path parts of the reference ``sequitr/utils.py``:
``CentroidWriter`` (utils.py:479-578), ``HDF5FileHandler`` (:431-474) and the
size-legality helpers (:230-240).  The per-frame / per-class SciPy loop is
replaced by one call into the CUDA library (``sq_label_centroids_host``).
"""
import logging
import os

import numpy as np

from . import hdf5min, ops

logger = logging.getLogger('worker_process')

try:                                    # h5py is optional in this image
    import h5py
except ImportError:                     # pragma: no cover - depends on the image
    h5py = None


def power_of_two(number):
    """ Return a bool testing whether number is power of two or not (utils.py:230-232) """
    return int(bin(number & (number - 1)), 2) == 0


def divisible_by_two_n_times(x, n):
    """ Check whether a number is divisible by two, n times (utils.py:234-240) """
    for i in range(n):
        x = x / 2.0
    return x % 1 == 0


class HDF5FileHandler(object):
    """ Base class for handling HDF files (utils.py:431-474).

    With h5py installed the file goes through it; otherwise through ``hdf5min``, this package's own
    writer / reader of the HDF5 subset the ``frames/frame_<i>/coords`` layout needs (version-0
    superblock, symbol-table groups, contiguous datasets).  Either way the result is a real HDF5 file.
    """

    def __init__(self, filename=None, read_only=False):
        if not isinstance(filename, str):
            raise TypeError('Filename must be specified as a string')
        pth, f = os.path.split(filename)
        if not os.path.exists(pth):
            raise IOError('Destination path {0:s} doesn\'t exist'.format(pth))
        f_noext, f_ext = os.path.splitext(f)
        if f_ext != '.hdf5':
            # (the reference drops the directory here, utils.py:446-448; we keep it)
            filename = os.path.join(pth, f_noext + '.hdf5')
        self.filename = filename
        self.read_only = read_only
        logger.info('Opening HDF file: {0:s}'.format(filename))
        if h5py is not None:
            self._hdf = h5py.File(filename, 'r+' if read_only else 'w')
        else:
            self._hdf = hdf5min.File(filename, 'r+' if read_only else 'w')

    @property
    def hdf(self):
        return self._hdf

    def __del__(self):
        if getattr(self, '_hdf', None) is not None:
            self.close()

    def close(self):
        """ Manually close the HDF5 file, to prevent HDF5 corruption """
        if self._hdf is None:
            return
        logger.info('Closing HDF file.')
        self._hdf.close()
        self._hdf = None


class CentroidWriter(HDF5FileHandler):
    """ CentroidWriter (utils.py:479-578)

    Using the segmentation output, find the centre of mass of each object and
    write these to the HDF file, grouped by the frame in which they were found.
    Works with both images (N,H,W) and volumes (N,Z,X,Y).
    """

    def __init__(self, filename=None, max_rows=4096):
        HDF5FileHandler.__init__(self, filename)
        self._hdf.create_group('frames')
        self.max_rows = max_rows

    @staticmethod
    def centroids(segmented, max_rows=4096, frame0=0):
        """The arithmetic of ``write`` without the file: list of per-frame (n_i,5)
        float32 tables with rows ``[frame, x, y, z, class]`` (utils.py:559-564)."""
        segmented = np.asarray(segmented)
        if segmented.ndim == 4:
            # volumetric: default input is N,Z,X,Y (utils.py:518-519)
            segmented = np.swapaxes(segmented, 1, -1)
        elif segmented.ndim != 3:
            logger.error("Incorrect image data shape.")
            raise ValueError("Incorrect image data shape.")
        if segmented.dtype != np.uint8:
            if segmented.size and (segmented.min() < 0 or segmented.max() > 255):
                raise ValueError("segmentation classes must fit in uint8")
            segmented = segmented.astype(np.uint8)
        return ops.label_centroids_host(np.ascontiguousarray(segmented), max_rows=max_rows,
                                        frame0=frame0)

    def write(self, segmented):
        """ Take a (large!) numpy array and output dataset """
        im_type = "Volumetric" if np.ndim(segmented) == 4 else "Image"
        n = np.shape(segmented)[0]
        chunk = 64
        for start in range(0, n, chunk):
            if start % 100 < chunk:
                logger.info('Written out {0:d} of {1:d} frames ({2:s})...'.format(start, n, im_type))
            tables = self.centroids(segmented[start:start + chunk], self.max_rows, frame0=start)
            for j, this_frame in enumerate(tables):
                grp = self._hdf['frames'].create_group('frame_' + str(start + j))
                if len(this_frame) == 0:
                    this_frame = []                  # the reference writes shape (0,) for a frame without objects (utils.py:570-578)
                grp.create_dataset('coords', data=this_frame, dtype='float32')


def pinned_array(shape, dtype=np.float32):
    """A page-locked host ndarray (backed by a torch pinned tensor).  Frame batches handed to
    ``UNet.segment_and_localise`` from such an array are DMA-copied at PCIe speed; from ordinary
    (pageable) memory the driver stages every copy through the CPU (~10 GB/s), which caps float32
    frames at ~600 frames/s of 2048^2 -- pin the batch buffer or hand over uint8/uint16 frames."""
    import torch
    tdt = {np.dtype(np.float32): torch.float32, np.dtype(np.uint16): torch.uint16,
           np.dtype(np.uint8): torch.uint8, np.dtype(np.int32): torch.int32}[np.dtype(dtype)]
    t = torch.empty(tuple(int(v) for v in shape), dtype=tdt, pin_memory=True)    # cudaHostAlloc, no pageable twin
    return t.numpy()                        # the array's .base keeps the pinned tensor alive


def pin_in_place(array, retries=3, wait_s=0.5):
    """Page-lock a host ndarray the caller already owns (e.g. a 16.8 GB frame stack) so that
    ``UNet.segment_and_localise`` copies from it by DMA.  The whole array is ONE registration: a host->device copy may
    not span two separately registered ranges (cudaMemcpyAsync answers "invalid argument"), so piecewise registration
    is not an option.  The OS sometimes refuses a multi-GB registration right after another process released one;
    the call is retried ``retries`` times and the CUDA error state is cleared each time (sq_host_register).  Returns
    1.0 when the array is now page-locked, 0.0 when the OS refused: the array simply stays pageable and the host
    calls run at the staged-copy rate."""
    import time
    from . import _lib
    lib = _lib.load()
    if array.nbytes == 0:
        return 1.0
    for attempt in range(retries):
        if lib.sq_host_register(array.ctypes.data, array.nbytes) == _lib.SQ_OK:
            return 1.0
        time.sleep(wait_s)
    return 0.0
