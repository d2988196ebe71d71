"""Octopus raw camera streams (reference ``dataio/octopus.py:34-306``).

A stream ``<stem><k>.dat`` / ``<stem><k>.dth`` is a run of numbered file pairs: ``.dat`` holds the frames
back to back as uint8 / uint16 (``:236``), ``.dth`` one text line of ``Key: value`` pairs per frame
(``:219-227``; keys include ``H``, ``W`` and optionally ``Bit_Depth``, ``:100-105``).  ``OctopusData``
keeps the reference's constructor, properties and indexing (``stream[n]`` -> float frame, ``:171-176,
243-247``) and adds ``frames_raw`` -- contiguous batches in the STORED integer type, which is what
``UNet.segment_and_localise`` sends across PCIe (1-2 bytes per pixel) and widens on the GPU.
"""
import os
import re
import time

import numpy as np

OCTOPUS_FILE_TYPES = ["uint8", "uint16"]


class OctopusData(object):
    """ OctopusData

    Read contiguous chunks of an octopus stream.

    Args:
        filename: path and stem name of the octopus stream
        contiguous: (bool) only use the first run of consecutively numbered files (default: True)
        header: (bool) only load the headers
        verbose: (bool) display extra information

    Properties:
        bit_depth: bit depth of the images (8 or 16)
        header_keys: names of the header fields
    """

    def __init__(self, filename, contiguous=True, header=False, verbose=False):
        self.data = None
        self.fileopen = -1
        self.framesize = -1
        self.filenum_to_framerange = {}
        self.currentfile = -1
        self.use_contig = contiguous
        self._header_only = header
        self._verbose = verbose
        self.filename = filename
        self.filelist = []
        self.num_frames = 0
        self._header = []
        self._header_keys = []
        # files modified more recently than this many seconds are assumed to be still being written
        # by the instrument and are skipped by refresh() (reference :283-286)
        self.timeout = 60

        self.refresh()
        if not self.filelist:
            raise IOError('No complete Octopus files for stream {0:s} yet.'.format(filename))
        self._open_header(self.filename + str(self.filelist[0]))
        first = self.header(0)
        self.framesize = (int(first['H']), int(first['W']))
        self._bit_depth = int(first['Bit_Depth']) if 'Bit_Depth' in first else 16
        if 'uint' + str(self._bit_depth) not in OCTOPUS_FILE_TYPES:
            raise IOError('Unsupported Octopus bit depth: {0}'.format(self._bit_depth))
        if self._verbose:
            print('Opened Octopus data file, size {0:d}x{1:d}, {2} frames ({3:d}-bit)...'.format(
                self.framesize[1], self.framesize[0], self.num_frames, self.bit_depth))

    @property
    def bit_depth(self):
        return self._bit_depth

    @property
    def header_keys(self):
        return self._header_keys

    def header(self, frame_num):
        return self._return_header(frame_num)

    def _find_file_range(self):
        """ File numbers of the stream (``<stem><k>.dth``); in contiguous mode only the first run of
        consecutive numbers. """
        datadir, stem = os.path.split(self.filename)
        self.filestem = stem
        try:
            files = os.listdir(datadir or '.')
        except (IOError, OSError):
            raise IOError('No files exist in directory: {0:s}'.format(datadir))
        filenums = []
        for name in files:
            m = re.match(re.escape(stem) + r'([0-9]+)\.dth$', name)
            if m:
                filenums.append(int(m.group(1)))
        if not filenums:
            raise IOError('No Octopus stream with pattern {0:s} found.'.format(stem))
        filenums.sort()
        if not self.use_contig:
            return filenums
        run = [filenums[0]]
        for k in filenums[1:]:
            if k != run[-1] + 1:
                break
            run.append(k)
        return run

    def refresh(self):
        """ Pick up files that appeared since the last call (instrument streaming).  Returns True if
        new files were added. """
        to_update = []
        for k in self._find_file_range():
            last_modified = os.stat(self.filename + str(k) + '.dth').st_mtime
            if k not in self.filelist and (time.time() - last_modified) > self.timeout:
                to_update.append(k)
        if not to_update:
            return False
        for k in to_update:
            self.filelist.append(k)
            self._open_header(self.filename + str(k))
            n = len(self._header)
            self.filenum_to_framerange[k] = (self.num_frames, self.num_frames + n - 1)
            self.num_frames += n
        self.currentfile = -1
        if self._verbose:
            print('Updated Octopus stream with {0:d} new files...'.format(len(to_update)))
        return True

    def _open_header(self, filename):
        try:
            with open(filename + '.dth') as fh:
                lines = [ln for ln in fh.readlines() if ln.strip()]
        except (IOError, OSError):
            raise IOError(filename + ' is not a valid file')
        self._header = [re.findall(r'\S+:\s*(\S+)', line) for line in lines]
        self._header_keys = re.findall(r'(\w+)\s*:\s*', lines[0]) if lines else []

    def _open_file(self, filename, num_frames):
        try:
            self.data = np.memmap(filename + '.dat', dtype='uint' + str(self.bit_depth), mode='r',
                                  shape=(num_frames, self.framesize[0], self.framesize[1]))
        except (IOError, OSError, ValueError):
            self.data = []
            raise IOError(filename + ' is not a valid file. Make sure the path to the images still exists!')
        self.fileopen = True

    def _select_file(self, abs_frame_num):
        """ Make the file holding ``abs_frame_num`` current; returns the frame's index inside it. """
        if abs_frame_num < 0 or abs_frame_num >= self.num_frames:
            raise IndexError('frame {0} outside the stream (0..{1})'.format(abs_frame_num, self.num_frames - 1))
        j = self.currentfile
        if j == -1 or not (self.filenum_to_framerange[j][0] <= abs_frame_num <= self.filenum_to_framerange[j][1]):
            for i in self.filelist:
                lo, hi = self.filenum_to_framerange[i]
                if lo <= abs_frame_num <= hi:
                    self.currentfile = i
                    self._open_header(self.filename + str(i))
                    if not self._header_only:
                        self._open_file(self.filename + str(i), len(self._header))
                    break
        return abs_frame_num - self.filenum_to_framerange[self.currentfile][0]

    def __getitem__(self, abs_frame_num):
        """ Frame ``abs_frame_num`` of the stream as a float array (reference :171-176, 243-247). """
        local = self._select_file(int(abs_frame_num))
        if self._header_only:
            return np.array(())
        return np.array(self.data[local, :, :], dtype='float')

    def info(self, abs_frame_num):
        """ Header dictionary of a frame with its absolute number under 'N' (reference :189-191). """
        local = self._select_file(int(abs_frame_num))
        d = self._return_header(local)
        d['N'] = abs_frame_num
        return d

    def _return_header(self, frame_num):
        return dict((self._header_keys[i], self._header[frame_num][i]) for i in range(len(self._header_keys)))

    def __len__(self):
        return self.num_frames

    def frames_raw(self, start, count, out=None):
        """ ``count`` frames from ``start`` as ONE contiguous (count,H,W) array in the stored integer
        type (crossing file boundaries) -- the batch ``UNet.segment_and_localise`` takes.  ``out``: an
        optional destination, e.g. a page-locked buffer from ``utils.pinned_array`` reused per batch. """
        count = min(int(count), self.num_frames - int(start))
        shape = (max(count, 0), self.framesize[0], self.framesize[1])
        if out is None:
            out = np.empty(shape, dtype='uint' + str(self.bit_depth))
        else:
            out = out[:shape[0]]
            if out.shape != shape or out.dtype != np.dtype('uint' + str(self.bit_depth)):
                raise ValueError('frames_raw: out must be a (>=count,H,W) array of the stream dtype')
        done = 0
        while done < count:
            local = self._select_file(start + done)
            lo, hi = self.filenum_to_framerange[self.currentfile]
            take = min(count - done, hi - (start + done) + 1)
            out[done:done + take] = self.data[local:local + take]
            done += take
        return out

    def to_array(self):
        """ The entire stack as one uint8 array (reference :305-310 narrows to uint8 as well). """
        image_data = np.zeros((len(self), self.framesize[0], self.framesize[1]), dtype='uint8')
        for i in range(len(self)):
            image_data[i, ...] = self[i]
        return image_data


def write_octopus_stream(stem, frames, frames_per_file=None, first_file=0, extra=None, age=120.0):
    """ Write ``frames`` (N,H,W) uint8 / uint16 as an octopus stream ``<stem><k>.dat/.dth`` (synthetic
    stacks for tests and benchmarks).  ``age`` back-dates the files so that ``refresh()`` accepts them. """
    frames = np.ascontiguousarray(frames)
    if frames.dtype not in (np.uint8, np.uint16) or frames.ndim != 3:
        raise ValueError('write_octopus_stream: (N,H,W) uint8 / uint16 frames')
    n, h, w = frames.shape
    per = int(frames_per_file or n)
    k = first_file
    for f0 in range(0, n, per):
        chunk = frames[f0:f0 + per]
        chunk.tofile(stem + str(k) + '.dat')
        with open(stem + str(k) + '.dth', 'w') as fh:
            for i in range(chunk.shape[0]):
                fields = [('N', f0 + i), ('H', h), ('W', w), ('Bit_Depth', frames.dtype.itemsize * 8)]
                fields += [(key, val[f0 + i]) for key, val in (extra or {}).items()]
                fh.write(' '.join('%s: %s' % kv for kv in fields) + '\n')
        if age:
            t = time.time() - age
            os.utime(stem + str(k) + '.dth', (t, t))
            os.utime(stem + str(k) + '.dat', (t, t))
        k += 1
    return k - first_file
