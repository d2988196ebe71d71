"""Micro-Manager 2.0 position folders (reference ``dataio/micromanager.py:30-262``).

A position folder holds ``metadata.txt`` (JSON: ``Summary``, ``Coords-<file>``, ``Metadata-<file>``
entries, ``:56-88``) and one TIFF per (channel, frame, slice).  ``MicromanagerMetadataParser`` and
``MicromanagerReader`` keep the reference's constructors, properties and indexing; ``frames_raw``
adds contiguous batches in the STORED integer type for ``UNet.segment_and_localise``.  The reference
reads TIFFs with ``tifffile`` (absent here); camera frames are baseline TIFFs (uncompressed strips,
8/16-bit), which ``read_tiff`` parses directly.
"""
import json
import os
import struct
from datetime import datetime

import numpy as np

_TIFF_TYPES = {1: 'B', 2: 'c', 3: 'H', 4: 'I', 5: 'II', 16: 'Q'}
_SAMPLE_DTYPES = {(1, 8): 'u1', (1, 16): 'u2', (1, 32): 'u4', (2, 8): 'i1', (2, 16): 'i2', (2, 32): 'i4',
                  (3, 32): 'f4', (3, 64): 'f8'}


def read_tiff(filename):
    """First image of a baseline TIFF (uncompressed, chunky, 1 sample per pixel) as a (rows, cols) array."""
    with open(filename, 'rb') as fh:
        buf = fh.read()
    if buf[:2] == b'II':
        bo = '<'
    elif buf[:2] == b'MM':
        bo = '>'
    else:
        raise IOError('%s is not a TIFF file' % filename)
    magic, ifd = struct.unpack(bo + 'HI', buf[2:8])
    if magic != 42:
        raise IOError('%s: BigTIFF / unknown TIFF flavour (magic %d)' % (filename, magic))
    (nent,) = struct.unpack(bo + 'H', buf[ifd:ifd + 2])
    tags = {}
    for i in range(nent):
        e = ifd + 2 + 12 * i
        tag, typ, cnt = struct.unpack(bo + 'HHI', buf[e:e + 8])
        code = _TIFF_TYPES.get(typ)
        if code is None or code == 'II':
            continue
        size = struct.calcsize(code) * cnt
        off = e + 8 if size <= 4 else struct.unpack(bo + 'I', buf[e + 8:e + 12])[0]
        tags[tag] = struct.unpack(bo + str(cnt) + code, buf[off:off + size]) if code != 'c' else buf[off:off + size]
    width, height = tags[256][0], tags[257][0]
    bits = tags.get(258, (1,))[0]
    if tags.get(259, (1,))[0] != 1:
        raise IOError('%s: compressed TIFFs are not supported' % filename)
    if tags.get(277, (1,))[0] != 1:
        raise IOError('%s: only single-sample (greyscale) TIFFs are supported' % filename)
    fmt = tags.get(339, (1,))[0]
    if (fmt, bits) not in _SAMPLE_DTYPES:
        raise IOError('%s: unsupported sample format %d / %d bits' % (filename, fmt, bits))
    dt = np.dtype(bo + _SAMPLE_DTYPES[(fmt, bits)])
    offsets, counts = tags[273], tags.get(279)
    if counts is None:
        counts = (width * height * dt.itemsize,)
    data = b''.join(buf[o:o + c] for o, c in zip(offsets, counts))
    img = np.frombuffer(data, dtype=dt, count=width * height).reshape(height, width)
    return img.astype(dt.newbyteorder('='))


def write_tiff(filename, image):
    """Baseline little-endian TIFF, one strip, uint8 / uint16 / float32 (what ``weightmap.py:205`` and
    the camera write)."""
    image = np.ascontiguousarray(image)
    fmt = {'u': 1, 'i': 2, 'f': 3}[image.dtype.kind]
    if image.ndim != 2 or (fmt, image.dtype.itemsize * 8) not in _SAMPLE_DTYPES:
        raise ValueError('write_tiff: 2-D integer / float32 / float64 images')
    h, w = image.shape
    data = image.astype(image.dtype.newbyteorder('<')).tobytes()
    entries = [(256, 4, w), (257, 4, h), (258, 3, image.dtype.itemsize * 8), (259, 3, 1), (262, 3, 1),
               (273, 4, 8), (277, 3, 1), (278, 4, h), (279, 4, len(data)), (339, 3, fmt)]
    ifd_off = 8 + len(data) + (len(data) & 1)
    with open(filename, 'wb') as fh:
        fh.write(b'II' + struct.pack('<HI', 42, ifd_off))
        fh.write(data + (b'\0' if len(data) & 1 else b''))
        fh.write(struct.pack('<H', len(entries)))
        for tag, typ, val in entries:
            fh.write(struct.pack('<HHI', tag, typ, 1) + (struct.pack('<HH', val, 0) if typ == 3 else struct.pack('<I', val)))
        fh.write(struct.pack('<I', 0))


class MicromanagerMetadataParser(object):
    """ Parse the micromanager metadata for a particular position and channel.

        Image filenames: img_channel000_position001_time000000002_z000.tif

    Args:
        filepath:   the folder containing the image data and metadata
        channel:    give a particular channel to return the data
    """

    def __init__(self, filepath, channel=None):
        root, position = os.path.split(filepath.rstrip(os.sep))
        with open(os.path.join(filepath, 'metadata.txt'), 'r') as metadata_file:
            self.raw = json.load(metadata_file)
        if channel is None:
            self.is_channel = lambda x: True
        else:
            assert isinstance(channel, int)
            self.is_channel = lambda x: x['ChannelIndex'] == channel
        self.pos_str = position + '/'
        self.root_str = root

    @property
    def summary(self):
        return self.raw['Summary']

    def _entries(self, prefix):
        r = [self.raw[m] for m in self.raw.keys() if m.startswith(prefix)]
        return sorted([m for m in r if self.is_channel(m)], key=lambda k: k['Frame'])

    @property
    def coords(self):
        return self._entries('Coords')

    @property
    def metadata(self):
        return self._entries('Meta')

    @property
    def x_pos(self):
        return [r['XPositionUm'] for r in self.metadata]

    @property
    def y_pos(self):
        return [r['YPositionUm'] for r in self.metadata]

    @property
    def z_pos(self):
        return [r['ZPositionUm'] for r in self.metadata]

    @property
    def image_filenames(self):
        return [r['FileName'].replace(self.pos_str, '') for r in self.metadata]

    @property
    def timestamps(self):
        return [self.convert_time_to_epoch(r['ReceivedTime']) for r in self.metadata]

    @property
    def shape(self):
        """ Return the shape of the stack: (frames, slices, width, height) """
        return (self.summary['Frames'], self.summary['Slices'], self.metadata[0]['Width'],
                self.metadata[0]['Height'])

    @property
    def start_time(self):
        return self.convert_time_to_epoch(self.summary['StartTime'])

    @staticmethod
    def convert_time_to_epoch(time_str):
        """ Micromanager time format: 2019-03-15 18:44:35.225 +0000 """
        utc_time = datetime.strptime(time_str[:23], "%Y-%m-%d %H:%M:%S.%f")
        return (utc_time - datetime(1970, 1, 1)).total_seconds()

    def get_metadata(self, idx):
        return {'filename': self.image_filenames[idx], 'x_position': self.x_pos[idx],
                'y_position': self.y_pos[idx], 'z_position': self.z_pos[idx],
                'timestamp': self.timestamps[idx]}


class MicromanagerReader(object):
    """ Reads in micromanager stacks """

    def __init__(self, filepath, channel=None):
        self.metadata = MicromanagerMetadataParser(filepath, channel)
        self._dir = filepath
        self._files = self.metadata.image_filenames
        self._n, self._s, self._w, self._h = self.metadata.shape
        self._dtype = self[0].dtype if len(self._files) else None

    @property
    def width(self):
        return self._w

    @property
    def height(self):
        return self._h

    def __len__(self):
        return len(self._files)

    def __getitem__(self, idx):
        if idx < 0 or idx >= len(self):
            raise IndexError('image %d outside the stack (0..%d)' % (idx, len(self) - 1))
        return read_tiff(os.path.join(self._dir, self._files[idx]))

    def get_metadata(self, idx):
        if idx < 0 or idx >= len(self):
            raise IndexError('image %d outside the stack (0..%d)' % (idx, len(self) - 1))
        return self.metadata.get_metadata(idx)

    @property
    def dtype(self):
        return self._dtype

    def frames_raw(self, start, count):
        """ ``count`` images from ``start`` as one contiguous (count, rows, cols) array in the stored
        type -- the batch ``UNet.segment_and_localise`` takes. """
        idx = range(int(start), min(int(start) + int(count), len(self)))
        first = self[idx[0]] if len(idx) else np.empty((0, 0), self._dtype or np.uint16)
        out = np.empty((len(idx),) + first.shape, dtype=first.dtype)
        for k, i in enumerate(idx):
            out[k] = first if k == 0 else self[i]
        return out


def write_micromanager_position(filepath, frames, channel_index=0, n_channels=1):
    """ Write ``frames`` (N,H,W) uint8 / uint16 as a Micro-Manager 2.0 position folder (synthetic stacks
    for tests): one TIFF per frame + ``metadata.txt``. """
    frames = np.asarray(frames)
    os.makedirs(filepath, exist_ok=True)
    position = os.path.basename(filepath.rstrip(os.sep))
    n, h, w = frames.shape
    meta_path = os.path.join(filepath, 'metadata.txt')
    raw = json.load(open(meta_path)) if os.path.exists(meta_path) else {}
    raw['Summary'] = {'Frames': n, 'Slices': 1, 'Channels': n_channels, 'Positions': 1,
                      'StartTime': '2019-03-15 18:44:35.225 +0000'}
    for i in range(n):
        name = 'img_channel%03d_position000_time%09d_z000.tif' % (channel_index, i)
        write_tiff(os.path.join(filepath, name), frames[i])
        key = '%s/%s' % (position, name)
        raw['Coords-' + key] = {'Frame': i, 'ChannelIndex': channel_index, 'SliceIndex': 0, 'PositionIndex': 0}
        raw['Metadata-' + key] = {'Frame': i, 'ChannelIndex': channel_index, 'Width': w, 'Height': h,
                                  'FileName': key, 'XPositionUm': 10.0 * i, 'YPositionUm': -5.0 * i,
                                  'ZPositionUm': 0.5, 'ReceivedTime': '2019-03-15 18:44:%02d.500 +0000' % (36 + i % 20)}
    with open(meta_path, 'w') as fh:
        json.dump(raw, fh)
