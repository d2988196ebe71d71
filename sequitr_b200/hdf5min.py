"""A minimal HDF5 writer and reader for the one layout the hot path produces:
``frames/frame_<i>/coords`` (``utils.py:431-474, 570-578`` of the reference write it through h5py).

h5py / libhdf5 are not part of this image, so ``CentroidWriter`` cannot lean on them.  This module
writes the subset of the HDF5 file format (version-0 superblock, "old style" groups = version-1
B-tree + local heap + symbol-table nodes, version-1 object headers, contiguous little-endian
datasets of fixed-point / IEEE types) that every libhdf5 >= 1.6 reads, and an independent parser of
the same structures (``File(name, 'r')``) that the tests use to check the bytes.  Layout follows
the "HDF5 File Format Specification Version 1.1" (III.A superblock v0, III.B B-trees v1, III.C symbol
table nodes, III.D local heaps, IV.A object headers v1, IV.A.2 messages 0x0001 dataspace v1, 0x0003
datatype v1, 0x0005 fill value v2, 0x0008 layout v3, 0x0011 symbol table).

Only what ``CentroidWriter`` needs is supported: create_group / create_dataset / item access /
keys / close.  No attributes, chunking, compression, or in-place modification.
"""
import struct

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF
SIGNATURE = b'\x89HDF\r\n\x1a\n'
LEAF_K, INTERNAL_K = 4, 16                   # symbol-table node / B-tree node fan-out parameters (defaults)
SNOD_ENTRIES, TREE_CHILDREN = 2 * LEAF_K, 2 * INTERNAL_K
HEAP_FREE_NULL = 1                           # libhdf5's "no free block" marker of a local heap


def _pad8(n):
    return (n + 7) & ~7


# --------------------------------------------------------------------------- in-memory tree
class Dataset(object):
    def __init__(self, data):
        self._data = data

    shape = property(lambda self: self._data.shape)
    dtype = property(lambda self: self._data.dtype)

    def __getitem__(self, key):
        return self._data[key]

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self._data, dtype=dtype)

    def __len__(self):
        return len(self._data)


class Group(object):
    def __init__(self):
        self._children = {}

    def _walk(self, name, create=False):
        node = self
        for part in [p for p in name.split('/') if p]:
            if part not in node._children:
                if not create:
                    raise KeyError(name)
                node._children[part] = Group()
            node = node._children[part]
            if not isinstance(node, Group) and create:
                raise ValueError('%s is a dataset' % part)
        return node

    def create_group(self, name):
        parent, _, leaf = name.rstrip('/').rpartition('/')
        node = self._walk(parent, create=True)
        if leaf in node._children:
            raise ValueError('Unable to create group (name already exists): %s' % name)
        node._children[leaf] = Group()
        return node._children[leaf]

    def create_dataset(self, name, data=None, dtype=None, shape=None):
        parent, _, leaf = name.rstrip('/').rpartition('/')
        node = self._walk(parent, create=True)
        if leaf in node._children:
            raise ValueError('Unable to create dataset (name already exists): %s' % name)
        if data is None:
            data = np.zeros(shape, dtype=dtype or 'float32')
        arr = np.asarray(data, dtype=dtype)
        if arr.ndim:                                 # (ascontiguousarray would turn a scalar into shape (1,))
            arr = np.ascontiguousarray(arr)
        _datatype_message(arr.dtype)             # fail now for unsupported types
        node._children[leaf] = Dataset(arr)
        return node._children[leaf]

    def __getitem__(self, name):
        return self._walk(name)

    def __contains__(self, name):
        try:
            self._walk(name)
            return True
        except KeyError:
            return False

    def keys(self):
        return sorted(self._children)

    def __iter__(self):
        return iter(self.keys())

    def __len__(self):
        return len(self._children)

    def items(self):
        return [(k, self._children[k]) for k in self.keys()]


# --------------------------------------------------------------------------- writer
def _datatype_message(dt):
    """Datatype message (0x0003) body, version 1, for little-endian fixed-point and IEEE floats."""
    dt = np.dtype(dt)
    if dt.byteorder == '>':
        raise TypeError('big-endian data is not supported')
    if dt.kind == 'f' and dt.itemsize in (4, 8):
        exp_bits, man_bits = (8, 23) if dt.itemsize == 4 else (11, 52)
        bits = dt.itemsize * 8
        head = struct.pack('<BBBBI', 0x11, 0x20, bits - 1, 0, dt.itemsize)   # class 1, implied-msb mantissa, sign bit
        prop = struct.pack('<HHBBBBI', 0, bits, man_bits, exp_bits, 0, man_bits, (1 << (exp_bits - 1)) - 1)
        return head + prop
    if dt.kind in 'iu' and dt.itemsize in (1, 2, 4, 8):
        head = struct.pack('<BBBBI', 0x10, 0x08 if dt.kind == 'i' else 0x00, 0, 0, dt.itemsize)
        return head + struct.pack('<HH', 0, dt.itemsize * 8)
    raise TypeError('unsupported dataset type %r' % dt)


def _message(mtype, body):
    body = body + b'\0' * (_pad8(len(body)) - len(body))
    return struct.pack('<HHBBBB', mtype, len(body), 0, 0, 0, 0) + body


def _object_header(messages):
    blob = b''.join(messages)
    return struct.pack('<BBHII', 1, 0, len(messages), 1, len(blob)) + b'\0' * 4 + blob


class _Writer(object):
    def __init__(self):
        self.buf = bytearray(96)                 # the superblock is filled in last

    def put(self, blob):
        """Append an 8-byte-aligned structure, return its address."""
        self.buf += b'\0' * (_pad8(len(self.buf)) - len(self.buf))
        addr = len(self.buf)
        self.buf += blob
        return addr

    def dataset(self, ds):
        arr = ds._data
        addr = self.put(arr.tobytes()) if arr.size else UNDEF
        space = struct.pack('<BBB5x', 1, arr.ndim, 0) + b''.join(struct.pack('<Q', d) for d in arr.shape)
        fill = struct.pack('<BBBBI', 2, 2, 0, 1, 0)          # v2: late allocation, fill at allocation, default value
        layout = struct.pack('<BBQQ', 3, 1, addr, arr.nbytes)
        return self.put(_object_header([_message(0x0001, space), _message(0x0003, _datatype_message(arr.dtype)),
                                        _message(0x0005, fill), _message(0x0008, layout)]))

    def group(self, grp):
        """Write a group (children first); returns (object header address, B-tree address, heap address)."""
        names = sorted(grp._children, key=lambda s: s.encode('utf-8'))
        entries = []                              # (name, header address, cache type, scratch)
        for name in names:
            child = grp._children[name]
            if isinstance(child, Group):
                oh, bt, hp = self.group(child)
                entries.append((name, oh, 1, struct.pack('<QQ', bt, hp)))
            else:
                entries.append((name, self.dataset(child), 0, b'\0' * 16))
        # local heap: offset 0 holds the empty string every leftmost B-tree key points at
        heap_data, offsets = bytearray(8), {}
        for name in names:
            offsets[name] = len(heap_data)
            raw = name.encode('utf-8') + b'\0'
            heap_data += raw + b'\0' * (_pad8(len(raw)) - len(raw))
        data_addr = self.put(bytes(heap_data))
        heap_addr = self.put(b'HEAP' + struct.pack('<B3xQQQ', 0, len(heap_data), HEAP_FREE_NULL, data_addr))
        # symbol-table nodes, up to 2*LEAF_K entries each
        level = []                                # (address, heap offset of the last name below it)
        for i in range(0, len(entries), SNOD_ENTRIES):
            chunk = entries[i:i + SNOD_ENTRIES]
            blob = b'SNOD' + struct.pack('<BBH', 1, 0, len(chunk))
            for name, oh, cache, scratch in chunk:
                blob += struct.pack('<QQII', offsets[name], oh, cache, 0) + scratch
            blob += b'\0' * (8 + SNOD_ENTRIES * 40 - len(blob))
            level.append((self.put(blob), offsets[chunk[-1][0]]))
        # version-1 B-tree over them, bottom-up; an empty group keeps one empty leaf node
        node_bytes = 24 + (2 * TREE_CHILDREN + 1) * 8
        depth = 0
        while True:
            groups = [level[i:i + TREE_CHILDREN] for i in range(0, len(level), TREE_CHILDREN)] or [[]]
            base = self.put(b'\0' * (node_bytes * len(groups)))
            nxt, first_key = [], 0
            for gi, kids in enumerate(groups):
                left = base + (gi - 1) * node_bytes if gi > 0 else UNDEF
                right = base + (gi + 1) * node_bytes if gi + 1 < len(groups) else UNDEF
                blob = b'TREE' + struct.pack('<BBHQQ', 0, depth, len(kids), left, right) + struct.pack('<Q', first_key)
                for addr, last in kids:
                    blob += struct.pack('<QQ', addr, last)
                    first_key = last
                a = base + gi * node_bytes
                self.buf[a:a + len(blob)] = blob
                nxt.append((a, first_key))
            if len(nxt) == 1:
                tree_addr = nxt[0][0]
                break
            level, depth = nxt, depth + 1
        oh_addr = self.put(_object_header([_message(0x0011, struct.pack('<QQ', tree_addr, heap_addr))]))
        return oh_addr, tree_addr, heap_addr

    def finish(self, root):
        oh, bt, hp = self.group(root)
        self.buf += b'\0' * (_pad8(len(self.buf)) - len(self.buf))
        sb = SIGNATURE + struct.pack('<BBBBBBBBHHI', 0, 0, 0, 0, 0, 8, 8, 0, LEAF_K, INTERNAL_K, 0)
        sb += struct.pack('<QQQQ', 0, UNDEF, len(self.buf), UNDEF)
        sb += struct.pack('<QQII', 0, oh, 1, 0) + struct.pack('<QQ', bt, hp)
        assert len(sb) == 96
        self.buf[0:96] = sb
        return bytes(self.buf)


# --------------------------------------------------------------------------- reader
class _Reader(object):
    """Parses the structures above straight from the bytes (independent of the writer's helpers)."""

    def __init__(self, raw):
        self.raw = raw
        if raw[:8] != SIGNATURE:
            raise IOError('not an HDF5 file (bad signature)')
        ver, _, _, _, _, so, sl = struct.unpack_from('<7B', raw, 8)
        if ver != 0 or so != 8 or sl != 8:
            raise IOError('unsupported superblock (version %d, offsets %d, lengths %d)' % (ver, so, sl))
        self.leaf_k, self.internal_k = struct.unpack_from('<HH', raw, 16)
        base, _, self.eof, _ = struct.unpack_from('<QQQQ', raw, 24)
        if base != 0 or self.eof > len(raw):
            raise IOError('truncated or relocated HDF5 file')
        _, self.root_header, _, _ = struct.unpack_from('<QQII', raw, 56)

    def messages(self, addr):
        ver, _, nmsg, _, size = struct.unpack_from('<BBHII', self.raw, addr)
        if ver != 1:
            raise IOError('unsupported object header version %d' % ver)
        out, pos, end = [], addr + 16, addr + 16 + size
        while pos < end and len(out) < nmsg:
            mtype, msize, _ = struct.unpack_from('<HHB', self.raw, pos)
            out.append((mtype, self.raw[pos + 8:pos + 8 + msize]))
            pos += 8 + msize
        return out

    def heap_string(self, heap_addr, off):
        if self.raw[heap_addr:heap_addr + 4] != b'HEAP':
            raise IOError('bad local heap signature')
        size, _, data = struct.unpack_from('<QQQ', self.raw, heap_addr + 8)
        if off >= size:
            raise IOError('name offset outside the local heap')
        end = self.raw.index(b'\0', data + off)
        return self.raw[data + off:end].decode('utf-8')

    def tree_entries(self, node, heap):
        sig = self.raw[node:node + 4]
        if sig == b'TREE':
            ntype, level, used = struct.unpack_from('<BBH', self.raw, node + 4)
            if ntype != 0:
                raise IOError('not a group B-tree')
            out = []
            for i in range(used):
                child, = struct.unpack_from('<Q', self.raw, node + 24 + 8 + i * 16)
                out += self.tree_entries(child, heap)
            return out
        if sig == b'SNOD':
            n, = struct.unpack_from('<H', self.raw, node + 6)
            out = []
            for i in range(n):
                off, oh, _, _ = struct.unpack_from('<QQII', self.raw, node + 8 + i * 40)
                out.append((self.heap_string(heap, off), oh))
            return out
        raise IOError('unknown group node signature %r' % sig)

    def obj(self, addr):
        msgs = dict(self.messages(addr))
        if 0x0011 in msgs:
            tree, heap = struct.unpack('<QQ', msgs[0x0011][:16])
            g = Group()
            for name, oh in self.tree_entries(tree, heap):
                g._children[name] = self.obj(oh)
            return g
        space, dtype, layout = msgs[0x0001], msgs[0x0003], msgs[0x0008]
        sver, rank = struct.unpack_from('<BB', space, 0)
        if sver != 1:
            raise IOError('unsupported dataspace version %d' % sver)
        shape = struct.unpack_from('<%dQ' % rank, space, 8)
        cls, bits0, _, _, size = struct.unpack_from('<BBBBI', dtype, 0)
        if cls & 0x0F == 1:
            npdt = np.dtype('<f%d' % size)
        elif cls & 0x0F == 0:
            npdt = np.dtype('<%s%d' % ('i' if bits0 & 0x08 else 'u', size))
        else:
            raise IOError('unsupported datatype class %d' % (cls & 0x0F))
        if bits0 & 1:
            raise IOError('big-endian data is not supported')
        lver, lclass, daddr, dsize = struct.unpack_from('<BBQQ', layout, 0)
        if lver != 3 or lclass != 1:
            raise IOError('unsupported data layout (version %d, class %d)' % (lver, lclass))
        count = int(np.prod(shape)) if rank else 1
        if dsize != count * npdt.itemsize:
            raise IOError('dataset size does not match its dataspace')
        if count == 0:
            return Dataset(np.zeros(shape, npdt))
        return Dataset(np.frombuffer(self.raw, npdt, count, daddr).reshape(shape).copy())


class File(Group):
    """``File(name, 'w')`` collects groups / datasets and writes the file on ``close()``;
    ``File(name, 'r')`` (or ``'r+'``: read access only) parses an existing one."""

    def __init__(self, filename, mode='r'):
        Group.__init__(self)
        self.filename, self.mode = filename, mode
        self._open = True
        if mode in ('r', 'r+'):
            with open(filename, 'rb') as fh:
                rd = _Reader(fh.read())
            self._children = rd.obj(rd.root_header)._children
        elif mode == 'w':
            open(filename, 'wb').close()          # fail early if the path is not writable
        else:
            raise ValueError('mode must be r, r+ or w')

    def close(self):
        if self._open and self.mode == 'w':
            with open(self.filename, 'wb') as fh:
                fh.write(_Writer().finish(self))
        self._open = False

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
