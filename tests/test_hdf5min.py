"""The package's own HDF5 writer / reader (``sequitr_b200/hdf5min.py``; h5py and libhdf5 are not in this
image): the ``frames/frame_<i>/coords`` layout of the reference's ``CentroidWriter`` (utils.py:570-578),
checked by an independent byte-level walk of the file and against known-answer bytes of the HDF5 format
specification."""
import struct

import numpy as np
import pytest

from sequitr_b200 import hdf5min, utils


def _tables(n, seed=0):
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n):
        rows = int(rng.integers(0, 9)) if i % 7 else 0
        t = rng.random((rows, 5)).astype(np.float32)
        t[:, 0] = i
        out.append(t if rows else np.zeros((0,), np.float32))        # empty frame: shape (0,) as the reference writes
    return out


def test_round_trip_of_the_centroid_layout(tmp_path):
    fn = str(tmp_path / 'tracks.hdf5')
    tables = _tables(2000)                                            # configs[2]: 2000 frames -> a two-level B-tree
    with hdf5min.File(fn, 'w') as f:
        f.create_group('frames')
        for i, t in enumerate(tables):
            f['frames'].create_group('frame_' + str(i)).create_dataset('coords', data=t, dtype='float32')
    with hdf5min.File(fn, 'r') as r:
        assert r.keys() == ['frames'] and len(r['frames']) == 2000
        assert 'frames/frame_1999/coords' in r and 'frames/frame_2000' not in r
        for i, t in enumerate(tables):
            got = r['frames']['frame_' + str(i)]['coords']
            assert got.shape == t.shape and got.dtype == np.float32
            np.testing.assert_array_equal(got[...], t)


def test_superblock_and_known_answer_bytes(tmp_path):
    """Byte-level checks that do not go through the reader: signature, version-0 superblock fields, the root
    symbol-table entry, and the IEEE float32 datatype message exactly as the format specification encodes it."""
    fn = str(tmp_path / 'k.hdf5')
    data = np.arange(10, dtype=np.float32).reshape(2, 5)
    with hdf5min.File(fn, 'w') as f:
        f.create_dataset('frames/frame_0/coords', data=data)
    raw = open(fn, 'rb').read()
    assert raw[:8] == b'\x89HDF\r\n\x1a\n'
    assert raw[8:16] == bytes([0, 0, 0, 0, 0, 8, 8, 0])              # versions, 8-byte offsets and lengths
    assert struct.unpack_from('<HH', raw, 16) == (4, 16)              # group leaf / internal node K
    base, free, eof, drv = struct.unpack_from('<QQQQ', raw, 24)
    assert (base, free, drv) == (0, hdf5min.UNDEF, hdf5min.UNDEF) and eof == len(raw)
    name_off, root, cache, _, btree, heap = struct.unpack_from('<QQIIQQ', raw, 56)
    assert name_off == 0 and cache == 1 and root % 8 == 0
    assert raw[btree:btree + 4] == b'TREE' and raw[heap:heap + 4] == b'HEAP'
    # root object header: version 1, one message (symbol table, type 0x11) naming the same B-tree and heap
    assert struct.unpack_from('<BBHII', raw, root) == (1, 0, 1, 1, 24)
    assert struct.unpack_from('<HH', raw, root + 16) == (0x11, 16)
    assert struct.unpack_from('<QQ', raw, root + 24) == (btree, heap)
    # IEEE_F32LE datatype message body (spec IV.A.2.d): class 1 v1, bit field 20 1f 00, size 4, then bit offset 0,
    # precision 32, exponent at bit 23 (8 bits), mantissa at bit 0 (23 bits), bias 127
    f32 = bytes.fromhex('11201f0004000000' '0000' '2000' '17' '08' '00' '17' '7f000000')
    assert raw.count(f32) == 1
    i32 = bytes.fromhex('1008000004000000' '0000' '2000')
    assert hdf5min._datatype_message(np.int32) == i32
    # the raw data sits contiguously, little-endian, at the address the layout message names
    at = raw.index(data.tobytes())
    assert at % 8 == 0 and struct.pack('<BBQQ', 3, 1, at, data.nbytes) in raw
    # every group node the reader will touch carries its signature; names live in a local heap
    assert raw.count(b'SNOD') == 3 and raw.count(b'HEAP') == 3 and b'frame_0\0' in raw and b'coords\0' in raw


def test_group_btree_invariants(tmp_path):
    """Walk the B-tree of a 300-entry group by hand: sorted names, keys = last name of the left subtree,
    sibling links, node occupancy within 2K."""
    fn = str(tmp_path / 'b.hdf5')
    with hdf5min.File(fn, 'w') as f:
        for i in range(300):
            f.create_dataset('g/d%03d' % (i * 7 % 300), data=np.int32([i]))
    raw = open(fn, 'rb').read()
    rd = hdf5min._Reader(raw)
    msgs = dict(rd.messages(rd.root_header))
    tree, heap = struct.unpack('<QQ', msgs[0x11][:16])
    (gname, goh), = rd.tree_entries(tree, heap)
    assert gname == 'g'
    tree, heap = struct.unpack('<QQ', dict(rd.messages(goh))[0x11][:16])
    level, used = struct.unpack_from('<BH', raw, tree + 5)
    assert level == 1 and used == 2                                   # 300 names / 8 per SNOD = 38 SNODs / 32 per node
    names, prev_right = [], hdf5min.UNDEF
    for i in range(used):
        child, key_after = struct.unpack_from('<QQ', raw, tree + 32 + 16 * i)
        clevel, cused, left, right = struct.unpack_from('<BHQQ', raw, child + 5)
        assert clevel == 0 and 1 <= cused <= 32
        assert left == (hdf5min.UNDEF if i == 0 else prev_child) and (i + 1 < used) == (right != hdf5min.UNDEF)
        for j in range(cused):
            snod, key = struct.unpack_from('<QQ', raw, child + 32 + 16 * j)
            n, = struct.unpack_from('<H', raw, snod + 6)
            assert raw[snod:snod + 4] == b'SNOD' and 1 <= n <= 8
            here = [rd.heap_string(heap, struct.unpack_from('<Q', raw, snod + 8 + 40 * e)[0]) for e in range(n)]
            assert rd.heap_string(heap, key) == here[-1]              # key = greatest name of the subtree to its left
            names += here
        assert rd.heap_string(heap, key_after) == names[-1]
        prev_child = child
    assert names == sorted(names) and len(names) == 300
    with hdf5min.File(fn, 'r') as r:
        assert [int(r['g'][k][0]) for k in r['g'].keys()][:3] == [0, 43, 86]     # d000, d001 (i=43: 301 % 300), d002


def test_errors_and_types(tmp_path):
    fn = str(tmp_path / 'e.hdf5')
    with hdf5min.File(fn, 'w') as f:
        g = f.create_group('a')
        with pytest.raises(ValueError):
            f.create_group('a')
        with pytest.raises(TypeError):
            g.create_dataset('c', data=np.array(['x']))
        g.create_dataset('u16', data=np.arange(6, dtype=np.uint16).reshape(2, 3))
        g.create_dataset('f64', data=np.float64([1.5, -2.25]))
        f.create_group('empty')
    with hdf5min.File(fn, 'r') as r:
        assert r['a/u16'].dtype == np.uint16 and r['a/u16'][1, 2] == 5
        np.testing.assert_array_equal(r['a/f64'][...], [1.5, -2.25])
        assert len(r['empty']) == 0
        with pytest.raises(KeyError):
            r['a/missing']
    bad = tmp_path / 'bad.hdf5'
    bad.write_bytes(b'not hdf5 at all, sorry' * 10)
    with pytest.raises(IOError):
        hdf5min.File(str(bad), 'r')


def test_handler_reads_back_what_it_wrote(tmp_path):
    """HDF5FileHandler (utils.py:431-474): write mode then read-only mode on the same file."""
    fn = str(tmp_path / 'h.hdf5')
    h = utils.HDF5FileHandler(fn)
    h.hdf.create_group('frames').create_group('frame_0').create_dataset('coords', data=np.ones((3, 5)), dtype='float32')
    h.close()
    r = utils.HDF5FileHandler(fn, read_only=True)
    assert r.hdf['frames']['frame_0']['coords'].shape == (3, 5)
    r.close()


def test_random_trees_round_trip(tmp_path):
    """Seeded random group trees (depth <= 4, 0-40 children per group, names of 1-40 characters incl. UTF-8, every
    supported dtype, ranks 0-3, empty arrays): written, parsed back, compared."""
    rng = np.random.default_rng(7)
    dtypes = [np.float32, np.float64, np.int8, np.uint8, np.int16, np.uint16, np.int32, np.uint32, np.int64, np.uint64]
    alphabet = list('abcXYZ_0123456789-. ') + ['é', 'µ']

    def name():
        return ''.join(rng.choice(alphabet, size=int(rng.integers(1, 41)))).strip() or 'n'

    def fill(group, depth, want):
        for _ in range(int(rng.integers(0, 41 if depth < 2 else 6))):
            nm = name()
            if nm in want or '/' in nm:
                continue
            if depth < 4 and rng.random() < 0.3:
                want[nm] = {}
                fill(group.create_group(nm), depth + 1, want[nm])
            else:
                shape = tuple(int(v) for v in rng.integers(0, 5, size=int(rng.integers(0, 4))))
                arr = (rng.random(shape) * 200 - 100).astype(dtypes[int(rng.integers(len(dtypes)))])
                group.create_dataset(nm, data=arr)
                want[nm] = arr

    def check(group, want):
        assert group.keys() == sorted(want)
        for nm, v in want.items():
            if isinstance(v, dict):
                check(group[nm], v)
            else:
                got = group[nm]
                assert got.shape == v.shape and got.dtype == v.dtype
                np.testing.assert_array_equal(np.asarray(got), v)

    for trial in range(5):
        fn = str(tmp_path / ('r%d.hdf5' % trial))
        want = {}
        with hdf5min.File(fn, 'w') as f:
            fill(f, 0, want)
        with hdf5min.File(fn, 'r') as r:
            check(r, want)
