"""
CPU tests of kpal_b200.h5lite, the in-tree HDF5 reader / writer behind the
profile file format (reference doc/fileformat.rst:23-39, kpal/__init__.py:85-111,
kpal/klib.py:63-76,227-256) on machines without h5py.

What pins it:
  * the reader against the one libhdf5-written file in this image (SciPy's
    MATLAB v7.3 fixture: superblock v0, v1 object headers, local heap, group
    B-tree, symbol-table node, attribute and datatype messages);
  * writer -> reader round trips of everything kPAL stores, and a structural
    walk of the written bytes against the format specification's invariants
    (node sizes, sibling links, key order, heap sizes, end-of-file address);
  * the reference's OWN test-suite (tests/test_klib.py, test_kmer.py, ...: 105
    tests, among them every HDF5 save / load / CLI test) run unmodified with
    h5lite standing in for h5py -- in the build container, where the reference
    tree exists.
Byte-level interoperability with real h5py cannot be executed here (no libhdf5).
"""
import importlib.util
import os
import shutil
import struct
import subprocess
import sys

import numpy as np
import pytest

from kpal_b200 import h5lite, klib, kmer
from oracle import kpal_oracle as ko, ref_loader

UNDEF = h5lite.UNDEF


def _matlab_fixture():
    spec = importlib.util.find_spec('scipy')
    if spec is None or not spec.submodule_search_locations:
        return None
    path = os.path.join(list(spec.submodule_search_locations)[0], 'io', 'matlab', 'tests', 'data',
                        'testhdf5_7.4_GLNX86.mat')
    return path if os.path.isfile(path) else None


@pytest.mark.skipif(_matlab_fixture() is None, reason="SciPy's HDF5 fixture is not installed")
def test_reader_on_a_file_written_by_libhdf5():
    """MATLAB v7.3 = HDF5 behind a 512-byte user block, written by libhdf5 1.6/1.8: the
    same superblock / group / object header family h5py writes by default."""
    assert h5lite.is_hdf5(_matlab_fixture())
    with h5lite.File(_matlab_fixture(), 'r') as f:
        assert f.keys() == ['testdouble'] and 'testdouble' in f and 'nope' not in f
        dataset = f['testdouble']
        assert dataset.shape == (9, 1) and dataset.dtype == np.dtype('<f8')
        assert np.allclose(dataset[...].ravel(), np.arange(9) * np.pi / 4, rtol=0, atol=1e-15)
        assert dataset.attrs['MATLAB_class'] == b'double'
        with pytest.raises(KeyError):
            f['missing']


def _profile_file(path, profiles, compression='gzip'):
    with h5lite.File(path, 'w') as f:
        f.attrs['format'] = 'kMer'
        f.attrs['version'] = '1.0.0'
        f.attrs['producer'] = 'kPAL-B200 tests'
        f.create_group('profiles')
        for name, counts in profiles.items():
            klib.Profile(counts).save(f, name=name)


def _fsck(path):
    """Walks every structure of a file written by h5lite and checks what libhdf5 relies on."""
    raw = open(path, 'rb').read()
    assert raw[:8] == h5lite.SIGNATURE
    version, _fs, _rg, _r, _sh, so, sl, _r2, leaf_k, group_k, _flags = struct.unpack_from('<8B2HI', raw, 8)
    assert (version, so, sl, leaf_k, group_k) == (0, 8, 8, 4, 16)
    base, free, eof, driver = struct.unpack_from('<4Q', raw, 24)
    assert base == 0 and free == UNDEF and driver == UNDEF
    assert eof == len(raw), "end-of-file address must equal the file size"
    _name, root, cache, _r3, root_btree, root_heap = struct.unpack_from('<QQIIQQ', raw, 56)
    assert cache == 1
    seen = {'groups': 0, 'datasets': 0, 'snod': 0, 'chunks': 0}

    def messages(address):
        v, _r, n, refs, size = struct.unpack_from('<BBHII', raw, address)
        assert v == 1 and refs == 1 and size % 8 == 0 and address % 8 == 0
        out, at, end = [], address + 16, address + 16 + size
        while at < end:
            mtype, msize, _flags = struct.unpack_from('<HHB', raw, at)
            assert msize % 8 == 0
            out.append((mtype, raw[at + 8: at + 8 + msize]))
            at += 8 + msize
        assert at == end and len(out) == n, "message count / header size must match"
        return out

    def heap_names(address):
        assert raw[address:address + 4] == b'HEAP'
        size, free_head, data = struct.unpack_from('<QQQ', raw, address + 8)
        assert free_head == 1 and size % 8 == 0 and data + size <= len(raw)     # 1 = no free block
        return raw[data:data + size]

    def name_at(names, offset):
        return names[offset:names.index(b'\0', offset)]

    def group(header, btree, heap):
        seen['groups'] += 1
        kinds = messages(header)
        assert kinds[0][0] == 0x11 and struct.unpack('<QQ', kinds[0][1][:16]) == (btree, heap)
        names = heap_names(heap)
        assert names[:8] == b'\0' * 8
        collected = []

        _walk_group_tree(raw, btree, names, collected, seen, name_at)
        ordered = [c[0] for c in collected]
        assert ordered == sorted(ordered) and len(set(ordered)) == len(ordered)
        for name, hdr, ctype, scratch in collected:
            if ctype == 1:
                group(hdr, scratch[0], scratch[1])
            else:
                dataset(hdr)
        return ordered

    def dataset(header):
        seen['datasets'] += 1
        kinds = dict(messages(header)[::-1])            # first occurrence wins
        shape = struct.unpack_from('<%dQ' % kinds[0x1][1], kinds[0x1], 8)
        layout = kinds[0x8]
        assert layout[0] == 3
        if layout[1] == 1:
            address, size = struct.unpack_from('<QQ', layout, 2)
            assert address + size <= len(raw)
            return
        assert layout[1] == 2 and layout[2] == len(shape) + 1
        btree = struct.unpack_from('<Q', layout, 3)[0]
        dims = struct.unpack_from('<%dI' % (len(shape) + 1), layout, 11)
        key_size = 8 + 8 * (len(shape) + 1)
        node_size = 24 + 65 * key_size + 64 * 8
        offsets = []

        def node(address, expect_level):
            assert raw[address:address + 4] == b'TREE' and raw[address + 4] == 1
            level, used = raw[address + 5], struct.unpack_from('<H', raw, address + 6)[0]
            assert 1 <= used <= 64 and address + node_size <= len(raw)
            assert expect_level is None or level == expect_level
            for i in range(used):
                at = address + 24 + i * (key_size + 8)
                nbytes, mask = struct.unpack_from('<II', raw, at)
                offset = struct.unpack_from('<%dQ' % (len(shape) + 1), raw, at + 8)
                child = struct.unpack_from('<Q', raw, at + key_size)[0]
                following = struct.unpack_from('<%dQ' % (len(shape) + 1), raw, at + key_size + 8 + 8)
                assert offset < following and offset[-1] == 0 and mask == 0
                if level:
                    node(child, level - 1)
                else:
                    assert child + nbytes <= len(raw) and all(o % d == 0 for o, d in zip(offset, dims))
                    offsets.append(offset[:-1])
                    seen['chunks'] += 1
        node(btree, None)
        assert offsets == sorted(offsets)
        expected = 1
        for s, d in zip(shape, dims):
            expected *= -(-s // d)
        assert len(offsets) == expected, "one chunk per grid cell"

    names = group(root, root_btree, root_heap)
    return names, seen


def _walk_group_tree(raw, address, names, collected, seen, name_at):
    """Depth-first walk of a group B-tree: key order, sibling links per level, node sizes."""
    per_level = {}

    def node(address):
        assert raw[address:address + 4] == b'TREE'
        ntype, level, used, left, right = struct.unpack_from('<BBHQQ', raw, address + 4)
        assert ntype == 0 and 1 <= used <= 32 and address + 544 <= len(raw)
        per_level.setdefault(level, []).append((address, left, right))
        keys = [struct.unpack_from('<Q', raw, address + 24 + 16 * i)[0] for i in range(used + 1)]
        children = [struct.unpack_from('<Q', raw, address + 32 + 16 * i)[0] for i in range(used)]
        for i, child in enumerate(children):
            low, high = name_at(names, keys[i]), name_at(names, keys[i + 1])
            before = len(collected)
            if level:
                node(child)
            else:
                assert raw[child:child + 4] == b'SNOD' and raw[child + 4] == 1
                count = struct.unpack_from('<H', raw, child + 6)[0]
                assert 1 <= count <= 8 and child + 328 <= len(raw)
                seen['snod'] += 1
                for e in range(count):
                    off, hdr, ctype, _r = struct.unpack_from('<QQII', raw, child + 8 + 40 * e)
                    scratch = struct.unpack_from('<QQ', raw, child + 8 + 40 * e + 24)
                    collected.append((name_at(names, off), hdr, ctype, scratch))
            inside = [c[0] for c in collected[before:]]
            assert inside and all(low < n <= high for n in inside), "names of child i lie in (key[i], key[i+1]]"
    used = struct.unpack_from('<H', raw, address + 6)[0]
    if used:
        node(address)
    for level, nodes in per_level.items():
        for i, (addr, left, right) in enumerate(nodes):
            assert left == (nodes[i - 1][0] if i else UNDEF)
            assert right == (nodes[i + 1][0] if i + 1 < len(nodes) else UNDEF)


@pytest.mark.parametrize("n_profiles", (0, 1, 8, 9, 300, 2100))
def test_round_trip_and_structure(tmp_path, n_profiles):
    """0 .. 2100 profiles: empty group, one symbol-table node, and group B-trees of one, two
    and three levels (8 links per node, 32 children per B-tree node)."""
    rng = np.random.default_rng(n_profiles)
    profiles = {}
    for i in range(n_profiles):
        k = int(rng.integers(1, 6))
        profiles['p%05d' % i] = rng.integers(0, 40, 4 ** k).astype(np.int64)
    path = str(tmp_path / 'profiles.k')
    _profile_file(path, profiles)
    names, seen = _fsck(path)
    assert names == [b'profiles'] and seen['datasets'] == n_profiles and seen['groups'] == 2
    with h5lite.File(path) as f:
        assert f.attrs['format'] == 'kMer' and isinstance(f.attrs['format'], str)
        assert f.attrs['version'] == '1.0.0' and f.attrs.get('nope') is None
        assert sorted(f.attrs.keys()) == ['format', 'producer', 'version']
        group = f['profiles']
        assert list(group) == sorted(profiles) and len(group) == n_profiles
        for name, counts in profiles.items():
            dataset = group[name]
            assert dataset.dtype == np.dtype('<i8') and dataset.shape == counts.shape
            assert dataset.compression == 'gzip' and dataset.chunks == counts.shape
            assert np.array_equal(dataset[:], counts) and np.array_equal(f['/profiles/' + name][...], counts)
            attrs = dataset.attrs
            assert attrs['length'] == klib.Profile(counts).length and attrs['total'] == counts.sum()
            assert attrs['non_zero'] == np.count_nonzero(counts)
            assert attrs['mean'] == counts.mean() and attrs['std'] == counts.std()
            assert attrs['median'] == np.median(counts)
            loaded = klib.Profile.from_file(f, name=name)
            assert np.array_equal(loaded.counts, counts) and loaded.name == name
        if profiles:
            assert klib.Profile.from_file(f).name == sorted(profiles)[0]


def test_large_profile_is_chunked_like_h5py(tmp_path):
    """k = 11 (33.5 MB): 64 KiB chunks by the rule h5py applies when chunking is left to it,
    a two-level chunk B-tree (512 chunks, 64 per node), parallel (de)compression."""
    counts = np.random.default_rng(3).poisson(3.0, 4 ** 11).astype(np.int64)
    path = str(tmp_path / 'big.k')
    _profile_file(path, {'big': counts, 'small': counts[:16].copy()})
    _names, seen = _fsck(path)
    assert seen['chunks'] == 512 + 1
    assert os.path.getsize(path) < counts.nbytes // 4
    with h5lite.File(path) as f:
        assert f['profiles/big'].chunks == (8192,)
        assert np.array_equal(f['profiles/big'][:], counts)
        out = np.empty_like(counts)
        f['profiles/big'].read_direct(out)
        assert np.array_equal(out, counts)
    assert h5lite._guess_chunk((4 ** 12,), 8) == (8192,) and h5lite._guess_chunk((4 ** 8,), 8) == (2048,)
    assert h5lite._guess_chunk((4 ** 6,), 8) == (1024,) and h5lite._guess_chunk((4,), 8) == (4,)


def test_read_direct_into_slab_rows(tmp_path):
    """The access pattern of kdistlib.distance_matrix_from_file: `group[name].shape`, then
    `read_direct` into one row of a preallocated [rows][4^k] slab -- multi-chunk (k = 9),
    single-chunk (k = 3) and uncompressed datasets, concurrently from several threads."""
    from concurrent.futures import ThreadPoolExecutor
    rng = np.random.default_rng(5)
    for k in (9, 3):
        profiles = dict(('s%02d' % i, rng.poisson(2.0, 4 ** k).astype(np.int64)) for i in range(12))
        path = str(tmp_path / ('slab%d.k' % k))
        _profile_file(path, profiles)
        with h5lite.File(path) as f:
            group = f['profiles']
            names = sorted(group)
            slab = np.full((len(names), 4 ** k), -1, dtype=np.int64)
            for row, name in enumerate(names):
                assert int(group[name].shape[0]) == 4 ** k
                group[name].read_direct(slab[row])
            assert all(np.array_equal(slab[row], profiles[name]) for row, name in enumerate(names))
            slab[:] = -1
            with ThreadPoolExecutor(4) as pool:             # datasets in parallel: reads are positional
                list(pool.map(lambda rn: group[rn[1]].read_direct(slab[rn[0]]), enumerate(names)))
            assert all(np.array_equal(slab[row], profiles[name]) for row, name in enumerate(names))
            wrong = np.empty(4 ** k, dtype=np.int32)         # other dtype: converted, not reinterpreted
            group[names[0]].read_direct(wrong)
            assert np.array_equal(wrong, profiles[names[0]])
    plain = str(tmp_path / 'plain.h5')
    with h5lite.File(plain, 'w') as f:
        f.create_dataset('x', data=np.arange(100, dtype=np.int64))
    with h5lite.File(plain) as f:
        out = np.empty(100, dtype=np.int64)
        f['x'].read_direct(out)
        assert np.array_equal(out, np.arange(100))


def test_values_and_layouts(tmp_path):
    path = str(tmp_path / 'misc.h5')
    matrix = np.arange(35 * 13, dtype=np.float64).reshape(35, 13)
    with h5lite.File(path, 'w') as f:
        f.create_dataset('plain', data=np.arange(10, dtype=np.int32))                       # contiguous
        f.create_dataset('deep/er/matrix', data=matrix, compression='gzip', chunks=(8, 5))   # edge chunks
        f.create_dataset('from_list', data=[1, 2, 3], dtype='int64', compression='gzip')
        f['plain'].attrs['i'] = 7
        f['plain'].attrs['f'] = 2.5
        f['plain'].attrs['np'] = np.float32(1.5)
        f['plain'].attrs['vec'] = np.array([1, 2, 3], dtype=np.int16)
        f['plain'].attrs['bytes'] = b'abc'
        f['plain'].attrs['text'] = u'café'
        f['plain'].attrs['empty'] = ''
        assert np.array_equal(f['deep/er/matrix'][...], matrix)        # readable before close
        with pytest.raises(ValueError):
            f.create_dataset('plain', data=[1])
        with pytest.raises(ValueError):
            f.create_group('deep')
    _fsck(path)
    with h5lite.File(path) as f:
        assert f.keys() == ['deep', 'from_list', 'plain'] and f['deep'].keys() == ['er']
        assert np.array_equal(f['plain'][:], np.arange(10)) and f['plain'].dtype == np.dtype('<i4')
        assert f['plain'].compression is None and f['plain'].chunks is None
        assert np.array_equal(f['deep/er/matrix'][...], matrix) and f['deep/er/matrix'].chunks == (8, 5)
        assert np.array_equal(f['deep']['er/matrix'][2:4, 1], matrix[2:4, 1])
        assert np.array_equal(f['from_list'][:], [1, 2, 3])
        attrs = f['plain'].attrs
        assert attrs['i'] == 7 and attrs['f'] == 2.5 and attrs['np'] == 1.5
        assert attrs['np'].dtype == np.float32 and np.array_equal(attrs['vec'], [1, 2, 3])
        assert attrs['bytes'] == b'abc' and attrs['text'] == u'café' and attrs['empty'] == ''
        assert 'i' in attrs and len(attrs) == 7
        with pytest.raises(ValueError):
            f.create_group('x')                 # read-only
    with pytest.raises(ValueError):
        f.keys() and f['plain']                 # closed
    with pytest.raises(IOError):
        h5lite.File(__file__)                   # not HDF5
    assert not h5lite.is_hdf5(__file__) and h5lite.is_hdf5(path)


def test_unclosed_file_is_completed_on_collection(tmp_path):
    path = str(tmp_path / 'leak.k')
    f = h5lite.File(path, 'w')
    f.create_group('profiles')
    klib.Profile(np.arange(16, dtype=np.int64)).save(f, name='a')
    del f
    import gc
    gc.collect()
    with h5lite.File(path) as f:
        assert np.array_equal(f['profiles/a'][:], np.arange(16))
    # ... and at interpreter exit
    path2 = str(tmp_path / 'exit.k')
    code = ("import numpy as np; from kpal_b200 import h5lite; f = h5lite.File(%r, 'w'); "
            "f.create_dataset('profiles/x', data=np.arange(4), dtype='int64', compression='gzip')" % path2)
    env = dict(os.environ, PYTHONPATH=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    subprocess.run([sys.executable, '-c', code], check=True, env=env)
    with h5lite.File(path2) as f:
        assert np.array_equal(f['profiles/x'][:], np.arange(4))


def test_profile_file_type(tmp_path, monkeypatch):
    """kmer.ProfileFileType (reference kpal/__init__.py:85-111) on top of h5lite: new files get
    format / version / producer and /profiles; existing ones are validated."""
    monkeypatch.setitem(sys.modules, 'h5py', None)          # as on a machine without h5py
    import argparse
    path = str(tmp_path / 'out.k')
    handle = kmer.ProfileFileType('w')(path)
    assert isinstance(handle, h5lite.File)
    counts = ko.count_sequences(['ACGTACGTTTGACA'], 3)
    assert klib.Profile(counts).save(handle) == '1'          # first free number
    assert klib.Profile(counts, name='named').save(handle) == 'named'
    assert klib.Profile(counts).save(handle) == '2'
    handle.close()
    with pytest.raises(argparse.ArgumentTypeError):
        kmer.ProfileFileType('w')(path)                      # file exists
    handle = kmer.ProfileFileType('r')(path)
    assert sorted(handle['profiles']) == ['1', '2', 'named']
    assert handle.attrs['format'] == 'kMer' and handle.attrs['producer'].startswith('kPAL-B200')
    assert np.array_equal(klib.Profile.from_file(handle, name='named').counts, counts)
    handle.close()
    other = str(tmp_path / 'other.h5')
    with h5lite.File(other, 'w') as f:
        f.attrs['format'] = 'something else'
    with pytest.raises(argparse.ArgumentTypeError):
        kmer.ProfileFileType('r')(other)
    with pytest.raises(argparse.ArgumentTypeError):
        kmer.ProfileFileType('r')(__file__)


@pytest.mark.skipif(not ref_loader.available() or not os.path.isdir('/root/reference/tests'),
                    reason="reference tree not present (build container only)")
def test_reference_test_suite_passes_on_h5lite(tmp_path):
    """The reference's own tests (44 klib, 39 kmer/CLI, 22 kdistlib / metrics), unmodified
    apart from the nose-style hook names pytest no longer calls, against the unmodified
    reference sources -- with kpal_b200.h5lite imported as `h5py`.  Every HDF5 save / load /
    CLI test of the reference therefore exercises this reader / writer."""
    work = tmp_path / 'suite'
    shutil.copytree('/root/reference/tests', str(work))     # scratch copy, never committed
    for name in os.listdir(str(work)):
        if name.endswith('.py'):
            path = work / name
            text = path.read_text()
            text = text.replace('def setup(self)', 'def setup_method(self)')
            text = text.replace('def teardown(self)', 'def teardown_method(self)')
            text = text.replace('.setup()', '.setup_method()').replace('.teardown()', '.teardown_method()')
            path.write_text(text)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    (work / 'conftest.py').write_text(
        "import sys\n"
        "sys.path.insert(0, %r)\n"
        "from kpal_b200 import h5lite\n"
        "sys.modules['h5py'] = h5lite\n"
        "from oracle import ref_loader\n"
        "ref_loader.load()\n" % root)
    result = subprocess.run([sys.executable, '-m', 'pytest', '-q', '-p', 'no:cacheprovider', str(work)],
                            capture_output=True, text=True, cwd=str(work))
    tail = result.stdout.strip().splitlines()[-1] if result.stdout.strip() else result.stderr[-400:]
    assert result.returncode == 0, result.stdout[-3000:]
    assert ' passed' in tail and 'failed' not in tail and int(tail.split()[0]) >= 105, tail


def test_bulk_datasets_equal_the_loop(tmp_path):
    """Group.create_datasets (one call for a batch of per-record profiles: native deflate of all
    chunks, bulk chunk B-trees and object headers) writes the content the create_dataset loop
    writes -- data, chunking, filter, attribute values and types -- and a structurally valid file;
    datasets of both kinds mix in one group, and a bulk dataset whose attributes are touched
    afterwards is serialised on its own."""
    from kpal_b200 import klib
    rng = np.random.default_rng(12)
    n, k = 70, 7
    rows = rng.poisson(0.05, (n, 4 ** k)).astype(np.int64)
    rows[3] = 0
    rows[5, ::11] = 10 ** 12
    names = ['rec%03d' % i for i in range(n)]
    bulk_path, loop_path = str(tmp_path / 'bulk.k'), str(tmp_path / 'loop.k')
    with h5lite.File(bulk_path, 'w') as handle:
        handle.create_group('profiles')
        klib.Profile(rows[0], name='first').save(handle)                # an ordinary dataset before ...
        klib.save_profiles(handle, names[:40], rows[:40])
        klib.save_profiles(handle, names[40:], rows[40:])               # ... two bulk calls ...
        handle['profiles/rec007'].attrs['note'] = 'touched'              # ... one of them modified ...
        klib.Profile(rows[1], name='last').save(handle)                 # ... and one after
        assert np.array_equal(handle['profiles/rec033'][:], rows[33])   # readable before close()
    with h5lite.File(loop_path, 'w') as handle:
        handle.create_group('profiles')
        for name, row in zip(names, rows):
            klib.Profile(row, name=name).save(handle)
    _fsck(bulk_path)
    with h5lite.File(bulk_path, 'r') as a, h5lite.File(loop_path, 'r') as b:
        assert a['profiles'].keys() == sorted(names + ['first', 'last'])
        assert a['profiles/rec007'].attrs['note'] == 'touched'
        for name in names:
            da, db = a['profiles/' + name], b['profiles/' + name]
            assert np.array_equal(da[:], db[:]) and da.dtype == db.dtype
            assert da.chunks == db.chunks and da.compression == db.compression
            for key in ('length', 'total', 'non_zero', 'mean', 'median', 'std'):
                va, vb = da.attrs[key], db.attrs[key]
                assert va == vb and np.asarray(va).dtype == np.asarray(vb).dtype, (name, key)
    with pytest.raises(ValueError):
        with h5lite.File(str(tmp_path / 'dup.k'), 'w') as handle:
            handle.create_group('profiles')
            klib.save_profiles(handle, ['a', 'a'], rows[:2])
