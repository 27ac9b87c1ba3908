"""
A small native reader / writer for the subset of HDF5 that *k*-mer profile
files use, with the part of the ``h5py`` interface that kPAL touches.

The on-disk format of kPAL profiles is HDF5 (reference doc/fileformat.rst:23-39):
three string attributes on the root group (``format``, ``version``,
``producer``, reference kpal/__init__.py:92-97), one group ``/profiles`` and one
``int64``, gzip-compressed dataset per profile with six numeric attributes
(reference kpal/klib.py:246-253).  The reference does all of this through
``h5py``; ``kpal_b200.kmer.ProfileFileType`` uses ``h5py`` as well when it is
installed and falls back to this module when it is not, so that ``kpal count`` /
``kpal matrix`` work end to end on a machine without libhdf5.

What is written is what libhdf5 writes for such a file with its default
("earliest") format bounds -- the structures of the HDF5 File Format
Specification, version 0 superblock family:

* superblock version 0, root symbol-table entry;
* groups as version-1 object headers with a symbol-table message: version-1
  B-tree (node type 0) over symbol-table nodes ``SNOD``, names in a local heap;
* datasets as version-1 object headers: dataspace v1, datatype (fixed point /
  IEEE float), fill value v2, layout v3 (chunked, version-1 B-tree node type 1
  over the chunks; or contiguous), filter pipeline v1 (deflate);
* attributes as version-1 attribute messages; Python ``str`` values as
  variable-length UTF-8 strings in a global heap collection (what h5py does, and
  what makes ``handle.attrs['format'] == 'kMer'`` true for the reference,
  kpal/__init__.py:99).

The reader accepts the same family (plus attribute message versions 2/3,
dataspace v2, compact layout, shuffle / fletcher32 filters and continuation
blocks), which covers files written by h5py with default settings.

PARITY NOTE: neither h5py nor libhdf5 exists in this image, so byte-level
interoperability with them cannot be executed here.  What is pinned: the reader
parses the one libhdf5-written file found in the image (SciPy's MATLAB v7.3
fixture, tests/test_h5lite.py), and writer -> reader round trips.  Everything
else follows the published format specification.

Only 'r' and 'w' modes exist.  A file opened with 'w' streams the (compressed)
chunk data to disk as datasets are created and writes all metadata on
``close()`` (also run at interpreter exit and on garbage collection).
"""
from __future__ import annotations

import atexit
import os
import struct
import weakref
import zlib
from concurrent.futures import ThreadPoolExecutor

import numpy as np

__all__ = ['File', 'Group', 'Dataset', 'AttributeManager', 'is_hdf5']

SIGNATURE = b'\x89HDF\r\n\x1a\n'
UNDEF = 0xFFFFFFFFFFFFFFFF
LEAF_K = 4              # symbol-table node holds up to 2 * LEAF_K entries
GROUP_K = 16            # group B-tree node holds up to 2 * GROUP_K children
CHUNK_K = 32            # chunk B-tree node holds up to 2 * CHUNK_K children (libhdf5 default)
GZIP_LEVEL = 4          # h5py's default for compression='gzip'
_DEFERRED_BYTES = 4 << 20       # datasets up to this size are deflated in the background (Dataset._write)
_MAX_PENDING = 64               # ... at most this many at a time

MSG_NIL, MSG_DATASPACE, MSG_DATATYPE, MSG_FILL_OLD, MSG_FILL = 0x0, 0x1, 0x3, 0x4, 0x5
MSG_LAYOUT, MSG_FILTERS, MSG_ATTRIBUTE, MSG_CONTINUATION, MSG_SYMBOL_TABLE = 0x8, 0xB, 0xC, 0x10, 0x11

_pool = None


def _workers():
    """Shared thread pool for (de)compressing chunks; zlib releases the GIL."""
    global _pool
    if _pool is None:
        _pool = ThreadPoolExecutor(max_workers=min(16, os.cpu_count() or 1))
    return _pool


def _pad8(n):
    return (n + 7) & ~7


def is_hdf5(path):
    """True if `path` starts with an HDF5 superblock (at offset 0, 512, 1024, ...)."""
    try:
        with open(path, 'rb') as handle:
            return _find_superblock(handle) is not None
    except OSError:
        return False


def _find_superblock(handle):
    handle.seek(0, os.SEEK_END)
    size = handle.tell()
    offset = 0
    while offset + 8 <= size:
        handle.seek(offset)
        if handle.read(8) == SIGNATURE:
            return offset
        offset = 512 if offset == 0 else offset * 2
    return None


# =========================================================================
# datatypes
# =========================================================================
class _VlenString(object):
    """Marker dtype for variable-length strings."""
    def __init__(self, utf8=True):
        self.utf8 = utf8


def _encode_datatype(dtype):
    """Datatype message body for a NumPy dtype or a string marker."""
    if isinstance(dtype, _VlenString):
        base = struct.pack('<B3BI2H', 0x10, 0, 0, 0, 1, 0, 8)            # 1-byte unsigned integer
        bits0 = 0x01                                                    # type = string, padding = null terminated
        return struct.pack('<B3BI', 0x19, bits0, 0x01 if dtype.utf8 else 0x00, 0, 16) + base
    dtype = np.dtype(dtype)
    if dtype.kind == 'S':                                               # fixed length, null padded, ASCII
        return struct.pack('<B3BI', 0x13, 0x01, 0, 0, dtype.itemsize)
    big = dtype.byteorder == '>'
    if dtype.kind in 'iu':
        bits0 = (1 if big else 0) | (0x08 if dtype.kind == 'i' else 0)
        return struct.pack('<B3BI2H', 0x10, bits0, 0, 0, dtype.itemsize, 0, 8 * dtype.itemsize)
    if dtype.kind == 'f' and dtype.itemsize in (4, 8):
        exp_bits, man_bits, bias = (8, 23, 127) if dtype.itemsize == 4 else (11, 52, 1023)
        bits0 = (1 if big else 0) | 0x20                                # mantissa normalisation: msb implied
        return struct.pack('<B3BI2H4BI', 0x11, bits0, 8 * dtype.itemsize - 1, 0, dtype.itemsize,
                           0, 8 * dtype.itemsize, man_bits, exp_bits, 0, man_bits, bias)
    raise TypeError('h5lite cannot store dtype %r' % (dtype,))


def _decode_datatype(buf, at=0):
    """-> (dtype or _VlenString, bytes consumed)."""
    head, b0, b1, _b2, size = struct.unpack_from('<B3BI', buf, at)
    cls, version = head & 0x0F, head >> 4
    if version not in (1, 2, 3):
        raise IOError('unsupported datatype message version %d' % version)
    if cls == 0:
        order = '>' if b0 & 1 else '<'
        kind = 'i' if b0 & 0x08 else 'u'
        return np.dtype('%s%s%d' % (order, kind, size)), 12
    if cls == 1:
        order = '>' if b0 & 1 else '<'
        return np.dtype('%sf%d' % (order, size)), 20
    if cls == 3:
        return np.dtype('S%d' % size), 8
    if cls == 9:
        if (b0 & 0x0F) != 1:
            raise IOError('variable-length sequences are not supported, only strings')
        _, used = _decode_datatype(buf, at + 8)
        return _VlenString(utf8=bool(b1 & 0x0F)), 8 + used
    raise IOError('unsupported datatype class %d' % cls)


# =========================================================================
# reading
# =========================================================================
class _Reader(object):
    """Low-level access to an existing file."""

    def __init__(self, handle):
        self.handle = handle
        start = _find_superblock(handle)
        if start is None:
            raise IOError('not an HDF5 file (no superblock signature)')
        handle.seek(start)
        block = handle.read(96 + 32)
        version = block[8]
        if version not in (0, 1):
            raise IOError('HDF5 superblock version %d is not supported (only the version 0 / 1 family '
                          'that h5py writes by default)' % version)
        if block[13] != 8 or block[14] != 8:
            raise IOError('only 8-byte offsets and lengths are supported')
        at = 24 if version == 0 else 28
        self.base, _free, self.eof, _driver = struct.unpack_from('<4Q', block, at)
        entry = at + 32
        _name, self.root_header, cache, _r = struct.unpack_from('<QQII', block, entry)
        self._global_heaps = {}

    def read(self, address, size):
        # positional read: no shared file offset, so chunk tasks may read concurrently
        data = os.pread(self.handle.fileno(), size, self.base + address)
        if len(data) != size:
            raise IOError('truncated HDF5 file (wanted %d bytes at %d)' % (size, address))
        return data

    # ---- object headers
    def messages(self, address):
        """[(type, flags, body bytes)] of a version-1 object header, continuation blocks included."""
        prefix = self.read(address, 16)
        if prefix[:4] == b'OHDR':
            raise IOError('version 2 object headers (libver="latest" files) are not supported')
        version, _r, n_messages, _refs, size = struct.unpack_from('<BBHII', prefix, 0)
        if version != 1:
            raise IOError('unsupported object header version %d' % version)
        blocks = [(address + 16, size)]
        out = []
        while blocks:
            start, length = blocks.pop(0)
            data = self.read(start, length)
            at = 0
            while at + 8 <= length and len(out) < n_messages:
                mtype, msize, flags = struct.unpack_from('<HHB', data, at)
                body = data[at + 8: at + 8 + msize]
                at += 8 + msize
                if mtype == MSG_CONTINUATION:
                    blocks.append(struct.unpack_from('<QQ', body, 0))
                out.append((mtype, flags, body))
        return out

    # ---- groups
    def group_links(self, btree, heap):
        """{name: object header address} of an old-style group."""
        head = self.read(heap, 32)
        if head[:4] != b'HEAP':
            raise IOError('bad local heap signature')
        data_size, _free, data_addr = struct.unpack_from('<QQQ', head, 8)
        names = self.read(data_addr, data_size)
        links = {}

        def walk(node_addr):
            node = self.read(node_addr, 24)
            if node[:4] != b'TREE':
                raise IOError('bad B-tree signature')
            node_type, level, used = struct.unpack_from('<BBH', node, 4)
            if node_type != 0:
                raise IOError('group B-tree expected')
            body = self.read(node_addr + 24, (2 * used + 1) * 8)
            for i in range(used):
                child = struct.unpack_from('<Q', body, 8 + 16 * i)[0]
                if level > 0:
                    walk(child)
                    continue
                snod = self.read(child, 8)
                if snod[:4] != b'SNOD':
                    raise IOError('bad symbol table node signature')
                count = struct.unpack_from('<H', snod, 6)[0]
                entries = self.read(child + 8, 40 * count)
                for e in range(count):
                    name_off, header = struct.unpack_from('<QQ', entries, 40 * e)
                    end = names.index(b'\0', name_off)
                    links[names[name_off:end].decode('utf-8')] = header
        if btree != UNDEF:
            walk(btree)
        return links

    # ---- global heap (variable-length data)
    def global_heap_object(self, address, index):
        heap = self._global_heaps.get(address)
        if heap is None:
            head = self.read(address, 16)
            if head[:4] != b'GCOL':
                raise IOError('bad global heap signature')
            size = struct.unpack_from('<Q', head, 8)[0]
            data = self.read(address, size)
            heap = {}
            at = 16
            while at + 16 <= size:
                idx, _refs, _r, osize = struct.unpack_from('<HHIQ', data, at)
                if idx == 0:
                    break
                heap[idx] = data[at + 16: at + 16 + osize]
                at += 16 + _pad8(osize)
            self._global_heaps[address] = heap
        return heap[index]

    # ---- values
    def decode_values(self, dtype, shape, raw):
        count = int(np.prod(shape)) if shape else 1
        if isinstance(dtype, _VlenString):
            values = []
            for i in range(count):
                length, heap, index = struct.unpack_from('<IQI', raw, 16 * i)
                data = self.global_heap_object(heap, index)[:length] if length else b''
                values.append(data.decode('utf-8') if dtype.utf8 else data.decode('latin-1'))
            if shape == ():
                return values[0]
            return np.array(values, dtype=object).reshape(shape)
        array = np.frombuffer(raw, dtype=dtype, count=count)
        if shape == ():
            return array[0]
        return array.reshape(shape).copy()


def _decode_dataspace(body):
    version, rank, flags = body[0], body[1], body[2]
    if version == 1:
        at = 8
    elif version == 2:
        if body[3] == 2:
            return None                                  # null dataspace
        at = 4
    else:
        raise IOError('unsupported dataspace version %d' % version)
    return tuple(struct.unpack_from('<%dQ' % rank, body, at)) if rank else ()


def _decode_attribute(reader, body):
    version = body[0]
    if version == 1:
        name_size, type_size, space_size = struct.unpack_from('<HHH', body, 2)
        at = 8
        pad = _pad8
    elif version in (2, 3):
        name_size, type_size, space_size = struct.unpack_from('<HHH', body, 2)
        at = 8 if version == 2 else 9
        pad = lambda n: n                              # noqa: E731 -- no padding from version 2 on
        if body[1] & 0x03:
            raise IOError('shared attribute datatypes / dataspaces are not supported')
    else:
        raise IOError('unsupported attribute message version %d' % version)
    name = body[at: at + name_size].split(b'\0', 1)[0].decode('utf-8')
    at += pad(name_size)
    dtype, _ = _decode_datatype(body, at)
    at += pad(type_size)
    shape = _decode_dataspace(body[at: at + space_size])
    at += pad(space_size)
    if shape is None:
        return name, None
    return name, reader.decode_values(dtype, shape, body[at:])


# =========================================================================
# the h5py-like objects
# =========================================================================
class AttributeManager(object):
    """``obj.attrs``: a mapping of names to scalars / small arrays / strings."""

    def __init__(self, owner):
        self._owner = owner

    def _load(self):
        return self._owner._attributes()

    def __getitem__(self, name):
        return self._load()[name]

    def get(self, name, default=None):
        return self._load().get(name, default)

    def __contains__(self, name):
        return name in self._load()

    def __iter__(self):
        return iter(self._load())

    def __len__(self):
        return len(self._load())

    def keys(self):
        return self._load().keys()

    def items(self):
        return self._load().items()

    def __setitem__(self, name, value):
        self._owner._file._require_writable()
        if isinstance(value, str):
            pass
        elif isinstance(value, bytes):
            value = np.bytes_(value)
        else:
            value = np.asarray(value)
            if value.dtype.kind == 'U':
                value = str(value[()]) if value.shape == () else value
            if not isinstance(value, str):
                if value.dtype.kind == 'b':
                    value = value.astype(np.int8)
                if value.dtype.kind not in 'iufS':
                    raise TypeError('h5lite cannot store an attribute of type %r' % (value.dtype,))
                value = value[()] if value.shape == () else value
        self._owner._attributes()[name] = value


class _Node(object):
    """Common part of groups and datasets."""

    def __init__(self, file, name, header=None):
        self._file = file
        self.name = name                   # absolute path
        self._header = header              # object header address (files opened for reading)
        self._attrs = {} if header is None else None
        self._messages = None

    @property
    def file(self):
        return self._file

    @property
    def attrs(self):
        return AttributeManager(self)

    def _header_messages(self):
        if self._messages is None:
            self._messages = self._file._reader.messages(self._header)
        return self._messages

    def _attributes(self):
        if self._attrs is None:
            self._file._require_open()
            self._attrs = dict(_decode_attribute(self._file._reader, body)
                               for mtype, _f, body in self._header_messages() if mtype == MSG_ATTRIBUTE)
        return self._attrs


class Group(_Node):
    def __init__(self, file, name, header=None):
        super(Group, self).__init__(file, name, header)
        self._children = {} if header is None else None

    def _links(self):
        if self._children is None:
            self._file._require_open()
            reader = self._file._reader
            for mtype, _f, body in self._header_messages():
                if mtype == MSG_SYMBOL_TABLE:
                    btree, heap = struct.unpack_from('<QQ', body, 0)
                    self._children = dict((n, None) for n in reader.group_links(btree, heap))
                    self._addresses = reader.group_links(btree, heap)
                    break
            else:
                raise IOError("'%s' is not an old-style group (no symbol table message)" % self.name)
        return self._children

    def _child_path(self, name):
        return (self.name.rstrip('/') + '/' + name) if self.name != '/' else '/' + name

    def _child(self, name):
        links = self._links()
        if name not in links:
            raise KeyError("Unable to open object (object '%s' doesn't exist)" % name)
        node = links[name]
        if node is None:
            header = self._addresses[name]
            kinds = set(m[0] for m in self._file._reader.messages(header))
            cls = Group if MSG_SYMBOL_TABLE in kinds else Dataset
            node = links[name] = cls(self._file, self._child_path(name), header)
        return node

    def __getitem__(self, path):
        self._file._require_open()
        node = self._file if path.startswith('/') else self
        for part in path.split('/'):
            if part:
                if not isinstance(node, Group):
                    raise KeyError("'%s' is not a group" % node.name)
                node = node._child(part)
        return node

    def get(self, path, default=None):
        try:
            return self[path]
        except KeyError:
            return default

    def __contains__(self, path):
        try:
            self[path]
            return True
        except KeyError:
            return False

    def keys(self):
        return sorted(self._links())           # h5py iterates old-style groups in name order

    def __iter__(self):
        return iter(self.keys())

    def __len__(self):
        return len(self._links())

    def values(self):
        return [self[name] for name in self.keys()]

    def items(self):
        return [(name, self[name]) for name in self.keys()]

    # ---- writing
    def _parent_for(self, path, create):
        node = self._file if path.startswith('/') else self
        parts = [p for p in path.split('/') if p]
        if not parts:
            raise ValueError('empty object name')
        for part in parts[:-1]:
            if part in node._links():
                node = node._child(part)
                if not isinstance(node, Group):
                    raise ValueError("'%s' is not a group" % node.name)
            elif create:
                node = node.create_group(part)
            else:
                raise KeyError(part)
        return node, parts[-1]

    def create_group(self, path):
        self._file._require_writable()
        parent, name = self._parent_for(path, create=True)
        if name in parent._links():
            raise ValueError('Unable to create group (name already exists)')
        group = Group(self._file, parent._child_path(name))
        parent._children[name] = group
        return group

    def create_dataset(self, path, shape=None, dtype=None, data=None, compression=None,
                       compression_opts=None, chunks=None, **unused):
        """The subset kPAL uses (reference kpal/klib.py:246-247): `data` given, optional
        ``compression='gzip'`` (level `compression_opts`, default 4) which implies chunking."""
        self._file._require_writable()
        if data is None:
            if shape is None:
                raise TypeError('one of data or shape is needed')
            data = np.zeros(shape, dtype=dtype or 'f4')
        data = np.ascontiguousarray(data, dtype=dtype)
        if shape is not None and tuple(np.atleast_1d(shape)) != data.shape:
            data = data.reshape(shape)
        if compression not in (None, 'gzip', True) and not isinstance(compression, int):
            raise ValueError("h5lite only has compression='gzip'")
        level = None
        if compression is not None:
            level = compression if isinstance(compression, int) and compression is not True else compression_opts
            level = GZIP_LEVEL if level is None else int(level)
        parent, name = self._parent_for(path, create=True)
        if name in parent._links():
            raise ValueError('Unable to create dataset (name already exists)')
        dataset = Dataset(self._file, parent._child_path(name))
        dataset._write(data, level, chunks)
        parent._children[name] = dataset
        return dataset

    def bulk_layout(self, rows, compression_opts=None):
        """``(chunk_elements, gzip_level)`` with which :meth:`create_datasets` stores the rows in
        bulk, or None when it would take the general path -- for callers that deflate the chunks
        themselves (``klib.save_profiles``: statistics and streams in one pass over the rows)."""
        if rows.ndim != 2 or rows.dtype.kind not in 'iuf' or not rows.shape[0] or not rows.shape[1]:
            return None
        m = rows.shape[1]
        chunk = _guess_chunk((m,), rows.dtype.itemsize)[0]
        if m % chunk or m // chunk > 2 * CHUNK_K:
            return None
        return chunk, GZIP_LEVEL if compression_opts is None else int(compression_opts)

    def create_datasets(self, names, rows, compression_opts=None, attrs=None, streams=None):
        """
        ``create_dataset(name, data=row, dtype=rows.dtype, compression='gzip')`` for every row
        of the C-contiguous 2-D array `rows`, plus ``dataset.attrs[key] = values[i]`` for every
        ``key: values`` of `attrs` (arrays of one int64 / float64 scalar per row) -- the bulk
        form behind ``klib.save_profiles``.  Same file content as the loop; the chunks of all
        rows are deflated by the native library on all host threads (``kpal_deflate_chunks``)
        and written with one call, the metadata is laid out in bulk on ``close()``.  `streams`
        = ``(blob, sizes)``: the chunks already deflated by the caller (layout: :meth:`bulk_layout`).
        (h5lite extension: h5py has no such call.)
        """
        self._file._require_writable()
        rows = np.ascontiguousarray(rows)
        if rows.ndim != 2 or rows.dtype.kind not in 'iuf' or len(names) != rows.shape[0]:
            raise ValueError('rows must be a 2-D numeric array with one row per name')
        attrs = dict((key, np.ascontiguousarray(values)) for key, values in (attrs or {}).items())
        level = GZIP_LEVEL if compression_opts is None else int(compression_opts)
        links = self._links()
        if len(set(names)) != len(names) or any(name in links or '/' in name for name in names):
            raise ValueError('Unable to create dataset (name already exists)')
        n, m = rows.shape
        chunk = _guess_chunk((m,), rows.dtype.itemsize)[0] if m else 0
        native = None
        try:
            from . import _cabi as native
            native.load()
        except Exception:
            native = None
        simple = all(v.shape == (n,) and v.dtype in (np.dtype('int64'), np.dtype('float64')) for v in attrs.values())
        if native is None or not n or not m or m % chunk or m // chunk > 2 * CHUNK_K or not simple:
            for i, name in enumerate(names):                    # the general path, one at a time
                dataset = self.create_dataset(name, data=rows[i], dtype=rows.dtype, compression='gzip',
                                              compression_opts=level)
                for key, values in attrs.items():
                    dataset.attrs[key] = values[i]
            return
        per = m // chunk
        if streams is not None:             # the zlib streams of the rows' chunks, back to back, and their sizes
            blob, sizes = streams
            if sizes.size != n * per or int(sizes.sum()) != blob.size:
                raise ValueError('streams do not match the rows')
        else:
            blob, sizes = native.deflate_chunks_packed(rows, chunk * rows.dtype.itemsize, level, sparse=True)
        self._file._drain(0)                                    # keep the file in creation order
        base = self._file._append(blob)
        ends = np.cumsum(sizes, dtype=np.uint64)
        addresses = (np.uint64(base) + ends - sizes.astype(np.uint64)).reshape(n, per)
        bulk = _Bulk((m,), rows.dtype, (chunk,), level, addresses, sizes.reshape(n, per), attrs)
        for i, name in enumerate(names):
            dataset = Dataset(self._file, self._child_path(name))
            dataset._bulk = (bulk, i)
            dataset._attrs = None                               # built from the bulk arrays on demand
            links[name] = dataset


class _Bulk(object):
    """The common part of the equally shaped, gzip-compressed 1-D datasets that one
    ``Group.create_datasets`` call writes (the profiles of a ``kpal count --by-record`` run:
    100 000 of them).  Their chunks are already in the file; addresses, stored sizes and
    attribute values are kept as arrays, and the serializer lays out all their chunk B-trees
    and object headers with a few array operations instead of a Python loop per dataset."""

    def __init__(self, shape, dtype, chunks, level, addresses, sizes, attrs):
        self.shape, self.dtype, self.chunks, self.level = shape, dtype, chunks, level
        self.addresses, self.sizes, self.attrs = addresses, sizes, attrs
        self.headers = None                 # object header addresses, set by the serializer

    def meta(self, i):
        per = self.addresses.shape[1]
        index = [((c * self.chunks[0],), int(self.addresses[i, c]), int(self.sizes[i, c]), 0) for c in range(per)]
        return {'shape': self.shape, 'dtype': self.dtype, 'filters': [(1, (self.level,))], 'chunks': self.chunks,
                'layout': ('chunked', None), 'index': index}


def _guess_chunk(shape, itemsize):
    """Chunk shape for a dataset whose chunking is left to the library: the rule h5py
    applies for compression='gzip' without `chunks` (so that files written here are
    chunked like the reference's): a target between 8 KiB and 1 MiB that grows with the
    dataset size, reached by halving the dimensions in turn."""
    chunks = [float(max(s, 1)) for s in shape]
    total = float(np.prod(chunks)) * itemsize
    target = 16384.0 * (2.0 ** np.log10(total / (1024.0 * 1024.0)))
    target = min(max(target, 8192.0), 1048576.0)
    index = 0
    while True:
        size = float(np.prod(chunks)) * itemsize
        if (size < target or abs(size - target) / target < 0.5) and size < 1048576.0:
            break
        if np.prod(chunks) == 1:
            break
        chunks[index % len(chunks)] = np.ceil(chunks[index % len(chunks)] / 2.0)
        index += 1
    return tuple(int(c) for c in chunks)


class Dataset(_Node):
    def __init__(self, file, name, header=None):
        super(Dataset, self).__init__(file, name, header)
        self._meta = None
        self._bulk = None                   # (_Bulk, row) for a dataset written by Group.create_datasets

    def _attributes(self):
        if self._attrs is None and self._bulk is not None:
            # somebody looks at (or is about to change) the attributes: from here on this is an
            # ordinary dataset, serialised on its own
            bulk, i = self._bulk
            self._meta = bulk.meta(i)
            self._attrs = dict((key, values[i]) for key, values in bulk.attrs.items())
            self._bulk = None
        return super(Dataset, self)._attributes()

    # ---- metadata
    def _load(self):
        if self._meta is not None:
            return self._meta
        if self._bulk is not None:
            self._meta = self._bulk[0].meta(self._bulk[1])
            return self._meta
        self._file._require_open()
        meta = {'filters': [], 'chunks': None, 'layout': None}
        for mtype, _f, body in self._header_messages():
            if mtype == MSG_DATASPACE:
                meta['shape'] = _decode_dataspace(body)
            elif mtype == MSG_DATATYPE:
                meta['dtype'] = _decode_datatype(body)[0]
            elif mtype == MSG_LAYOUT:
                version, cls = body[0], body[1]
                if version in (1, 2):                   # libhdf5 < 1.6.3 and files it keeps compatible
                    rank_l, cls = body[1], body[2]
                    at = 8
                    address = UNDEF
                    if cls != 0:
                        address = struct.unpack_from('<Q', body, at)[0]
                        at += 8
                    dims = struct.unpack_from('<%dI' % rank_l, body, at)
                    at += 4 * rank_l
                    if cls == 0:
                        size = struct.unpack_from('<I', body, at)[0]
                        meta['layout'] = ('compact', body[at + 4: at + 4 + size])
                    elif cls == 1:
                        meta['layout'] = ('contiguous-dims', address, dims)
                    elif cls == 2:
                        meta['layout'] = ('chunked', address)
                        meta['chunks'] = tuple(dims[:-1])
                    else:
                        raise IOError('unsupported data layout class %d' % cls)
                    continue
                if version != 3:
                    raise IOError('unsupported data layout message version %d' % version)
                if cls == 0:
                    size = struct.unpack_from('<H', body, 2)[0]
                    meta['layout'] = ('compact', body[4:4 + size])
                elif cls == 1:
                    meta['layout'] = ('contiguous',) + struct.unpack_from('<QQ', body, 2)
                elif cls == 2:
                    rank1 = body[2]
                    btree = struct.unpack_from('<Q', body, 3)[0]
                    dims = struct.unpack_from('<%dI' % rank1, body, 11)
                    meta['layout'] = ('chunked', btree)
                    meta['chunks'] = tuple(dims[:-1])
                else:
                    raise IOError('unsupported data layout class %d' % cls)
            elif mtype == MSG_FILTERS:
                version, count = body[0], body[1]
                at = 8 if version == 1 else 2
                for _ in range(count):
                    fid, name_len, _flags, n_values = struct.unpack_from('<HHHH', body, at)
                    if version == 2 and fid < 256:
                        n_values, name_len = struct.unpack_from('<H', body, at + 4)[0], 0
                        at += 6
                    else:
                        at += 8
                    at += _pad8(name_len) if version == 1 else name_len
                    values = struct.unpack_from('<%dI' % n_values, body, at)
                    at += 4 * n_values
                    if version == 1 and n_values % 2:
                        at += 4
                    meta['filters'].append((fid, values))
        if meta.get('shape') is None or 'dtype' not in meta or meta['layout'] is None:
            raise IOError("'%s' is not a readable dataset" % self.name)
        if meta['layout'][0] == 'contiguous-dims':      # old layout message: the size is the dataspace's
            _kind, address, _dims = meta['layout']
            meta['layout'] = ('contiguous', address, int(np.prod(meta['shape'])) * meta['dtype'].itemsize)
        self._meta = meta
        return meta

    @property
    def shape(self):
        return self._load()['shape']

    @property
    def dtype(self):
        return self._load()['dtype']

    @property
    def chunks(self):
        return self._load()['chunks']

    @property
    def compression(self):
        return 'gzip' if any(fid == 1 for fid, _ in self._load()['filters']) else None

    @property
    def size(self):
        return int(np.prod(self.shape))

    def __len__(self):
        return self.shape[0]

    # ---- reading
    def _chunk_index(self):
        """[(offsets, address, stored bytes, filter mask)] from the chunk B-tree."""
        meta = self._load()
        if 'index' not in meta and getattr(self, '_future', None) is not None:
            self._file._drain(0)                                # still being deflated in the background
        if self._meta is None:
            self._meta = meta                                   # (a bulk dataset: keep what was built)
        if 'index' in meta:
            return meta['index']
        reader = self._file._reader
        rank = len(meta['shape'])
        key_size = 8 + 8 * (rank + 1)
        out = []

        def walk(address):
            head = reader.read(address, 24)
            if head[:4] != b'TREE' or head[4] != 1:
                raise IOError('bad chunk B-tree node')
            level, used = head[5], struct.unpack_from('<H', head, 6)[0]
            body = reader.read(address + 24, used * (key_size + 8) + key_size)
            for i in range(used):
                at = i * (key_size + 8)
                nbytes, mask = struct.unpack_from('<II', body, at)
                offsets = struct.unpack_from('<%dQ' % rank, body, at + 8)
                child = struct.unpack_from('<Q', body, at + key_size)[0]
                if level:
                    walk(child)
                else:
                    out.append((offsets, child, nbytes, mask))
        if meta['layout'][1] != UNDEF:
            walk(meta['layout'][1])
        meta['index'] = out
        return out

    def _decode_chunk(self, raw, mask):
        meta = self._meta
        for position, (fid, values) in reversed(list(enumerate(meta['filters']))):
            if mask & (1 << position):
                continue
            if fid == 1:
                raw = zlib.decompress(raw)
            elif fid == 2:                          # shuffle
                width = values[0] if values else meta['dtype'].itemsize
                count = len(raw) // width
                raw = np.frombuffer(raw, np.uint8, count * width).reshape(width, count).T.tobytes() + raw[count * width:]
            elif fid == 3:                          # fletcher32: checksum trails the data
                raw = raw[:-4]
            else:
                raise IOError('unsupported filter %d' % fid)
        return raw

    def _read_all(self):
        meta = self._load()
        reader = self._file._reader
        shape, dtype = meta['shape'], meta['dtype']
        if isinstance(dtype, _VlenString):
            raise IOError('variable-length datasets are not supported')
        kind = meta['layout'][0]
        if kind == 'compact':
            return np.frombuffer(meta['layout'][1], dtype, int(np.prod(shape))).reshape(shape).copy()
        if kind == 'contiguous':
            _k, address, size = meta['layout']
            if address == UNDEF:
                return np.zeros(shape, dtype)
            return np.frombuffer(reader.read(address, size), dtype, int(np.prod(shape))).reshape(shape).copy()
        out = np.zeros(shape, dtype)
        self._read_chunks_into(out)
        return out

    def _read_chunks_into(self, out):
        """Chunked layout: every chunk is read, decoded and placed by a pool task of its own
        (positional reads and zlib both run without the GIL)."""
        meta = self._meta
        reader = self._file._reader
        shape, dtype, chunk = meta['shape'], meta['dtype'], meta['chunks']
        count = int(np.prod(chunk))

        def place(entry):
            offsets, address, nbytes, mask = entry
            raw = self._decode_chunk(reader.read(address, nbytes), mask)
            block = np.frombuffer(raw, dtype, count).reshape(chunk)
            where = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offsets, chunk, shape))
            out[where] = block[tuple(slice(0, w.stop - w.start) for w in where)]
        index = self._chunk_index()
        if len(index) > 1:
            list(_workers().map(place, index))
        else:
            for entry in index:
                place(entry)

    def __getitem__(self, key):
        self._file._require_open()
        data = self._read_all()
        if key is Ellipsis or key == () or (isinstance(key, slice) and key == slice(None)):
            return data
        return data[key]

    def read_direct(self, dest):
        """Read the whole dataset into the array `dest` (same shape)."""
        self._file._require_open()
        meta = self._load()
        if (meta['layout'][0] == 'chunked' and isinstance(dest, np.ndarray) and dest.shape == meta['shape']
                and dest.dtype == meta['dtype']):
            n_chunks = 1
            for s_, c_ in zip(meta['shape'], meta['chunks']):
                n_chunks *= -(-s_ // c_)
            if len(self._chunk_index()) < n_chunks:
                dest[...] = 0                            # chunks that were never written read as the fill value
            self._read_chunks_into(dest)                 # decoded straight into the caller's array
        else:
            dest[...] = self._read_all()

    def __array__(self, dtype=None, copy=None):
        data = self._read_all()
        return data if dtype is None else data.astype(dtype)

    # ---- writing (called once, by Group.create_dataset)
    def _write(self, data, gzip_level, chunks):
        file = self._file
        meta = {'shape': data.shape, 'dtype': data.dtype, 'filters': [], 'chunks': None}
        if data.dtype.kind not in 'iuf':
            raise TypeError('h5lite cannot store datasets of dtype %r' % (data.dtype,))
        if gzip_level is None and chunks is None or data.ndim == 0:
            address = file._append(data.tobytes()) if data.size else UNDEF
            meta['layout'] = ('contiguous', address, data.nbytes)
        else:
            if chunks is None or chunks is True:
                chunks = _guess_chunk(data.shape, data.dtype.itemsize)
            chunks = tuple(int(c) for c in np.atleast_1d(chunks))
            if len(chunks) != data.ndim or any(c < 1 for c in chunks):
                raise ValueError('bad chunk shape %r' % (chunks,))
            meta['chunks'] = chunks
            if gzip_level is not None:
                meta['filters'] = [(1, (gzip_level,))]
            grid = [range(0, s, c) for s, c in zip(data.shape, chunks)]
            origins = [()]
            for axis in grid:
                origins = [o + (x,) for o in origins for x in axis]

            def encode(origin):
                where = tuple(slice(o, min(o + c, s)) for o, c, s in zip(origin, chunks, data.shape))
                block = data[where]
                if block.shape != chunks:                       # edge chunks are stored whole
                    whole = np.zeros(chunks, data.dtype)
                    whole[tuple(slice(0, w.stop - w.start) for w in where)] = block
                    block = whole
                raw = block.tobytes()
                return zlib.compress(raw, gzip_level) if gzip_level is not None else raw
            meta['layout'] = ('chunked', None)
            if data.nbytes <= _DEFERRED_BYTES and gzip_level is not None:
                # A small dataset (one profile of a --by-record run: 100 k of them) is deflated
                # as ONE background task on a private copy, and its chunks reach the file when
                # the task is done (File._drain): the caller goes on to the next profile, and
                # the parallelism is across datasets instead of inside one.
                data = np.array(data, copy=True)
                self._future = _workers().submit(lambda: [encode(o) for o in origins])
                self._origins = origins
                file._pending.append(self)
                self._meta = meta
                file._drain(_MAX_PENDING)
                return
            index = []
            file._drain(0)                                       # keep the file in creation order
            for origin, raw in zip(origins, _workers().map(encode, origins)):
                index.append((origin, file._append(raw), len(raw), 0))
            meta['index'] = index
        self._meta = meta


class File(Group):
    """``h5lite.File(path, mode)`` with mode 'r' or 'w' (``h5py.File`` look-alike)."""

    def __init__(self, name, mode='r'):
        if mode not in ('r', 'w', 'w-', 'x'):
            raise ValueError("h5lite supports the modes 'r' and 'w' only, not %r" % (mode,))
        self.filename = name
        self.mode = 'r' if mode == 'r' else 'r+'
        self._closed = False
        if mode == 'r':
            self._handle = open(name, 'rb')
            try:
                self._reader = _Reader(self._handle)
            except Exception:
                self._handle.close()
                raise
            Group.__init__(self, self, '/', self._reader.root_header)
            self._writable = False
        else:
            if mode in ('w-', 'x') and os.path.exists(name):
                raise IOError("Unable to create file (file exists): '%s'" % name)
            self._handle = open(name, 'w+b')
            self._handle.write(b'\0' * 96)              # the superblock goes here on close()
            self._end = 96
            self._pending = []                          # datasets whose chunks are still being deflated
            self._reader = _WriteSideReader(self)
            Group.__init__(self, self, '/')
            self._writable = True
            # an unclosed file is completed when it is collected (__del__) or, at the latest,
            # when the interpreter exits
            self._at_exit = _close_at_exit(weakref.ref(self))
            atexit.register(self._at_exit)

    # ---- state
    def _require_open(self):
        if self._closed:
            raise ValueError('Invalid file handle (the file is closed)')

    def _require_writable(self):
        self._require_open()
        if not self._writable:
            raise ValueError('the file is open read-only')

    def __bool__(self):
        return not self._closed

    def __enter__(self):
        return self

    def __exit__(self, *unused):
        self.close()

    def __repr__(self):
        return '<h5lite file "%s" (mode %s)>' % (os.path.basename(str(self.filename)),
                                                  self.mode) if not self._closed else '<Closed h5lite file>'

    def _append(self, raw):
        """Raw bytes to the end of the file (8-byte aligned) -> their address."""
        address = _pad8(self._end)
        self._handle.seek(address)
        self._handle.write(raw)
        self._end = address + len(raw)
        return address

    def _drain(self, keep):
        """Appends the chunks of pending datasets (oldest first) until at most `keep` are left."""
        while len(self._pending) > keep:
            dataset = self._pending.pop(0)
            chunks = dataset._future.result()
            dataset._meta['index'] = [(origin, self._append(raw), len(raw), 0)
                                      for origin, raw in zip(dataset._origins, chunks)]
            dataset._future = dataset._origins = None

    def flush(self):
        """Pushes the data written so far to the OS.  Chunks still being deflated in the
        background follow as they finish; the metadata (object headers, group B-trees,
        superblock) is written by close()."""
        self._require_open()
        if self._writable:
            self._handle.flush()

    def __del__(self):
        try:
            if not getattr(self, '_closed', True):
                self.close()
        except Exception:
            pass

    def close(self):
        if self._closed:
            return
        try:
            if self._writable:
                self._drain(0)
                _Serializer(self).run()
        finally:
            self._closed = True
            self._handle.close()
            if self._writable:
                atexit.unregister(self._at_exit)


def _close_at_exit(reference):
    def run():
        file = reference()
        if file is not None:
            file.close()
    return run


class _WriteSideReader(object):
    """Lets datasets of a file that is still being written read their chunks back."""

    def __init__(self, file):
        self._file = file

    def read(self, address, size):
        handle = self._file._handle
        handle.flush()
        return os.pread(handle.fileno(), size, address)


# =========================================================================
# writing the metadata
# =========================================================================
def _message(mtype, body, flags=0):
    body = body + b'\0' * (_pad8(len(body)) - len(body))
    return struct.pack('<HHB3x', mtype, len(body), flags) + body


def _object_header(messages):
    """Version-1 object header holding `messages` (already encoded) in one block."""
    data = b''.join(messages)
    return struct.pack('<BBHII4x', 1, 0, len(messages), 1, len(data)) + data


def _dataspace_message(shape):
    body = struct.pack('<BBB5x', 1, len(shape), 0)
    return body + b''.join(struct.pack('<Q', s) for s in shape)


class _Serializer(object):
    """Lays out and writes everything that is not raw chunk data: one global heap collection
    for the variable-length strings, then bottom-up every object (datasets, then the groups
    that link to them), finally the superblock."""

    def __init__(self, file):
        self.file = file
        self.heap_objects = []              # global heap payloads, index = position + 1
        self.heap_address = None

    # ---- global heap
    def _vlen_reference(self, text):
        data = text.encode('utf-8')
        self.heap_objects.append(data)
        return len(data), len(self.heap_objects)

    def _collect_strings(self, node):
        if isinstance(node, Dataset) and node._bulk is not None:
            return                                      # written in bulk: numeric attributes only
        for value in node._attributes().values():
            if isinstance(value, str):
                self.n_strings += 1
                self.string_bytes += 16 + _pad8(len(value.encode('utf-8')))
            elif isinstance(value, np.ndarray) and value.dtype == object:
                raise TypeError('arrays of Python objects cannot be stored')
        if isinstance(node, Group):
            for child in node._links().values():
                self._collect_strings(child)

    def _write_global_heap(self):
        size = 16
        for data in self.heap_objects:
            size += 16 + _pad8(len(data))
        size = max(4096, _pad8(size + 16))              # libhdf5's minimum collection size
        out = bytearray(struct.pack('<4sB3xQ', b'GCOL', 1, size))
        for index, data in enumerate(self.heap_objects, 1):
            out += struct.pack('<HHIQ', index, 1, 0, len(data))
            out += data + b'\0' * (_pad8(len(data)) - len(data))
        free = size - len(out)
        out += struct.pack('<HHIQ', 0, 0, 0, free)      # object 0: the free space (size includes this header)
        out += b'\0' * (size - len(out))
        self.file._handle.seek(self.heap_address)
        self.file._handle.write(bytes(out))

    # ---- attributes
    def _attribute_message(self, name, value):
        if isinstance(value, str):
            dtype, shape = _VlenString(utf8=True), ()
            length, index = self._vlen_reference(value)
            raw = struct.pack('<IQI', length, self.heap_address, index)
        else:
            array = np.asarray(value)
            dtype, shape = array.dtype, array.shape
            raw = np.ascontiguousarray(array).tobytes()
        encoded_name = name.encode('utf-8') + b'\0'
        dt = _encode_datatype(dtype)
        ds = _dataspace_message(shape)
        body = struct.pack('<BBHHH', 1, 0, len(encoded_name), len(dt), len(ds))
        for part in (encoded_name, dt, ds):
            body += part + b'\0' * (_pad8(len(part)) - len(part))
        body += raw
        if len(body) > 65528:
            raise ValueError("attribute '%s' is too large for an object header message" % name)
        return _message(MSG_ATTRIBUTE, body)

    # ---- datasets
    def _chunk_btree(self, dataset):
        """Writes the version-1 B-tree over the dataset's chunks -> address of its root."""
        meta = dataset._meta
        rank = len(meta['shape'])
        file = self.file
        key_size = 8 + 8 * (rank + 1)
        node_size = 24 + (2 * CHUNK_K + 1) * key_size + 2 * CHUNK_K * 8

        def key(nbytes, mask, offsets):
            return struct.pack('<II', nbytes, mask) + struct.pack('<%dQ' % (rank + 1), *(tuple(offsets) + (0,)))
        # one past the last chunk, in the row-major order of the chunk offsets
        chunks = meta['chunks']
        end = tuple(((s + c - 1) // c) * c if axis == 0 else 0
                    for axis, (s, c) in enumerate(zip(meta['shape'], chunks)))
        entries = [(key(nbytes, mask, offsets), address) for offsets, address, nbytes, mask in meta['index']]
        if not entries:
            return UNDEF
        level = 0
        while True:
            groups = [entries[i:i + 2 * CHUNK_K] for i in range(0, len(entries), 2 * CHUNK_K)]
            addresses = []
            base = _pad8(file._end)
            for g in range(len(groups)):
                addresses.append(base + g * node_size)
            parents = []
            for g, group in enumerate(groups):
                left = addresses[g - 1] if g else UNDEF
                right = addresses[g + 1] if g + 1 < len(groups) else UNDEF
                node = bytearray(struct.pack('<4sBBHQQ', b'TREE', 1, level, len(group), left, right))
                for k, child in group:
                    node += k + struct.pack('<Q', child)
                # the key after the last child: the first key of the right sibling, or the end
                node += groups[g + 1][0][0] if g + 1 < len(groups) else key(0, 0, end)
                node += b'\0' * (node_size - len(node))
                file._append(bytes(node))
                parents.append((group[0][0], addresses[g]))
            if len(groups) == 1:
                return addresses[0]
            entries = parents
            level += 1

    def _write_bulk(self, bulk):
        """Chunk B-trees and object headers of all datasets of a Group.create_datasets call:
        the same bytes _write_dataset produces, built as two arrays (one leaf node and one
        header per dataset) from a template whose variable fields are patched column-wise."""
        file = self.file
        n, per = bulk.addresses.shape
        # ---- one leaf node per dataset (per <= 2 * CHUNK_K chunks, rank 1)
        key = [('nbytes', '<u4'), ('mask', '<u4'), ('off0', '<u8'), ('off1', '<u8')]
        used = 24 + per * 32 + 24
        node_size = 24 + (2 * CHUNK_K + 1) * 24 + 2 * CHUNK_K * 8
        node = np.dtype([('sig', 'S4'), ('type', 'u1'), ('level', 'u1'), ('used', '<u2'), ('left', '<u8'),
                         ('right', '<u8'), ('entries', key + [('child', '<u8')], (per,)), ('last', key),
                         ('pad', 'u1', (node_size - used,))])
        assert node.itemsize == node_size
        nodes = np.zeros(n, dtype=node)
        nodes['sig'], nodes['type'], nodes['used'] = b'TREE', 1, per
        nodes['left'] = nodes['right'] = UNDEF
        nodes['entries']['nbytes'] = bulk.sizes
        nodes['entries']['off0'] = (np.arange(per, dtype=np.uint64) * np.uint64(bulk.chunks[0]))[None, :]
        nodes['entries']['child'] = bulk.addresses
        nodes['last']['off0'] = -(-bulk.shape[0] // bulk.chunks[0]) * bulk.chunks[0]
        base = file._append(nodes.tobytes())
        btrees = np.uint64(base) + np.arange(n, dtype=np.uint64) * np.uint64(node_size)
        # ---- one object header per dataset: a template with marked fields, then patched copies
        marks = {}

        def mark(label, dtype):
            value = np.array([0x7E57A77B00000000 + len(marks)], dtype='<u8')
            marks[label] = value.tobytes()
            return value.view(dtype)[0]
        template = Dataset(file, '/template')
        template._meta = dict(bulk.meta(0), index=None)
        template._attrs = dict((name, mark(('attr', name), values.dtype)) for name, values in bulk.attrs.items())
        header = bytearray(self._dataset_header(template, int(mark('btree', '<u8'))))
        offsets = {}
        for label, pattern in marks.items():
            at = bytes(header).find(pattern)
            if at < 0 or bytes(header).find(pattern, at + 1) >= 0:
                raise AssertionError('h5lite: cannot locate a field of the bulk object header')
            offsets[label] = at
        headers = np.tile(np.frombuffer(bytes(header), dtype=np.uint8), (n, 1))
        headers[:, offsets['btree']:offsets['btree'] + 8] = btrees.astype('<u8').view(np.uint8).reshape(n, 8)
        for name, values in bulk.attrs.items():
            at = offsets[('attr', name)]
            headers[:, at:at + 8] = np.ascontiguousarray(values).view(np.uint8).reshape(n, 8)
        base = file._append(headers.tobytes())
        bulk.headers = np.uint64(base) + np.arange(n, dtype=np.uint64) * np.uint64(len(header))

    def _write_dataset(self, dataset):
        meta = dataset._load()
        btree = self._chunk_btree(dataset) if meta['layout'][0] == 'chunked' else None
        return self.file._append(self._dataset_header(dataset, btree))

    def _dataset_header(self, dataset, btree):
        """The object header of `dataset` (bytes) whose chunk B-tree, if any, is at `btree`."""
        meta = dataset._load()
        messages = [_message(MSG_DATASPACE, _dataspace_message(meta['shape']), 0),
                    _message(MSG_DATATYPE, _encode_datatype(meta['dtype']), 1)]
        if meta['layout'][0] == 'chunked':
            rank = len(meta['shape'])
            # fill value v2: allocation incremental, written if set, no user-defined value
            messages.append(_message(MSG_FILL, struct.pack('<BBBB', 2, 3, 2, 0), 1))
            layout = struct.pack('<BBBQ', 3, 2, rank + 1, btree)
            layout += struct.pack('<%dI' % (rank + 1), *(meta['chunks'] + (meta['dtype'].itemsize,)))
            if meta['filters']:
                body = struct.pack('<BB6x', 1, len(meta['filters']))
                for fid, values in meta['filters']:
                    name = b'deflate\0'
                    body += struct.pack('<HHHH', fid, len(name), 1, len(values)) + name
                    body += struct.pack('<%dI' % len(values), *values)
                    if len(values) % 2:
                        body += b'\0' * 4
                messages.append(_message(MSG_FILTERS, body, 1))
            messages.append(_message(MSG_LAYOUT, layout))
        else:
            _kind, address, size = meta['layout']
            messages.append(_message(MSG_FILL, struct.pack('<BBBB', 2, 2, 2, 0), 1))
            messages.append(_message(MSG_LAYOUT, struct.pack('<BBQQ', 3, 1, address, size)))
        for name, value in dataset._attributes().items():
            messages.append(self._attribute_message(name, value))
        return _object_header(messages)

    # ---- groups
    def _write_group(self, group):
        """-> (object header address, B-tree address, local heap address)."""
        file = self.file
        children = []
        for child in group._links().values():                   # datasets written in bulk: all at once
            if isinstance(child, Dataset) and child._bulk is not None and child._bulk[0].headers is None:
                self._write_bulk(child._bulk[0])
        for name in sorted(group._links(), key=lambda n: n.encode('utf-8')):
            child = group._links()[name]
            if isinstance(child, Group):
                header, btree, heap = self._write_group(child)
                children.append((name, header, 1, btree, heap))
            elif child._bulk is not None:
                children.append((name, int(child._bulk[0].headers[child._bulk[1]]), 0, 0, 0))
            else:
                children.append((name, self._write_dataset(child), 0, 0, 0))
        # local heap: "" at offset 0, then the names, each padded to 8 bytes
        heap_data = bytearray(8)
        offsets = {}
        for name, *_rest in children:
            offsets[name] = len(heap_data)
            raw = name.encode('utf-8') + b'\0'
            heap_data += raw + b'\0' * (_pad8(len(raw)) - len(raw))
        heap_address = file._append(b'')
        data_address = heap_address + 32
        file._append(struct.pack('<4sB3xQQQ', b'HEAP', 0, len(heap_data), 1, data_address) + bytes(heap_data))
        # symbol table nodes of up to 2 * LEAF_K entries, then the B-tree levels over them
        snod_size = 8 + 2 * LEAF_K * 40
        entries = []                                    # (key = heap offset of the largest name below, address)
        for i in range(0, len(children), 2 * LEAF_K):
            part = children[i:i + 2 * LEAF_K]
            node = bytearray(struct.pack('<4sBBH', b'SNOD', 1, 0, len(part)))
            for name, header, cache, btree, heap in part:
                node += struct.pack('<QQII', offsets[name], header, cache, 0)
                node += struct.pack('<QQ', btree, heap) if cache == 1 else b'\0' * 16
            node += b'\0' * (snod_size - len(node))
            entries.append((offsets[part[-1][0]], file._append(bytes(node))))
        node_size = 24 + (2 * GROUP_K + 1) * 8 + 2 * GROUP_K * 8
        level = 0
        if not entries:                                 # empty group: a B-tree root without children
            node = struct.pack('<4sBBHQQ', b'TREE', 0, 0, 0, UNDEF, UNDEF) + struct.pack('<Q', 0)
            btree_address = file._append(node + b'\0' * (node_size - len(node)))
        while entries:
            groups = [entries[i:i + 2 * GROUP_K] for i in range(0, len(entries), 2 * GROUP_K)]
            base = _pad8(file._end)
            addresses = [base + g * node_size for g in range(len(groups))]
            parents = []
            for g, part in enumerate(groups):
                left = addresses[g - 1] if g else UNDEF
                right = addresses[g + 1] if g + 1 < len(groups) else UNDEF
                node = bytearray(struct.pack('<4sBBHQQ', b'TREE', 0, level, len(part), left, right))
                # key 0: the largest name to the left of this node ("" for the leftmost)
                node += struct.pack('<Q', groups[g - 1][-1][0] if g else 0)
                for k, child in part:
                    node += struct.pack('<QQ', child, k)
                node += b'\0' * (node_size - len(node))
                file._append(bytes(node))
                parents.append((part[-1][0], addresses[g]))
            if len(groups) == 1:
                btree_address = addresses[0]
                break
            entries = parents
            level += 1
        messages = [_message(MSG_SYMBOL_TABLE, struct.pack('<QQ', btree_address, heap_address))]
        for name, value in group._attributes().items():
            messages.append(self._attribute_message(name, value))
        return file._append(_object_header(messages)), btree_address, heap_address

    def run(self):
        file = self.file
        self.n_strings, self.string_bytes = 0, 0
        self._collect_strings(file)
        if self.n_strings:
            # reserve the collection now (attribute messages need its address), fill it in last
            size = max(4096, _pad8(16 + self.string_bytes + 16))
            self.heap_address = file._append(b'\0' * size)
        header, btree, heap = self._write_group(file)
        if self.n_strings:
            self._write_global_heap()
        end = _pad8(file._end)
        file._handle.seek(0, os.SEEK_END)
        if file._handle.tell() < end:
            file._handle.write(b'\0' * (end - file._handle.tell()))
        superblock = SIGNATURE + struct.pack('<BBBBBBBBHHI', 0, 0, 0, 0, 0, 8, 8, 0, LEAF_K, GROUP_K, 0)
        superblock += struct.pack('<QQQQ', 0, UNDEF, end, UNDEF)
        superblock += struct.pack('<QQII', 0, header, 1, 0) + struct.pack('<QQ', btree, heap)
        assert len(superblock) == 96
        file._handle.seek(0)
        file._handle.write(superblock)
        file._handle.flush()
