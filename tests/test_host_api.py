"""
CPU tests of the host-side mirror of the reference interface (everything in
kpal_b200.klib / kdistlib / metrics / kmer that does not need the GPU): names,
container behaviour, the host-path options of ProfileDistance, HDF5 save/load
through a handle double, and the host metrics -- compared with the reference
itself when its tree is present.  Mirrors reference tests/test_klib.py:101-304,
tests/test_metrics.py and tests/test_kdistlib.py:22-102.
"""
import io
import itertools
from collections import Counter

import numpy as np
import pytest

from conftest import dense
from kpal_b200 import kdistlib, klib, metrics
from oracle import kpal_oracle as ko, ref_loader


class FakeDataset(object):
    def __init__(self, data):
        self.data = np.array(data, dtype='int64')
        self.attrs = {}

    def __getitem__(self, item):
        return np.array(self.data[item])        # h5py hands out copies, never views

    @property
    def shape(self):
        return self.data.shape

    def read_direct(self, dest):
        dest[...] = self.data


class FakeGroup(dict):
    pass


class FakeH5(object):
    """Just enough of h5py.File for Profile.save / from_file (duck typing as in
    reference kpal/klib.py:74-75,243-254)."""
    def __init__(self):
        self.groups = {'profiles': FakeGroup()}
        self.attrs = {}
        self.flushed = 0

    def __getitem__(self, path):
        node = self.groups
        for part in path.strip('/').split('/'):
            node = node[part]
        return node

    def create_dataset(self, path, data=None, dtype=None, compression=None):
        assert dtype == 'int64' and compression == 'gzip'
        group, name = path.rsplit('/', 1)
        dataset = FakeDataset(data)
        self[group][name] = dataset
        return dataset

    def flush(self):
        self.flushed += 1


def as_array(counts, k):
    return np.array([counts[''.join(s)] for s in itertools.product('ACGT', repeat=k)])


def test_profile_container(golden):
    counts = ko.count_sequences(golden["fixtures"]["LENGTH_60"], 8)
    profile = klib.Profile(counts, name='abc')
    assert profile.length == 8 and profile.number == 4 ** 8
    assert profile.total == counts.sum() and profile.non_zero == np.count_nonzero(counts)
    assert profile.mean == counts.mean() and profile.std == counts.std()
    assert profile.median == np.median(counts)
    copy = profile.copy()
    copy.counts[0] += 5
    assert profile.counts[0] == counts[0] and copy.name == 'abc'
    assert klib.Profile(np.zeros(20)).length == 2          # float-log behaviour kept (klib.py:59)


def test_profile_name_rules():
    counts = np.zeros(16, dtype=np.int64)
    for bad in ('abc/def', 'a.b'):
        with pytest.raises(ValueError):
            klib.Profile(counts, name=bad)
        with pytest.raises(ValueError):
            klib.Profile(counts).save(FakeH5(), name=bad)
    assert klib.Profile(counts, name=None).name is None


def test_profile_save_and_from_file(golden):
    counts = ko.count_sequences(golden["fixtures"]["LENGTH_60"], 4)
    handle = FakeH5()
    assert klib.Profile(counts).save(handle) == '1'          # first free number
    assert klib.Profile(counts).save(handle) == '2'
    assert klib.Profile(counts, name='x').save(handle) == 'x'
    assert klib.Profile(counts, name='x').save(handle, name='y') == 'y'
    dataset = handle['profiles/x']
    assert dataset.attrs['length'] == 4 and dataset.attrs['total'] == counts.sum()
    assert dataset.attrs['non_zero'] == np.count_nonzero(counts)
    assert set(dataset.attrs) == {'length', 'total', 'non_zero', 'mean', 'median', 'std'}
    loaded = klib.Profile.from_file(handle)
    assert loaded.name == '1' and np.array_equal(loaded.counts, counts)
    assert klib.Profile.from_file(handle, name='y').name == 'y'
    assert handle.flushed == 4


def test_profile_from_file_old_format(golden):
    counts = ko.count_sequences(golden["fixtures"]["LENGTH_60"], 4)
    text = '%d\n%d\n%d\n' % (4, counts.sum(), np.count_nonzero(counts))
    text += '\n'.join(str(c) for c in counts) + '\n'
    profile = klib.Profile.from_file_old_format(io.StringIO(text), name='old')
    assert np.array_equal(profile.counts, counts) and profile.length == 4


def test_host_profile_operations(golden):
    left = ko.count_sequences(golden["fixtures"]["LENGTH_60"], 8)
    right = ko.count_sequences(golden["fixtures"]["LENGTH_60_MORE"], 8)
    p = klib.Profile(left.copy())
    p.merge(klib.Profile(right))
    assert np.array_equal(p.counts, left + right)
    p = klib.Profile(left.copy())
    p.merge(klib.Profile(right), merger=metrics.mergers['xor'])
    assert np.array_equal(p.counts, (left + right) * np.logical_xor(left, right))
    # shrink: reference test_klib.py:235-279
    p = klib.Profile(left.copy())
    p.shrink(1)
    assert p.length == 7 and np.array_equal(p.counts, left.reshape(-1, 4).sum(axis=1))
    p.shrink(3)
    assert p.length == 4 and p.counts.sum() == left.sum()
    with pytest.raises(ValueError):
        klib.Profile(np.zeros(16, dtype=np.int64)).shrink(2)
    # dna / binary / reverse complement: reference test_klib.py:219-233,292-304
    p = klib.Profile(np.zeros(4 ** 4, dtype=np.int64))
    assert p.dna_to_binary('ACGT') == 0b00011011 and p.binary_to_dna(0b00011011) == 'ACGT'
    assert p.reverse_complement(p.dna_to_binary('AACG')) == p.dna_to_binary('CGTT')
    for entry in golden["reverse_complement"]:
        q = klib.Profile(np.zeros(4 ** entry["k"], dtype=np.int64))
        assert [q.reverse_complement(i) for i in range(q.number)] == entry["table"]
        assert q._rc_table().tolist() == entry["table"]
    # shuffle keeps the multiset of counts (seeded like reference test_klib.py:281-290)
    p = klib.Profile(left.copy())
    np.random.seed(100)
    p.shuffle()
    assert sorted(p.counts) == sorted(left) and not np.array_equal(p.counts, left)


def test_metrics_host_functions():
    rng = np.random.default_rng(0)
    a, b = rng.integers(0, 21, 100), rng.integers(0, 21, 100)
    assert metrics.distribution(a) == sorted(Counter(a).items())
    assert metrics.vector_length(a) == pytest.approx(np.sqrt(float(np.sum(a * a))))
    sa, sb = metrics.get_scale(a, b)
    assert (sa, sb) == ((b.sum() / a.sum(), 1.0) if a.sum() < b.sum() else (1.0, a.sum() / b.sum()))
    assert metrics.scale_down(1.0, 1.5) == (1.0 / 1.5, 1.0)
    assert np.array_equal(metrics.positive(a, b), [i if j else 0 for i, j in zip(a, b)])
    for name in ('prod', 'sum'):
        f = metrics.pairwise[name]
        values = [f(float(i), float(j)) for i, j in zip(a, b) if i or j]
        assert metrics.multiset(a, b, f) == pytest.approx(sum(values) / (len(values) + 1))
        assert metrics.multiset(a, b, f) == ko.multiset(a, b, ko.PAIRWISE[name])
    assert metrics.euclidean(a, b) == ko.euclidean(a, b)
    assert metrics.cosine_similarity(a, b) == ko.cosine_similarity(a, b)
    assert set(metrics.vector_distance) == {'default', 'euclidean', 'cosine'}
    assert set(metrics.summary) == {'min', 'average', 'median'}
    assert set(metrics.mergers) == {'sum', 'xor', 'int', 'nint'}


def test_dynamic_smooth_golden():
    """reference tests/test_kdistlib.py:76-102"""
    a = Counter(['AC', 'AG', 'AT', 'CA', 'CC', 'CG', 'CT', 'GA', 'GC', 'GG', 'GT', 'TA', 'TG', 'TT'])
    b = Counter(['AC', 'AT', 'CA', 'CC', 'CG', 'CT', 'GA', 'GC', 'GG', 'GT', 'TA', 'TC', 'TG', 'TT'])
    pa, pb = klib.Profile(as_array(a, 2)), klib.Profile(as_array(b, 2))
    kdistlib.ProfileDistance().dynamic_smooth(pa, pb)
    ea = Counter(['AA', 'AA', 'AA', 'CA', 'CC', 'CG', 'CT', 'GA', 'GC', 'GG', 'GT', 'TA', 'TA', 'TA'])
    eb = Counter(['AA', 'AA', 'CA', 'CC', 'CG', 'CT', 'GA', 'GC', 'GG', 'GT', 'TA', 'TA', 'TA', 'TA'])
    assert np.array_equal(pa.counts, as_array(ea, 2)) and np.array_equal(pb.counts, as_array(eb, 2))
    v = np.random.default_rng(1).integers(0, 21, 100)
    got = kdistlib.ProfileDistance()._collapse(v, 30, 40)
    assert got.tolist() == [v[30 + 10 * i:40 + 10 * i].sum() for i in range(4)]


def test_gpu_dispatch_rules():
    """Which options go to the device fast path."""
    PD = kdistlib.ProfileDistance
    assert PD()._gpu_options() == dict(metric='multiset', pairwise='prod', do_balance=False,
                                       do_scale=False, down=False)
    assert PD(pairwise=metrics.pairwise['sum'], do_scale=True, down=True)._gpu_options()['pairwise'] == 'sum'
    assert PD(distance_function=metrics.euclidean)._gpu_options()['metric'] == 'euclidean'
    assert PD(distance_function=metrics.vector_distance['cosine'])._gpu_options()['metric'] == 'cosine'
    assert PD(do_positive=True, do_scale=True)._gpu_options() == dict(
        metric='multiset', pairwise='prod', do_balance=False, do_scale=True, down=False,
        do_positive=True)                      # device, pair by pair (kpal_pair_distance_positive)
    assert PD(do_smooth=True)._gpu_options() is None
    assert PD(do_positive=True, do_smooth=True)._gpu_options() is None
    assert PD(pairwise=lambda x, y: abs(x - y))._gpu_options() is None
    assert PD(distance_function=lambda x, y: 0.0)._gpu_options() is None


def test_host_path_distance_and_matrix_text(golden):
    """Smoothing / custom callables run the host pipeline: works without a
    GPU and equals the oracle's reading of the reference."""
    left = klib.Profile(ko.count_sequences(golden["fixtures"]["LENGTH_60"], 4), 'a')
    right = klib.Profile(ko.count_sequences(golden["fixtures"]["LENGTH_60_MORE"], 4), 'b')
    third = klib.Profile(left.counts.copy(), None)
    dist = kdistlib.ProfileDistance(do_positive=True, do_scale=True, do_smooth=True, threshold=-1)
    mask = (left.counts != 0) & (right.counts != 0)           # threshold -1: smoothing is a no-op
    want = ko.distance(left.counts * mask, right.counts * mask, do_scale=True)
    assert dist.distance(left, right) == want
    assert ko.distance(left.counts, right.counts, do_scale=True, do_positive=True) == want
    out = io.StringIO()
    custom = kdistlib.ProfileDistance(pairwise=lambda x, y: abs(x - y) / (x + y + 1))
    kdistlib.distance_matrix([left, right, third], out, 3, custom)
    d = ko.distance(left.counts, right.counts, pairwise="sum")
    assert out.getvalue() == '3\na\nb\nNone\n%.3f\n%.3f %.3f\n' % (d, 0.0, d)
    out = io.StringIO()
    kdistlib.distance_matrix([left], out, 2, kdistlib.ProfileDistance())
    assert out.getvalue() == '1\na\n'                     # reference test_kdistlib.py:38-47


@pytest.mark.skipif(not ref_loader.available(), reason="reference tree not present")
def test_host_path_equals_reference_randomised():
    rklib, rkdist, rmetrics = ref_loader.load()
    rng = np.random.default_rng(8)
    for _ in range(20):
        k = int(rng.integers(2, 5))
        l, r = rng.poisson(1.5, 4 ** k), rng.poisson(2.5, 4 ** k)
        for opts in (dict(do_smooth=True, threshold=1),
                     dict(do_smooth=True, do_scale=True, down=True),
                     dict(do_positive=True, do_smooth=True, do_scale=True)):
            for summary in ('min', 'average', 'median'):
                ref = rkdist.ProfileDistance(summary=rmetrics.summary[summary], **opts).distance(
                    rklib.Profile(l.copy()), rklib.Profile(r.copy()))
                got = kdistlib.ProfileDistance(summary=metrics.summary[summary], **opts).distance(
                    klib.Profile(l.copy()), klib.Profile(r.copy()))
                assert got == ref or (np.isnan(got) and np.isnan(ref))


def test_cli_commands_with_handle_double(golden, tmp_path):
    from kpal_b200 import kmer
    k = 4
    left = ko.count_sequences(golden["fixtures"]["LENGTH_60"], k)
    right = ko.count_sequences(golden["fixtures"]["LENGTH_60_MORE"], k)
    store = FakeH5()
    klib.Profile(left, 'b').save(store)
    klib.Profile(right, 'a').save(store)
    klib.Profile(left, '10').save(store)
    out = io.StringIO()
    kmer.distance_matrix(store, out, do_smooth=True, threshold=0, precision=2)
    lines = out.getvalue().split('\n')
    assert lines[:4] == ['3', '10', 'a', 'b']              # sorted(names): string order
    with pytest.raises(ValueError):
        kmer.distance_matrix(store, io.StringIO(), names=['a'])
    other = FakeH5()
    klib.Profile(np.zeros(4 ** 3, dtype=np.int64), 'a').save(other)
    out = io.StringIO()
    with pytest.raises(ValueError):
        kmer.distance(store, other, out, names_left=['a'], names_right=['a'], do_positive=True)
    with pytest.raises(ValueError):
        kmer.distance(store, other, out, names_left=['a', 'b'], names_right=['a'])
    out = io.StringIO()
    kmer.distance(store, store, out, names_left=['a'], names_right=['b'],
                  custom_pairwise='abs(left - right) / (left + right + 1)', precision=3)
    assert out.getvalue() == 'a b %.3f\n' % ko.distance(right, left, pairwise='sum')
    with pytest.raises(ValueError):
        kmer.count([io.StringIO('>a\nACGT\n')], FakeH5(), 3, names=['x', 'y'])
    assert kmer._name_from_handle(io.StringIO('x')) is None
    path = tmp_path / 'sample.one.fasta'
    path.write_text('>r\nACGT\n')
    with open(str(path)) as handle:
        assert kmer._name_from_handle(handle) == 'sample.one'
    # argparse wiring: missing sub-command / unknown file -> SystemExit(2) like parser.error
    with pytest.raises(SystemExit):
        kmer.main([])
    with pytest.raises(SystemExit):
        kmer.main(['matrix', str(tmp_path / 'missing.k'), '-'])


def test_read_text_takes_ascii_files_as_bytes(tmp_path):
    """klib._read_text: a text-mode FASTA file at its start is read through its binary buffer
    when that cannot change the result (plain ASCII, no '\\r'); every other handle goes through
    handle.read() and so keeps universal newlines, encodings and partial reads."""
    from kpal_b200.klib import _read_text

    def write(name, data):
        path = tmp_path / name
        path.write_bytes(data)
        return str(path)
    plain = write('plain.fa', b">r1 x\nACGT\nNN\n>r2\nTTGA")
    with open(plain) as handle:
        assert _read_text(handle) == b">r1 x\nACGT\nNN\n>r2\nTTGA"
    with open(plain) as handle:
        handle.readline()
        assert _read_text(handle) == "ACGT\nNN\n>r2\nTTGA"          # partly consumed: the rest, as text
    with open(plain, 'rb') as handle:
        assert _read_text(handle) == b">r1 x\nACGT\nNN\n>r2\nTTGA"
    with open(write('dos.fa', b">r1\r\nACGT\r\n")) as handle:
        assert _read_text(handle) == ">r1\nACGT\n"
    with open(write('mac.fa', b">r1\rACGT\r>r2\rTT\r")) as handle:
        assert _read_text(handle) == ">r1\nACGT\n>r2\nTT\n"
    with open(write('utf8.fa', u">ré\nACGT\n".encode('utf-8')), encoding='utf-8') as handle:
        assert _read_text(handle) == u">ré\nACGT\n"
    assert _read_text(io.StringIO(">x\nAC\n")) == ">x\nAC\n"
    assert _read_text(io.BytesIO(b">x\nAC\n")) == b">x\nAC\n"
    # ASCII-incompatible encodings whose bytes are all below 0x80 are decoded by the text layer
    for codec in ('utf-16-le', 'utf-32-be'):
        with open(write(codec + '.fa', u">r1\nACGT\n".encode(codec)), encoding=codec) as handle:
            assert _read_text(handle) == u">r1\nACGT\n"
    with open(plain, encoding='latin-1') as handle:
        assert _read_text(handle) == b">r1 x\nACGT\nNN\n>r2\nTTGA"
    # either form parses to the same records
    for text in (b">r1 x\nACGT\nNN\n>r2\nTTGA", ">r1 x\nACGT\nNN\n>r2\nTTGA"):
        assert ko.parse_fasta(text) == [("r1", "ACGTNN"), ("r2", "TTGA")]


def test_non_integer_counts_take_the_host_pipeline(monkeypatch):
    """The device path works on int64 counts.  Profiles holding anything else (float counts
    from a custom merger, say) are never truncated: they run the reference's NumPy pipeline."""
    def no_gpu(*args, **kwargs):
        raise AssertionError('device path used for non-integer counts')
    monkeypatch.setattr(kdistlib._cabi, 'pair_distance', no_gpu)
    monkeypatch.setattr(kdistlib._cabi, 'distance_matrix', no_gpu)
    rng = np.random.default_rng(5)
    a = klib.Profile(rng.integers(0, 9, 256) + 0.5, 'a')
    b = klib.Profile(rng.integers(0, 9, 256) * 1.25, 'b')
    c = klib.Profile(rng.integers(0, 9, 256).astype(np.float32), 'c')
    dist = kdistlib.ProfileDistance(do_scale=True)
    # kpal/kdistlib.py:150-161 + kpal/metrics.py:49-72,101-123 on the float counts themselves
    x, y = a.counts * (b.counts.sum() / a.counts.sum()), b.counts * 1.0
    if a.counts.sum() >= b.counts.sum():
        x, y = a.counts * 1.0, b.counts * (a.counts.sum() / b.counts.sum())
    nz = np.where(np.logical_or(x, y))
    want = (np.abs(x[nz] - y[nz]) / ((x[nz] + 1) * (y[nz] + 1))).sum() / (len(nz[0]) + 1)
    assert abs(dist.distance(a, b) - want) <= 1e-12 * abs(want)
    values = kdistlib.distance_matrix_values([a, b, c], dist)
    assert abs(values[1, 0] - want) <= 1e-12 * abs(want) and values[1, 0] == values[0, 1]
    assert kdistlib._integer_counts(klib.Profile(np.zeros(16, dtype=np.uint16)))
    assert kdistlib._integer_counts(klib.Profile([0] * 16))
    assert not kdistlib._integer_counts(klib.Profile(np.zeros(16)))
    big = np.zeros(16, dtype=np.uint64)
    big[3] = 2 ** 63
    assert not kdistlib._integer_counts(klib.Profile(big))


def test_h5lite_writer_warns_once(tmp_path, monkeypatch):
    """Without h5py, writing a profile file with the in-tree HDF5 writer says so (once)."""
    import builtins
    import warnings
    from kpal_b200 import kmer
    real_import = builtins.__import__

    def no_h5py(name, *args, **kwargs):
        if name == 'h5py':
            raise ImportError('no h5py')
        return real_import(name, *args, **kwargs)
    monkeypatch.setattr(builtins, '__import__', no_h5py)
    monkeypatch.setattr(kmer, '_h5lite_warned', False)
    monkeypatch.delenv('KPAL_B200_H5LITE', raising=False)
    with warnings.catch_warnings(record=True) as caught:
        warnings.simplefilter('always')
        kmer.ProfileFileType('w')(str(tmp_path / 'a.k')).close()
        kmer.ProfileFileType('w')(str(tmp_path / 'b.k')).close()
        kmer.ProfileFileType('r')(str(tmp_path / 'a.k')).close()
    assert len([w for w in caught if 'h5lite' in str(w.message)]) == 1
