"""
CPU test of the HOST LOGIC of the command line (kpal_b200.kmer.main -> klib / kdistlib ->
h5lite files -> text output) with the GPU entry points replaced by a TEST DOUBLE.

This is test infrastructure: the double below (oracle-backed, monkeypatched into
kpal_b200._cabi for the duration of one test) stands in for libkpal_b200's compute calls so
that the Python glue around them -- argument handling, profile naming, file layout, the
slab-wise matrix loader, the matrix text -- is exercised on a machine without a GPU.  The
product has no such path: without the double every one of these commands raises
(tests/test_cabi.py::test_no_cpu_fallback_without_gpu).  The same commands run against the
real kernels in tests/test_gpu_cli.py.  Mirrors reference tests/test_kmer.py:83-125,185-202,
373-382,455-467.
"""
import io

import numpy as np
import pytest

from kpal_b200 import _cabi, h5lite, kmer
from oracle import kpal_oracle as ko
from test_cabi import unpack


class _SessionDouble(object):
    def __init__(self, n, k, **options):
        self.n, self.k, self.options = n, k, options
        self.slab = np.zeros((min(n, 3), 4 ** k), dtype=np.int64)       # small: several pushes
        self.rows = []

    def push(self, m):
        self.rows.extend(self.slab[:m].copy())

    def finish(self):
        assert len(self.rows) == self.n
        return ko.distance_matrix_values(self.rows, **self.options)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        pass


@pytest.fixture
def gpu_double(monkeypatch):
    def by_record(codes, valid, n_bases, rec_starts, first, n, k, balance=False, out=None):
        stream = unpack(codes, valid, n_bases)
        rows = []
        for r in range(first, first + n):
            seq = ''.join('ACGTN'[c] for c in stream[int(rec_starts[r]):int(rec_starts[r + 1]) - 1])
            counts = ko.count_sequences([seq], k)
            rows.append(ko.balance(counts) if balance else counts)
        rows = np.array(rows, dtype=np.int64).reshape(n, 4 ** k)
        if out is not None:                      # the reused batch buffer of Profile.record_batches
            view = out.reshape(-1)[:n * 4 ** k].reshape(n, 4 ** k)
            view[:] = rows
            return view
        return rows

    def count_fasta(text, k, balance=False, out=None):
        counts = ko.count_fasta(text if isinstance(text, str) else text.decode('latin-1'), k)
        return ko.balance(counts) if balance else counts

    def balance(counts):
        counts[:] = ko.balance(counts)
        return counts

    def pair_distance(left, right, **options):
        return ko.distance(left, right, **options)

    monkeypatch.setattr(_cabi, 'require_gpu', lambda: None)
    monkeypatch.setattr(_cabi, 'count_fasta', count_fasta)
    monkeypatch.setattr(_cabi, 'count_sequences',
                        lambda seqs, k, balance=False: (ko.balance(ko.count_sequences(list(seqs), k)) if balance
                                                        else ko.count_sequences(list(seqs), k)))
    monkeypatch.setattr(_cabi, 'count_by_record', by_record)
    monkeypatch.setattr(_cabi, 'balance', balance)
    monkeypatch.setattr(_cabi, 'split', ko.split)
    monkeypatch.setattr(_cabi, 'show_balance', ko.show_balance)
    monkeypatch.setattr(_cabi, 'pair_distance', pair_distance)
    monkeypatch.setattr(_cabi, 'distance_matrix', lambda profiles, **options: ko.distance_matrix_values(list(profiles), **options))
    monkeypatch.setattr(_cabi, 'MatrixSession', _SessionDouble)


def test_command_line_host_logic(gpu_double, golden, tutorial_texts, tmp_path, monkeypatch, capsys):
    import sys
    monkeypatch.setitem(sys.modules, 'h5py', None)            # profile files through h5lite
    names = sorted(tutorial_texts)
    paths = []
    for name in names:
        path = tmp_path / (name + '.fa')
        path.write_text(tutorial_texts[name])
        paths.append(str(path))
    counts = dict((n, ko.count_fasta(tutorial_texts[n], 9)) for n in names)

    merged = str(tmp_path / 'merged.k9')
    kmer.main(['count', '-k', '9'] + paths + [merged])
    with h5lite.File(merged) as f:
        assert f.attrs['format'] == 'kMer' and f.attrs['version'] == '1.0.0'
        assert sorted(f['profiles']) == names
        for name in names:
            dataset = f['profiles/' + name]
            assert np.array_equal(dataset[:], counts[name]) and dataset.compression == 'gzip'
            assert dataset.attrs['length'] == 9 and dataset.attrs['total'] == counts[name].sum()
            assert dataset.attrs['non_zero'] == np.count_nonzero(counts[name])
            assert dataset.attrs['median'] == np.median(counts[name]) and dataset.attrs['std'] == counts[name].std()

    renamed = str(tmp_path / 'renamed.k9')
    kmer.main(['count', '-k', '9', '-p', 'x', 'y', '--', paths[0], paths[1], renamed])
    with h5lite.File(renamed) as f:
        assert sorted(f['profiles']) == ['x', 'y'] and np.array_equal(f['profiles/y'][:], counts[names[1]])
    with pytest.raises(SystemExit):                           # names and inputs differ in number
        kmer.main(['count', '-k', '9', '-p', 'x', '--', paths[0], paths[1], str(tmp_path / 'bad.k9')])

    balanced = str(tmp_path / 'balanced.k9')
    kmer.main(['balance', merged, balanced, '-p', names[2], names[0]])
    with h5lite.File(balanced) as f:
        assert sorted(f['profiles']) == sorted([names[0], names[2]])
        assert np.array_equal(f['profiles/' + names[2]][:], ko.balance(counts[names[2]]))

    capsys.readouterr()
    kmer.main(['showbalance', merged, '-n', '4'])
    lines = capsys.readouterr().out.strip().split('\n')
    assert [line.split()[0] for line in lines] == names
    assert lines[0].split()[1] == '%.4f' % ko.show_balance(counts[names[0]])

    kmer.main(['distance', merged, merged, '-l', names[0], '-r', names[3], '-S', '-P', 'sum', '-n', '8'])
    left, right, value = capsys.readouterr().out.split()
    assert (left, right) == (names[0], names[3])
    assert value == '%.8f' % ko.distance(counts[names[0]], counts[names[3]], do_scale=True, pairwise='sum')

    # matrix: slab-wise loader (three rows per push in the double), then the text writer
    matrix = str(tmp_path / 'matrix.txt')
    kmer.main(['matrix', merged, matrix, '-S', '-b', '-n', '6'])
    want = ko.format_matrix(names, ko.distance_matrix_values([counts[n] for n in names], do_balance=True, do_scale=True), 6)
    assert open(matrix).read() == want
    subset = str(tmp_path / 'subset.txt')
    kmer.main(['matrix', merged, subset, '-p', names[4], names[1], names[6], '-D', 'euclidean'])
    picked = [names[4], names[1], names[6]]
    want = ko.format_matrix(picked, ko.distance_matrix_values([counts[n] for n in picked], metric='euclidean'), 10)
    assert open(subset).read() == want
    # host-path options (smoothing) take the list-of-profiles route and the reference's pipeline
    smooth = str(tmp_path / 'smooth.txt')
    kmer.main(['matrix', merged, smooth, '-p', names[0], names[1], '-m', '-t', '2', '-n', '5'])
    rows = open(smooth).read().strip().split('\n')
    assert rows[:3] == ['2', names[0], names[1]] and len(rows) == 4 and float(rows[3]) >= 0.0

    records = str(tmp_path / 'records.k4')
    first = str(tmp_path / 'first.fa')
    seqs = golden["fixtures"]["LENGTH_60"]
    with open(first, 'w') as handle:
        handle.write('\n'.join('>' + n + ' description\n' + s for n, s in zip('abcd', seqs)) + '\n')
    kmer.main(['count', '-k', '4', '--by-record', first, records])
    with h5lite.File(records) as f:
        assert sorted(f['profiles']) == ['a', 'b', 'c', 'd']
        for name, seq in zip('abcd', seqs):
            assert np.array_equal(f['profiles/' + name][:], ko.count_sequences([seq], 4))
    with pytest.raises(SystemExit):                           # OUTPUT exists
        kmer.main(['count', '-k', '4', first, records])
    with pytest.raises(SystemExit):                           # not a profile file
        kmer.main(['balance', first, str(tmp_path / 'never.k4')])
