"""
GPU tests of the command functions on the accelerated path (kpal count /
balance / distance / matrix) through an HDF5 handle double -- mirrors reference
tests/test_kmer.py:83-125 (count), 185-193 (balance), 373-382 (distance),
455-467 (matrix).  Real h5py is not available in this image.
"""
import io

import numpy as np
import pytest

from kpal_b200 import kmer
from oracle import kpal_oracle as ko
from test_host_api import FakeH5

pytestmark = pytest.mark.gpu


def fasta_handle(tmp_path, name, sequences, names=None):
    names = names or ['sequence_%d' % (i + 1) for i in range(len(sequences))]
    path = tmp_path / name
    path.write_text('\n'.join('>' + n + '\n' + s for n, s in zip(names, sequences)) + '\n')
    return open(str(path))


def test_count_commands(golden, tmp_path):
    seqs = golden["fixtures"]["LENGTH_60"]
    more = golden["fixtures"]["LENGTH_60_MORE"]
    out = FakeH5()
    with fasta_handle(tmp_path, 'left.fa', seqs) as a, fasta_handle(tmp_path, 'right.fa', more) as b:
        kmer.count([a, b], out, 8)
    assert sorted(out['profiles']) == ['left', 'right']
    assert np.array_equal(out['profiles/left'][:], ko.count_sequences(seqs, 8))
    assert np.array_equal(out['profiles/right'][:], ko.count_sequences(more, 8))
    assert out['profiles/left'].attrs['length'] == 8
    assert out['profiles/left'].attrs['total'] == ko.count_sequences(seqs, 8).sum()

    out = FakeH5()
    with fasta_handle(tmp_path, 'x.fa', seqs) as a:
        kmer.count([a], out, 8, names=['custom'])
    assert list(out['profiles']) == ['custom']

    out = FakeH5()                                           # --by-record, one input: no prefix
    with fasta_handle(tmp_path, 'r.fa', seqs, names=['a', 'b', 'c', 'd']) as a:
        kmer.count([a], out, 5, by_record=True)
    assert sorted(out['profiles']) == ['a', 'b', 'c', 'd']
    for name, seq in zip('abcd', seqs):
        assert np.array_equal(out['profiles/' + name][:], ko.count_sequences([seq], 5))

    out = FakeH5()                                           # several inputs: file-name prefix
    with fasta_handle(tmp_path, 'p.fa', seqs[:2], names=['a', 'b']) as a, \
            fasta_handle(tmp_path, 'q.fa', more[:2], names=['a', 'b']) as b:
        kmer.count([a, b], out, 5, by_record=True)
    assert sorted(out['profiles']) == ['p_a', 'p_b', 'q_a', 'q_b']
    assert np.array_equal(out['profiles/q_b'][:], ko.count_sequences([more[1]], 5))

    out = FakeH5()                                           # nameless handle: numbered from 1
    kmer.count([io.StringIO('>x\n' + seqs[0] + '\n')], out, 4)
    assert list(out['profiles']) == ['1']

    with pytest.raises(ValueError):                          # record name with '.' (klib.py:181-182)
        kmer.count([io.StringIO('>read.1\nACGT\n')], FakeH5(), 2, by_record=True)


def test_balance_distance_matrix_commands(golden):
    left = ko.count_sequences(golden["fixtures"]["LENGTH_60"], 8)
    right = ko.count_sequences(golden["fixtures"]["LENGTH_60_MORE"], 8)
    store = FakeH5()
    from kpal_b200 import klib
    klib.Profile(left, 'a').save(store)
    klib.Profile(right, 'b').save(store)
    klib.Profile(left, 'c').save(store)

    balanced = FakeH5()
    kmer.balance(store, balanced)
    assert np.array_equal(balanced['profiles/a'][:], ko.balance(left))
    assert np.array_equal(balanced['profiles/b'][:], ko.balance(right))

    out = io.StringIO()
    kmer.distance(store, store, out, names_left=['a'], names_right=['b'])
    name_l, name_r, value = out.getvalue().split()
    assert (name_l, name_r) == ('a', 'b')
    assert float(value) == pytest.approx(0.4626209323, abs=1.01e-10)   # test_kmer.py:373-382

    out = io.StringIO()
    kmer.distance_matrix(store, out, precision=2)
    assert out.getvalue().strip().split('\n') == ['3', 'a', 'b', 'c', '0.46', '0.00 0.46']
    out = io.StringIO()
    kmer.distance_matrix(store, out, do_scale=True, do_balance=True, distance_function='euclidean',
                         precision=6)
    want = ko.distance(left, right, do_balance=True, do_scale=True, metric='euclidean')
    assert out.getvalue().split('\n')[4] == '%.6f' % want


def test_command_line_with_real_profile_files(golden, tutorial_texts, tmp_path, monkeypatch, capsys):
    """`kpal count / balance / showbalance / distance / matrix` as a user runs them, on real
    files: FASTA in, HDF5 profile files in between (kpal_b200.h5lite when h5py is absent),
    the matrix text out.  Mirrors reference tests/test_kmer.py:83-125,185-202,373-382,455-467
    and the tutorial (doc/tutorial.rst:36-146)."""
    import sys
    from kpal_b200 import h5lite
    if 'h5py' not in sys.modules:
        monkeypatch.setitem(sys.modules, 'h5py', None)       # make the fallback explicit
    names = sorted(tutorial_texts)
    paths = []
    for name in names:
        path = tmp_path / (name + '.fa')
        path.write_text(tutorial_texts[name])
        paths.append(str(path))
    merged = str(tmp_path / 'merged.k9')
    kmer.main(['count', '-k', '9'] + paths + [merged])
    with h5lite.File(merged) as f:
        assert f.attrs['format'] == 'kMer' and sorted(f['profiles']) == names
        for name in names:
            want = ko.count_fasta(tutorial_texts[name], 9)
            assert np.array_equal(f['profiles/' + name][:], want)
            assert f['profiles/' + name].attrs['total'] == want.sum()
            assert f['profiles/' + name].attrs['non_zero'] == np.count_nonzero(want)

    balanced = str(tmp_path / 'balanced.k9')
    kmer.main(['balance', merged, balanced])
    with h5lite.File(balanced) as f:
        assert np.array_equal(f['profiles/' + names[0]][:],
                              ko.balance(ko.count_fasta(tutorial_texts[names[0]], 9)))

    capsys.readouterr()
    kmer.main(['showbalance', merged, '-n', '3'])
    lines = capsys.readouterr().out.strip().split('\n')
    assert [line.split()[0] for line in lines] == names

    kmer.main(['distance', merged, balanced, '-l', names[0], '-r', names[1]])
    left, right, value = capsys.readouterr().out.split()
    want = ko.distance(ko.count_fasta(tutorial_texts[names[0]], 9),
                       ko.balance(ko.count_fasta(tutorial_texts[names[1]], 9)))
    assert (left, right) == (names[0], names[1]) and float(value) == pytest.approx(want, abs=1.01e-10)

    matrix = str(tmp_path / 'matrix.txt')
    kmer.main(['matrix', merged, matrix, '-S', '-b', '-n', '6'])
    rows = open(matrix).read().strip().split('\n')
    assert rows[0] == str(len(names)) and rows[1:1 + len(names)] == names
    counts = [ko.count_fasta(tutorial_texts[n], 9) for n in names]
    for i in range(1, len(names)):
        got = [float(x) for x in rows[len(names) + i].split()]
        want = [ko.distance(counts[i], counts[j], do_balance=True, do_scale=True) for j in range(i)]
        assert got == pytest.approx(want, abs=1.01e-6)

    by_record = str(tmp_path / 'records.k4')
    first = str(tmp_path / 'first.fa')
    with open(first, 'w') as handle:
        handle.write('\n'.join('>' + n + '\n' + s for n, s in zip('abcd', golden["fixtures"]["LENGTH_60"])) + '\n')
    kmer.main(['count', '-k', '4', '--by-record', first, by_record])
    with h5lite.File(by_record) as f:
        assert sorted(f['profiles']) == ['a', 'b', 'c', 'd']
        for name, seq in zip('abcd', golden["fixtures"]["LENGTH_60"]):
            assert np.array_equal(f['profiles/' + name][:], ko.count_sequences([seq], 4))

    with pytest.raises(SystemExit):                           # OUTPUT exists: argparse error, exit 2
        kmer.main(['count', '-k', '4', first, by_record])
