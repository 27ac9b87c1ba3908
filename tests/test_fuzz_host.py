"""
Property tests (hypothesis) of the host-side pieces, CPU only:

  * the C++ FASTA packer (csrc/pack.cpp: kpal_fasta_scan / kpal_fasta_pack) against the
    oracle's reader on arbitrary texts over an alphabet rich in the troublesome bytes
    (headers in odd places, '\\r', blanks, tabs, VT/FF, FS..US, empty lines, no final
    newline) -- the corners of Biopython's FastaIterator that the reference does not pin
     must at least be handled identically by both of our restatements;
  * the sequence-list packer against the plain text;
  * h5lite round trips of arbitrary names, dataset sizes and attribute values.
"""
import os
import tempfile

import numpy as np
from hypothesis import HealthCheck, given, settings, strategies as st

from kpal_b200 import _cabi, h5lite
from oracle import kpal_oracle as ko
from test_cabi import expected_stream, unpack

ALPHABET = "ACGTacgtNn>;- \t\r\n\x0b\x0c\x1c\x1f*"
COMMON = dict(deadline=None, suppress_health_check=[HealthCheck.too_slow])


@settings(max_examples=400, **COMMON)
@given(st.text(alphabet=ALPHABET, max_size=300))
def test_fasta_pack_equals_oracle_reader_on_arbitrary_text(text):
    records = ko.parse_fasta(text)
    codes, valid, rec_starts, names, n_bases = _cabi.fasta_pack(text)
    assert names == [n for n, _ in records]
    assert unpack(codes, valid, n_bases) == expected_stream([s for _, s in records])
    assert rec_starts.tolist() == np.cumsum([0] + [len(s) + 1 for _, s in records]).tolist()


@settings(max_examples=200, **COMMON)
@given(st.lists(st.text(alphabet="ACGTacgtNRYKM-*", max_size=120), max_size=12))
def test_pack_sequences_equals_text(seqs):
    codes, valid, rec_starts, n_bases = _cabi.pack_sequences(seqs)
    assert n_bases == sum(len(s) + 1 for s in seqs)
    assert unpack(codes, valid, n_bases) == expected_stream(seqs)
    assert rec_starts.tolist() == np.cumsum([0] + [len(s) + 1 for s in seqs]).tolist()


names = st.text(alphabet=st.characters(blacklist_characters="/.\x00", blacklist_categories=("Cs",)),
                min_size=1, max_size=24)
values = st.one_of(st.integers(-2 ** 62, 2 ** 62), st.floats(allow_nan=False, width=64),
                   st.text(alphabet=st.characters(blacklist_characters="\x00", blacklist_categories=("Cs",)),
                           max_size=40))


@settings(max_examples=60, **COMMON)
@given(st.dictionaries(names, st.tuples(st.integers(0, 6), st.integers(0, 2 ** 31), st.dictionaries(names, values, max_size=4)),
                       max_size=20))
def test_h5lite_round_trip_of_arbitrary_profiles(profiles):
    with tempfile.TemporaryDirectory() as directory:
        path = os.path.join(directory, 'fuzz.h5')
        expected = {}
        with h5lite.File(path, 'w') as f:
            group = f.create_group('profiles')
            for name, (k, seed, attrs) in profiles.items():
                counts = np.random.default_rng(seed).integers(0, 1000, 4 ** k).astype(np.int64)
                dataset = group.create_dataset(name, data=counts, dtype='int64', compression='gzip')
                for key, value in attrs.items():
                    dataset.attrs[key] = value
                expected[name] = (counts, attrs)
        with h5lite.File(path) as f:
            assert sorted(f['profiles'].keys(), key=lambda n: n.encode('utf-8')) == \
                sorted(expected, key=lambda n: n.encode('utf-8'))
            for name, (counts, attrs) in expected.items():
                dataset = f['profiles'][name]
                assert np.array_equal(dataset[:], counts)
                assert set(dataset.attrs.keys()) == set(attrs)
                for key, value in attrs.items():
                    got = dataset.attrs[key]
                    assert got == value and type(got) is (str if isinstance(value, str) else type(got))
