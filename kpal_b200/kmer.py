"""
Command line front end for the accelerated path -- mirror of the reference's
``kpal count``, ``kpal balance``, ``kpal distance`` and ``kpal matrix``
(reference kpal/kmer.py:112-146, 203-219, 541-700, 703-975): same arguments,
defaults, profile naming rules, output text and error behaviour
(``ValueError`` -> ``parser.error``).  The other fourteen kPAL sub-commands
are host-side analysis helpers outside the scope of this package
(SURVEY.md section 2, rows 8-10).

Profile files are HDF5 (reference doc/fileformat.rst:23-39); ``h5py`` is
imported lazily so that the package and its library API import without it.
"""
import argparse
import importlib
import os
import re
import sys

import numpy as np

from . import __version__, kdistlib, klib, metrics

LENGTH_ERROR = 'k-mer lengths of the files differ'
NAMES_COUNT_ERROR = ('number of profile names does not match number of '
                     'profiles')
PAIRED_NAMES_COUNT_ERROR = ('number of left and right profile names do not '
                            'match')

# dotted importable name, e.g. ``package.module.function``
_IMPORTABLE = re.compile(r'[_a-zA-Z][_a-zA-Z0-9]*(\.[_a-zA-Z][_a-zA-Z0-9]*)+$')

FORMAT_VERSION = '1.0.0'


# ------------------------------------------------------------------ file types
class FileType(object):
    """``argparse`` type for text files that refuses to overwrite existing
    files; ``-`` means stdin / stdout (reference kpal/__init__.py:47-82)."""
    def __init__(self, mode='r'):
        self._mode = mode

    def __call__(self, string):
        if string == '-':
            return sys.stdin if 'r' in self._mode else sys.stdout
        try:
            if 'w' in self._mode and os.path.exists(string):
                raise OSError('file exists')
            return open(string, self._mode)
        except OSError as error:
            raise argparse.ArgumentTypeError("can't open '%s': %s" % (string, error))


class ProfileFileType(object):
    """``argparse`` type for HDF5 *k*-mer profile files (reference
    kpal/__init__.py:85-111): new files get the ``format``, ``version`` and
    ``producer`` attributes and a ``/profiles`` group; existing files are
    validated."""
    def __init__(self, mode='r'):
        self._mode = mode

    def __call__(self, string):
        try:
            import h5py
        except ImportError:
            # no libhdf5 on this machine: the in-tree reader / writer of the same format
            from . import h5lite as h5py
            if 'w' in self._mode:
                _warn_h5lite_writer()
        try:
            if 'w' in self._mode and os.path.exists(string):
                raise IOError('file exists')
            handle = h5py.File(string, self._mode)
            if 'w' in self._mode:
                handle.attrs['format'] = 'kMer'
                handle.attrs['version'] = FORMAT_VERSION
                handle.attrs['producer'] = 'kPAL-B200 %s' % __version__
                handle.create_group('profiles')
            else:
                if handle.attrs.get('format') != 'kMer':
                    raise IOError('not a k-mer profile file')
                major = str(_text(handle.attrs['version'])).split('.')[0]
                if major != '1':
                    raise IOError('file format version %s not supported'
                                  % _text(handle.attrs['version']))
            return handle
        except IOError as error:
            raise argparse.ArgumentTypeError("can't open '%s': %s" % (string, error))


_h5lite_warned = False


def _warn_h5lite_writer():
    """One warning per process when a profile file is WRITTEN without h5py: the in-tree
    writer follows the HDF5 specification for the subset kPAL uses and round-trips through
    its own reader and the reference's test-suite, but its files have never been opened
    by libhdf5 itself (there is none in this image).  ``KPAL_B200_H5LITE=1`` acknowledges
    that and silences the warning."""
    global _h5lite_warned
    if _h5lite_warned or os.environ.get('KPAL_B200_H5LITE') == '1':
        return
    _h5lite_warned = True
    import warnings
    warnings.warn('h5py is not installed: writing the profile file with the in-tree HDF5 writer '
                  '(kpal_b200.h5lite), whose interoperability with libhdf5 / upstream kPAL is '
                  'unverified; install h5py for guaranteed-compatible files '
                  '(set KPAL_B200_H5LITE=1 to silence this warning)', RuntimeWarning, stacklevel=3)


def _text(value):
    return value.decode() if isinstance(value, bytes) else value


def _name_from_handle(handle):
    """Profile name from the file name of `handle`; ``None`` for streams and
    buffers (reference kpal/kmer.py:42-49)."""
    if not hasattr(handle, 'name') or handle.name.startswith('<'):
        return None
    return os.path.splitext(os.path.basename(str(handle.name)))[0]


def _resolve(custom, builtin, table, arguments):
    """A custom callable given as dotted name or expression, else the named
    built-in (reference kpal/kmer.py:657-681)."""
    if not custom:
        return table[builtin]
    if _IMPORTABLE.match(custom):
        module, name = custom.rsplit('.', 1)
        return getattr(importlib.import_module(module), name)
    return eval('lambda %s: %s' % (arguments, custom), {'np': np})


def _make_dist(distance_function, pairwise, custom_pairwise, do_smooth, summary,
               custom_summary, threshold, do_scale, down, do_positive, do_balance):
    return kdistlib.ProfileDistance(
        do_balance=do_balance, do_positive=do_positive, do_smooth=do_smooth,
        summary=_resolve(custom_summary, summary, metrics.summary, 'values'),
        threshold=threshold, do_scale=do_scale, down=down,
        pairwise=_resolve(custom_pairwise, pairwise, metrics.pairwise, 'left, right'),
        distance_function=metrics.vector_distance[distance_function])


# -------------------------------------------------------------------- commands
def count(input_handles, output_handle, size, names=None, by_record=False):
    """
    Make k-mer profiles from FASTA files.

    One profile per input file, named after the file (or numbered from 1), or
    with `by_record` one profile per FASTA record, named after the record and
    prefixed with the file's name when there are several inputs (reference
    kpal/kmer.py:112-146).
    """
    names = names or [_name_from_handle(h) for h in input_handles]
    if len(names) != len(input_handles):
        raise ValueError(NAMES_COUNT_ERROR)
    for input_handle, name in zip(input_handles, names):
        if by_record:
            # a device batch of rows at a time: counted on the GPU, statistics and deflate on
            # the host threads, all datasets of the batch created with one call
            prefix = name if len(input_handles) > 1 else None
            for batch_names, rows in klib.Profile.record_batches(input_handle, size, prefix=prefix):
                klib.save_profiles(output_handle, batch_names, rows)
        else:
            klib.Profile.from_fasta(input_handle, size, name=name).save(output_handle)


def balance(input_handle, output_handle, names=None):
    """
    Balance k-mer profiles.

    Every selected profile (default: all, alphabetically) is balanced and
    saved under its own name (reference kpal/kmer.py:203-219).
    """
    for name in names or sorted(input_handle['profiles']):
        profile = klib.Profile.from_file(input_handle, name=name)
        profile.balance()
        profile.save(output_handle)


def get_balance(input_handle, output_handle, precision=10, names=None):
    """
    Show the balance of k-mer profiles.

    One ``<name> <figure>`` line per profile: the multiset/prod distance
    between the forward and reverse-complement halves of the profile
    (reference kpal/kmer.py:222-247); one pass on the GPU
    (``kpal_show_balance``).
    """
    from . import _cabi
    template = '{{0:.{0}f}}'.format(precision)
    for name in names or sorted(input_handle['profiles']):
        profile = klib.Profile.from_file(input_handle, name=name)
        print(name, template.format(_cabi.show_balance(profile.counts)), file=output_handle)


def distance(input_handle_left, input_handle_right, output_handle,
             names_left=None, names_right=None, distance_function='default',
             pairwise='prod', custom_pairwise=None, do_smooth=False,
             summary='min', custom_summary=None, threshold=0, do_scale=False,
             down=False, do_positive=False, do_balance=False, precision=10):
    """
    Calculate the distance between two k-mer profiles.

    Profiles of the two files are paired in order and one
    ``<left> <right> <distance>`` line is written per pair (reference
    kpal/kmer.py:541-620).
    """
    names_left = names_left or sorted(input_handle_left['profiles'])
    names_right = names_right or sorted(input_handle_right['profiles'])
    if len(names_left) != len(names_right):
        raise ValueError(PAIRED_NAMES_COUNT_ERROR)
    dist = _make_dist(distance_function, pairwise, custom_pairwise, do_smooth, summary,
                      custom_summary, threshold, do_scale, down, do_positive, do_balance)
    template = '{{0:.{0}f}}'.format(precision)
    for name_left, name_right in zip(names_left, names_right):
        left = klib.Profile.from_file(input_handle_left, name=name_left)
        right = klib.Profile.from_file(input_handle_right, name=name_right)
        if left.length != right.length:
            raise ValueError(LENGTH_ERROR)
        print(name_left, name_right, template.format(dist.distance(left, right)),
              file=output_handle)


def distance_matrix(input_handle, output_handle, names=None,
                    distance_function='default', pairwise='prod',
                    custom_pairwise=None, do_smooth=False, summary='min',
                    custom_summary=None, threshold=0, do_scale=False,
                    down=False, do_positive=False, do_balance=False,
                    precision=10):
    """
    Make a distance matrix between any number of k-mer profiles.

    Profiles default to all of the file in alphabetical (string) order; at
    least two are needed (reference kpal/kmer.py:623-700).
    """
    names = names or sorted(input_handle['profiles'])
    if len(names) < 2:
        raise ValueError('you must give at least two k-mer profiles')
    dist = _make_dist(distance_function, pairwise, custom_pairwise, do_smooth, summary,
                      custom_summary, threshold, do_scale, down, do_positive, do_balance)
    # device fast path: datasets stream through a pinned slab straight to the GPU
    if kdistlib.distance_matrix_from_file(input_handle, names, output_handle, precision, dist):
        return
    profiles = []
    for name in names:
        profiles.append(klib.Profile.from_file(input_handle, name=name))
        if profiles[0].length != profiles[-1].length:
            raise ValueError(LENGTH_ERROR)
    kdistlib.distance_matrix(profiles, output_handle, precision, dist)


# ------------------------------------------------------------------------ main
def _first_paragraph(function):
    return function.__doc__.strip().split('\n\n')[0]


def main(args=None):
    """Command line interface (reference kpal/kmer.py:703-975, restricted to
    the four sub-commands on the accelerated path)."""
    profile_in = argparse.ArgumentParser(add_help=False)
    profile_in.add_argument('input_handle', metavar='INPUT', type=ProfileFileType('r'),
                            help='input k-mer profile file')
    profile_in.add_argument('-p', '--profiles', dest='names', metavar='NAME', type=str, nargs='+',
                            help='names of the k-mer profiles to consider (default: all '
                            'profiles in INPUT, in alphabetical order)')

    fasta_in = argparse.ArgumentParser(add_help=False)
    fasta_in.add_argument('input_handles', metavar='INPUT', type=FileType('r'), nargs='*',
                          default=[sys.stdin], help='input file (default: stdin)')

    profile_out = argparse.ArgumentParser(add_help=False)
    profile_out.add_argument('output_handle', metavar='OUTPUT', type=ProfileFileType('w'),
                             help='output k-mer profile file')

    dist_options = argparse.ArgumentParser(add_help=False)
    dist_options.add_argument('-d', dest='down', action='store_true', help='scale down')
    dist_options.add_argument('-s', dest='summary', type=str, default='min',
                              choices=metrics.summary,
                              help='summary function for dynamic smoothing (default: %(default)s)')
    dist_options.add_argument('-M', '--custom-summary', metavar='STRING', type=str,
                              dest='custom_summary',
                              help='custom summary function: an expression over the ndarray '
                              '"values" or an importable name')
    dist_options.add_argument('-t', dest='threshold', metavar='INT', type=int, default=0,
                              help='threshold for the summary function (default: %(default)s)')
    dist_options.add_argument('-n', metavar='INT', dest='precision', type=int, default=10,
                              help='precision in number of decimals (default: %(default)s)')
    dist_options.add_argument('-b', '--balance', dest='do_balance', action='store_true',
                              help='balance the profiles')
    dist_options.add_argument('--positive', dest='do_positive', action='store_true',
                              help='use only positive values')
    dist_options.add_argument('-S', '--scale', dest='do_scale', action='store_true',
                              help='scale the profiles')
    dist_options.add_argument('-m', '--smooth', dest='do_smooth', action='store_true',
                              help='smooth the profiles')
    dist_options.add_argument('-D', dest='distance_function', type=str, default='default',
                              choices=metrics.vector_distance,
                              help='choose distance function (default: %(default)s)')
    dist_options.add_argument('-P', dest='pairwise', type=str, default='prod',
                              choices=metrics.pairwise,
                              help='pairwise distance function for the multiset distance '
                              '(default: %(default)s)')
    dist_options.add_argument('-f', '--pairwise-function', metavar='STRING',
                              dest='custom_pairwise', type=str,
                              help='custom pairwise function: an expression over the ndarrays '
                              '"left" and "right" or an importable name')

    parser = argparse.ArgumentParser(
        prog='kpal', formatter_class=argparse.RawDescriptionHelpFormatter,
        description='kPAL hot path on B200: k-mer counting, balancing and profile distances.')
    parser.add_argument('-v', action='version', version='%s version %s' % ('kpal-b200', __version__))
    subparsers = parser.add_subparsers(dest='subcommand')
    subparsers.required = True

    p = subparsers.add_parser('count', parents=[fasta_in, profile_out],
                              description=_first_paragraph(count))
    p.add_argument('-p', '--profiles', dest='names', metavar='NAME', type=str, nargs='+',
                   help='names for the created k-mer profiles, one per INPUT (default: named '
                   'after the input files, or numbered from 1)')
    p.add_argument('-k', dest='size', metavar='SIZE', type=int, default=9,
                   help='k-mer size (%(type)s default: %(default)s)')
    p.add_argument('--by-record', '-r', dest='by_record', action='store_true',
                   help='make a k-mer profile per FASTA record instead of per FASTA file')
    p.set_defaults(func=count)

    p = subparsers.add_parser('balance', parents=[profile_in, profile_out],
                              description=_first_paragraph(balance))
    p.set_defaults(func=balance)

    p = subparsers.add_parser('showbalance', parents=[profile_in],
                              description=_first_paragraph(get_balance))
    p.add_argument('-n', metavar='INT', dest='precision', type=int, default=10,
                   help='precision in number of decimals (default: %(default)s)')
    p.set_defaults(func=get_balance, output_handle=sys.stdout)

    p = subparsers.add_parser('distance', parents=[dist_options],
                              description=_first_paragraph(distance))
    p.add_argument('input_handle_left', metavar='INPUT_LEFT', type=ProfileFileType('r'),
                   help='input k-mer profile file (left)')
    p.add_argument('input_handle_right', metavar='INPUT_RIGHT', type=ProfileFileType('r'),
                   help='input k-mer profile file (right)')
    p.add_argument('-l', '--profiles-left', dest='names_left', metavar='NAME', type=str,
                   nargs='+', help='names of the k-mer profiles to consider (left)')
    p.add_argument('-r', '--profiles-right', dest='names_right', metavar='NAME', type=str,
                   nargs='+', help='names of the k-mer profiles to consider (right)')
    p.set_defaults(func=distance, output_handle=sys.stdout)

    p = subparsers.add_parser('matrix', parents=[profile_in, dist_options],
                              description=_first_paragraph(distance_matrix))
    p.add_argument('output_handle', metavar='OUTPUT', type=FileType('w'), help='output file')
    p.set_defaults(func=distance_matrix)

    try:
        arguments = parser.parse_args(args)
    except IOError as error:
        parser.error(error)

    try:
        arguments.func(**dict((k, v) for k, v in vars(arguments).items()
                              if k not in ('func', 'subcommand')))
    except ValueError as error:
        parser.error(error)
    finally:
        # profile files are complete once closed (h5py does this at interpreter exit; here it
        # happens as soon as the command is done, so main() can be called in-process)
        for value in vars(arguments).values():
            for handle in (value if isinstance(value, list) else [value]):
                if hasattr(handle, 'create_dataset') and hasattr(handle, 'close'):
                    handle.close()


if __name__ == '__main__':
    main()
