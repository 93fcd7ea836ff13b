"""Drop-in for the reference's lib/mercat2_kmers.py -- k-mer counting on the GPU.

Same entry points and return types (``dict[str, int]``).  Keys are raw substrings exactly as in
the reference: no canonicalisation, no alphabet check, case-sensitive; windows never cross records.
Differences, on purpose and loud: input must be 7-bit ASCII (``NonAsciiError`` otherwise -- the
reference would decode UTF-8 and count code points) and ``kmer`` must be in 1..128.
"""
from __future__ import annotations

import gzip
from pathlib import Path

from . import _native


def read_text_bytes(file) -> bytes:
    """The bytes the reference's parser sees (lib/mercat2_kmers.py:47): gunzip when the suffix is
    '.gz'.  Newline translation ('\\r\\n', lone '\\r') is done on device, not here."""
    file = Path(file)
    if file.suffix == ".gz":
        with gzip.open(file, "rb") as handle:
            return handle.read()
    with open(file, "rb") as handle:
        return handle.read()


def calculateKmerCount(seq: str, kmer: int) -> dict:
    """Counts of every length-``kmer`` substring of ``seq`` (lib/mercat2_kmers.py:10-28)."""
    if kmer < 1:
        raise ValueError("kmer must be >= 1")
    if len(seq) < kmer:
        return {}
    table = _native.default_engine().count_symbols(seq, kmer, 1)
    return table.to_dict()


def find_kmers(file: Path, kmer: int, min_count: int) -> dict:
    """k-mer counts of one (optionally gzipped) FASTA file, keeping counts >= ``min_count``
    (lib/mercat2_kmers.py:32-78).  ``file`` must be a ``pathlib.Path`` like in the reference
    (``.suffix`` decides gzip)."""
    if kmer < 1:
        raise ValueError("kmer must be >= 1")
    file = Path(file)
    sample = _native.default_engine().sample(kmer, min_count)
    sample.add_file(file, 0, gunzip=file.suffix == ".gz")          # one piece = the whole file (kmers.py:47-78)
    return sample.finish().to_dict()
