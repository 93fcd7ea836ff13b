"""Pure-Python restatement of MerCat2's k-mer hot path -- TEST INFRASTRUCTURE ONLY.

This module is the parity oracle for ``mercat2_b200``.  It restates, in plain
Python and with the same data structures (``str`` slices in a ``dict``), what the
reference does on the path

    chunk file -> parse FASTA -> slide k-window -> dict count -> per-file -c filter
    -> per-sample sum -> sorted TSV            (+ per-sequence pI / MW / hydropathy)

Every function cites the reference lines it follows (paths relative to the
reference checkout).  The restatement is *pinned*: ``oracle/make_golden.py`` runs the
reference's own modules (imported from ``/root/reference/lib`` in the build
container) and this oracle on the same inputs and stores the reference's answers
under ``tests/golden/``; ``tests/test_oracle_golden.py`` re-checks the oracle
against those vectors everywhere (no reference checkout needed at test time).

Nothing in the product (``mercat2_b200/``) imports this file.
"""
from __future__ import annotations

import glob
import gzip
import os
from pathlib import Path

# --------------------------------------------------------------------------------------
# A1 / A1b  k-mer counting                                    lib/mercat2_kmers.py:10-78
# --------------------------------------------------------------------------------------

def _open_text(path):
    """Text-mode open, gzip by suffix (lib/mercat2_kmers.py:47, lib/mercat2_Chunker.py:42)."""
    path = Path(path)
    return gzip.open(path, "rt") if path.suffix == ".gz" else open(path, "r")


def calculate_kmer_count(seq: str, kmer: int) -> dict:
    """All length-``kmer`` substrings of ``seq`` with multiplicity (lib/mercat2_kmers.py:10-28)."""
    table: dict = {}
    for start in range(len(seq) - kmer + 1):
        word = seq[start:start + kmer]
        table[word] = table.get(word, 0) + 1
    return table


def _count_into(table: dict, seq: str, kmer: int) -> None:
    for start in range(len(seq) - kmer + 1):
        word = seq[start:start + kmer]
        table[word] = table.get(word, 0) + 1


def record_sequences(lines):
    """Yield the sequence text of each record of a FASTA line stream.

    lib/mercat2_kmers.py:49-69 -- each line is ``strip()``-ed; a stripped line starting
    with '>' closes the record in progress (only non-empty ones are counted there, the
    final one always is -- an empty string has no windows for kmer >= 1, so yielding it
    or not is the same thing); any other line has every '*' removed and is appended.
    """
    pieces: list = []
    for raw in lines:
        text = raw.strip()
        if text.startswith(">"):
            if pieces:
                yield "".join(pieces)
                pieces = []
        else:
            pieces.append(text.replace("*", ""))
    yield "".join(pieces)


def find_kmers(file, kmer: int, min_count: int) -> dict:
    """Count k-mers of one FASTA file and keep those with count >= min_count.

    lib/mercat2_kmers.py:32-78.  Windows never cross records; the filter is applied to
    this file only (the caller sums filtered tables of several chunk files).
    """
    table: dict = {}
    with _open_text(file) as handle:
        for seq in record_sequences(handle):
            _count_into(table, seq, kmer)
    return {word: n for word, n in table.items() if n >= min_count}


def find_kmers_text(text: str, kmer: int, min_count: int) -> dict:
    """``find_kmers`` on an in-memory text (universal-newline semantics applied here)."""
    import io
    table: dict = {}
    for seq in record_sequences(io.StringIO(text, newline=None)):
        _count_into(table, seq, kmer)
    return {word: n for word, n in table.items() if n >= min_count}


# --------------------------------------------------------------------------------------
# A2  Chunker                                               lib/mercat2_Chunker.py:14-139
# --------------------------------------------------------------------------------------

_UNIT_SETS = (
    ("B", "K", "M", "G", "T", "P", "E", "Z", "Y"),
    ("byte", "kilo", "mega", "giga", "tera", "peta", "exa", "zetta", "iotta"),
    ("Bi", "Ki", "Mi", "Gi", "Ti", "Pi", "Ei", "Zi", "Yi"),
    ("byte", "kibi", "mebi", "gibi", "tebi", "pebi", "exbi", "zebi", "yobi"),
)


def human2bytes(text: str) -> int:
    """'100M' -> 104857600 etc. (lib/mercat2_Chunker.py:82-139): leading digits/dots are
    the number, the stripped remainder is a unit looked up in four unit families
    ('k' is accepted for 'K'); each step is a factor 1024; unknown unit -> ValueError."""
    original = text
    number = ""
    while text and text[0:1].isdigit() or text[0:1] == ".":
        number += text[0]
        text = text[1:]
    value = float(number)
    unit = text.strip()
    for family in _UNIT_SETS:
        if unit in family:
            break
    else:
        if unit == "k":
            family, unit = _UNIT_SETS[0], "K"
        else:
            raise ValueError("can't interpret %r" % original)
    return int(value * (1 << (10 * family.index(unit))))


def chunk_piece_name(path, index: int) -> str:
    """Piece file name ``<stem0>.<%05d><suffixes[:-1]>`` (lib/mercat2_Chunker.py:25-26,41)."""
    p = Path(path)
    return "%s.%05d%s" % (p.stem.split(".")[0], index, "".join(p.suffixes[:-1]))


def chunker_pieces(lines, chunksize: int, delim: str = ">"):
    """Split a line stream into pieces the way ``Chunker.stream_delim`` does.

    lib/mercat2_Chunker.py:39-59: at every line that *contains* ``delim`` the size of
    the current piece (bytes written so far; text is ASCII so chars == bytes, with
    universal newlines already folded to '\\n') is compared with ``chunksize``; if it is
    >= a new piece is started with that line.  Returns a list of lists of lines.
    """
    pieces = [[]]
    size = 0
    for line in lines:
        if delim in line and size >= chunksize:
            pieces.append([])
            size = 0
        pieces[-1].append(line)
        size += len(line.encode())
    return pieces


class Chunker:
    """File-writing restatement of lib/mercat2_Chunker.py:14-59 (delimiter mode only --
    ``stream_lines`` is never used by MerCat2)."""

    def __init__(self, path, dest, chunksize="1000M", delim=">"):
        self.path = str(path)
        self.dest = dest
        self.chunksize = human2bytes(chunksize)
        os.makedirs(dest, exist_ok=True)
        with _open_text_by_name(self.path) as handle:
            pieces = chunker_pieces(handle, self.chunksize, delim)
        for index, piece in enumerate(pieces):
            with open(os.path.join(dest, chunk_piece_name(path, index)), "w") as out:
                out.writelines(piece)
        self.files = glob.glob(os.path.join(dest, "*"))


def _open_text_by_name(path: str):
    return gzip.open(path, "rt") if path.endswith(".gz") else open(path, "r")


# --------------------------------------------------------------------------------------
# A3-A6  per-sample driver glue                                  bin/mercat2.py:86-137
# --------------------------------------------------------------------------------------

def chunk_files(filename, chunk_size_mb: int, outpath) -> list:
    """bin/mercat2.py:86-106 (+ the ``-s 0`` switch at :314/:417): chunk only when the
    ON-DISK size (gz size for .gz) reaches ``chunk_size_mb`` MiB."""
    if chunk_size_mb > 0 and os.stat(filename).st_size >= chunk_size_mb * 1024 * 1024:
        return sorted(Chunker(filename, outpath, "%dM" % chunk_size_mb, ">").files)
    return [str(filename)]


def merge_counts(tables) -> dict:
    """Sum already-filtered per-file tables (bin/mercat2.py:121-127)."""
    total: dict = {}
    for table in tables:
        for word, n in table.items():
            total[word] = total.get(word, 0) + n
    return total


def tsv_bytes(basename: str, table: dict) -> bytes:
    """The per-sample TSV (bin/mercat2.py:130-133): header ``k-mer\\t<base>_Count`` then
    rows sorted by k-mer string."""
    rows = ["k-mer\t%s_Count\n" % basename]
    rows += ["%s\t%d\n" % (word, n) for word, n in sorted(table.items())]
    return "".join(rows).encode()


def _find_kmers_job(args):
    return find_kmers(Path(args[0]), args[1], args[2])


def run_sample(basename: str, files, out_file, kmer: int, min_count: int, pool=None):
    """bin/mercat2.py:115-137 without Ray: one ``find_kmers`` task per file (on ``pool``
    when given), serial merge, sorted TSV; returns ``(basename, out_file or None)``."""
    jobs = [(str(f), kmer, min_count) for f in files]
    tables = pool.map(_find_kmers_job, jobs) if pool is not None else map(_find_kmers_job, jobs)
    total = merge_counts(tables)
    if not total:
        return basename, None
    with open(out_file, "wb") as out:
        out.write(tsv_bytes(basename, total))
    return basename, out_file


def count_sample(filename, kmer: int, min_count: int, chunk_size_mb: int, workdir, pool=None) -> dict:
    """chunk_files + per-chunk find_kmers + merge, returning the merged dict."""
    files = chunk_files(filename, chunk_size_mb, os.path.join(workdir, "chunks"))
    jobs = [(str(f), kmer, min_count) for f in files]
    tables = pool.map(_find_kmers_job, jobs) if pool is not None else map(_find_kmers_job, jobs)
    return merge_counts(tables)


# --------------------------------------------------------------------------------------
# A7-A9  protein metrics                                    lib/mercat2_metrics.py:16-170
# --------------------------------------------------------------------------------------

# pK triples (N-terminal, internal, C-terminal) of the ionisable residues, :16-25
PK_IONISABLE = {
    "K": (10.00, 9.80, 10.30), "R": (11.50, 12.50, 11.50), "H": (4.89, 6.08, 6.89),
    "D": (3.57, 4.07, 4.57), "E": (4.15, 4.45, 4.75), "C": (8.00, 8.28, 9.00),
    "Y": (9.34, 9.84, 10.34), "U": (5.20, 5.43, 5.60),
}
# terminal pK pairs (as first residue -> [0] used for the C-term test!, [1] ...) :27-54.
# NB the reference indexes these tables in a crossed way (first residue uses the
# ionisable table's index 2 or this table's index 1; last residue uses index 0 of
# either); the oracle keeps that exactly.
PK_TERMINAL = {
    "G": (7.50, 3.70), "A": (7.58, 3.75), "S": (6.86, 3.61), "P": (8.36, 3.40),
    "V": (7.44, 3.69), "T": (7.02, 3.57), "C": (8.12, 3.10), "I": (7.48, 3.72),
    "L": (7.46, 3.73), "J": (7.46, 3.73), "N": (7.22, 3.64), "D": (7.70, 3.50),
    "Q": (6.73, 3.57), "K": (6.67, 3.40), "E": (7.19, 3.50), "M": (6.98, 3.68),
    "H": (7.18, 3.17), "F": (6.96, 3.98), "R": (6.76, 3.41), "Y": (6.83, 3.60),
    "W": (7.11, 3.78), "X": (7.26, 3.57), "Z": (6.96, 3.535), "B": (7.46, 3.57),
    "U": (5.20, 5.60), "O": (7.00, 3.50),
}
# average residue masses :104-130 and Kyte-Doolittle scores :133-155
RESIDUE_MASS = {
    "A": 71.0788, "B": 114.6686, "C": 103.1388, "D": 115.0886, "E": 129.1155,
    "F": 147.1766, "G": 57.0519, "H": 137.1411, "I": 113.1594, "K": 128.1741,
    "L": 113.1594, "M": 131.1926, "N": 114.1038, "O": 237.3018, "P": 97.1167,
    "Q": 128.1307, "R": 156.1875, "S": 87.0782, "T": 101.1051, "U": 150.0388,
    "V": 99.1326, "W": 186.2132, "X": 111.1138, "Y": 163.176, "Z": 128.7531,
}
WATER_MASS = 18.01524
HYDROPATHY = {
    "A": 1.8, "R": -4.5, "N": -3.5, "D": -3.5, "C": 2.5, "Q": -3.5, "E": -3.5,
    "G": -0.4, "H": -3.2, "I": 4.5, "L": 3.8, "K": -3.9, "M": 1.9, "F": 2.8,
    "P": -1.6, "S": -0.8, "T": -0.7, "W": -0.9, "Y": -1.3, "V": 4.2,
}


def isoelectric_point_unrounded(seq: str):
    """ProMoST bisection (lib/mercat2_metrics.py:57-101) returning the pH *before* the
    final ``round(pH, 2)``; ``None`` when the last residue has no pK entry (:75-77);
    ``KeyError`` when the first residue has none (:66-69)."""
    first, last = seq[0], seq[-1]
    n_asp, n_glu, n_cys, n_tyr = seq.count("D"), seq.count("E"), seq.count("C"), seq.count("Y")
    n_his, n_lys, n_arg = seq.count("H"), seq.count("K"), seq.count("R")
    pk_first = PK_IONISABLE[first][2] if first in PK_IONISABLE else PK_TERMINAL[first][1]
    if last in PK_IONISABLE:
        pk_last = PK_IONISABLE[last][0]
    elif last in PK_TERMINAL:
        pk_last = PK_TERMINAL[last][0]
    else:
        return None
    ph, lo, hi = 6.51, 0.0, 14.0
    while True:
        charge = -1.0 / (1.0 + pow(10, pk_first - ph))
        charge_order = (
            -n_asp / (1.0 + pow(10, PK_IONISABLE["D"][1] - ph)),
            -n_glu / (1.0 + pow(10, PK_IONISABLE["E"][1] - ph)),
            -n_cys / (1.0 + pow(10, PK_IONISABLE["C"][1] - ph)),
            -n_tyr / (1.0 + pow(10, PK_IONISABLE["Y"][1] - ph)),
            n_his / (1.0 + pow(10, ph - PK_IONISABLE["H"][1])),
            1.0 / (1.0 + pow(10, ph - pk_last)),
            n_lys / (1.0 + pow(10, ph - PK_IONISABLE["K"][1])),
            n_arg / (1.0 + pow(10, ph - PK_IONISABLE["R"][1])),
        )
        for term in charge_order:          # same left-to-right order as :87
            charge += term
        if charge < 0.0:
            ph, hi = ph - (ph - lo) / 2.0, ph
        else:
            ph, lo = ph + (hi - ph) / 2.0, ph
        if ph - lo < 0.01 and hi - ph < 0.01:
            return ph


def predict_isoelectric_point_ProMoST(seq: str):
    """lib/mercat2_metrics.py:57-101 (prints ``"<c> not found!"`` and returns None for an
    unknown last residue, exactly like the reference)."""
    ph = isoelectric_point_unrounded(seq)
    if ph is None:
        print(seq[-1] + " not found!")
        return None
    return round(ph, 2)


def molecular_weight_unrounded(seq: str) -> float:
    total = 0.0
    for residue in seq:
        total += RESIDUE_MASS.get(residue, 0.0)
    return total + WATER_MASS


def calculate_MW(seq: str) -> float:
    """lib/mercat2_metrics.py:158-163."""
    return round(molecular_weight_unrounded(seq), 2)


def hydropathy_unrounded(seq: str) -> float:
    total = 0.0
    for residue in seq:
        total += HYDROPATHY.get(residue, 0.0)
    return total


def calculate_hydro(seq: str) -> float:
    """lib/mercat2_metrics.py:166-170."""
    return round(hydropathy_unrounded(seq), 2)


# --------------------------------------------------------------------------------------
# A10  per-file metrics table                               lib/mercat2_figures.py:150-186
# --------------------------------------------------------------------------------------

def protein_records(lines):
    """Records as ``plot_sample_metrics`` sees them (lib/mercat2_figures.py:157-173):
    per line ``strip()`` then ``rstrip('*')`` (internal '*' stay); only text after a
    header belongs to a record; yields ``(header_without_gt, sequence)``."""
    name = None
    pieces: list = []
    for raw in lines:
        text = raw.strip().rstrip("*")
        if text.startswith(">"):
            if name is not None:
                yield name, "".join(pieces)
            name, pieces = text[1:], []
        elif name is not None:
            pieces.append(text)
    if name is not None:
        yield name, "".join(pieces)


def sample_metrics(file) -> list:
    """Rows ``(header, first_word, length, pI, MW, hydro)`` of one protein file, empty
    sequences skipped, ordered by length descending (lib/mercat2_figures.py:174-186;
    tie order is unspecified there -- this oracle keeps file order among ties).
    A later record with the same full header overwrites the earlier one's values in
    place (``DataFrame.at[name, ...]``, :175-179)."""
    rows: dict = {}
    with _open_text(file) as handle:
        for name, seq in protein_records(handle):
            if not seq:
                continue
            rows[name] = (name, name.split()[0], float(len(seq)),
                          predict_isoelectric_point_ProMoST(seq), calculate_MW(seq), calculate_hydro(seq))
    return sorted(rows.values(), key=lambda r: -r[2])


# ---- merge_tsv / merge_tsv_T (lib/mercat2_report.py:98-194; row N3 of SURVEY.md 8f) ------------------------------
def _read_counts_tsv(path):
    with open(path) as handle:
        head = handle.readline()
        rows = [line.split() for line in handle if line.strip()]
    return head.split("\t")[0], {r[0]: r[1] for r in rows}


def merge_tsv(tsv_list: dict, out_file) -> None:
    """The reference's k-way merge, step for step (lib/mercat2_report.py:98-160).  Every file has a current row; the
    row label of an output line is the smallest NEXT key among the files that advanced on the previous line (all files
    at the start); a file emits its count and advances when its current key is <= the label, else emits '0'.
    NB: because files that did not advance are left out of that minimum, a pending smaller key can be emitted under a
    larger label -- with differing key sets the reference's table has out-of-order / repeated labels and misattributed
    counts (visible in tests/golden/merged_protein_k3.tsv.gz: 'CCR' twice around 'CCQ').  With identical key sets
    (e.g. nucleotide k=3) it is the plain sorted union.  ``merge_tsv_union`` below is the intended table."""
    names = sorted(tsv_list.keys())
    handles, lines, header = {}, {}, ""
    for name in names:
        handles[name] = open(tsv_list[name])
        head = handles[name].readline()
        if not header:
            header = head.split("\t")[0]
    try:
        with open(out_file, "w") as writer:
            print(header, "\t".join(names), sep="\t", file=writer)
            kmers = set()
            for name in names:
                lines[name] = handles[name].readline().split()
                kmers.add(lines[name][0])
            kmer = sorted(kmers)[0]
            while True:
                line = [kmer]
                kmers = set()
                for name in names:
                    if not lines[name] or lines[name][0] > kmer:
                        line.append("0")
                    else:
                        line.append(lines[name][1])
                        nxt = handles[name].readline().strip("\n").split("\t")
                        lines[name] = [x for x in nxt if len(x) > 0]
                        if lines[name]:
                            kmers.add(lines[name][0])
                print("\t".join(line), file=writer)
                if not kmers:
                    break
                kmer = sorted(kmers)[0]
    finally:
        for h in handles.values():
            h.close()


def merge_tsv_union(tsv_list: dict, out_file) -> None:
    """What merge_tsv is meant to produce: one row per k-mer of the sorted union of all samples, '0' where a sample
    lacks it.  Equal to the reference's output whenever all samples hold the same k-mers."""
    names = sorted(tsv_list.keys())
    header, tables = "", {}
    for name in names:
        head, table = _read_counts_tsv(tsv_list[name])
        header = header or head
        tables[name] = table
    kmers = sorted(set().union(*[set(t) for t in tables.values()])) if tables else []
    with open(out_file, "w") as writer:
        print(header, "\t".join(names), sep="\t", file=writer)
        for kmer in kmers:
            print("\t".join([kmer] + [tables[n].get(kmer, "0") for n in names]), file=writer)


def merge_tsv_T(tsv_list: dict, out_file) -> None:
    """Transposed table (lib/mercat2_report.py:164-194).  The reference's column order is a Python set's iteration
    order; this restatement sorts the columns (only the cell contents are comparable)."""
    names = sorted(tsv_list.keys())
    tables = {n: _read_counts_tsv(tsv_list[n])[1] for n in names}
    kmers = sorted(set().union(*[set(t) for t in tables.values()])) if tables else []
    with open(out_file, "w") as writer:
        print("sample", "\t".join(kmers), sep="\t", file=writer)
        for n in names:
            writer.write(n)
            for kmer in kmers:
                writer.write("\t" + tables[n].get(kmer, "0"))
            writer.write("\n")
