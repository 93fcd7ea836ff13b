"""merge_tsv / merge_tsv_T of the reference's lib/mercat2_report.py (:98-160, :164-194) on the engine: the per-sample
tables are parsed on the device (or taken as they are, still on the device).  Same names and arguments as the reference
(``tsv_list``: {sample name: TSV path}).

``merge_tsv`` writes the reference's table BYTE FOR BYTE by default.  The reference's k-way merge takes the label of a
row from the files that advanced on the previous line only, so with differing k-mer sets its table has repeated /
out-of-order labels and drops some rows; everything downstream of the reference (PCA, diversity) consumes that table, so
the drop-in reproduces it.  ``union=True`` writes what the merge is meant to be -- one row per k-mer of the sorted union,
'0' where a sample lacks it -- formatted on the GPU; the two are identical whenever all samples hold the same k-mers."""
from __future__ import annotations

import os

from . import _native


def _tables(tsv_list: dict, engine):
    names = sorted(tsv_list.keys())
    header, tables = "", []
    for name in names:
        with open(tsv_list[name], "rb") as handle:
            data = handle.read()
        if not header:
            header = data.split(b"\n", 1)[0].split(b"\t")[0].decode()       # :107-110 first column name of the first file
        tables.append(engine.table_from_tsv(data))
    return names, header, tables


def merge_tsv(tsv_list: dict, out_file: os.PathLike, engine=None, union: bool = False):
    """One column per sample (sorted by name); rows as the reference's merge_tsv writes them (default) or the sorted
    union of the k-mers (``union=True``)."""
    engine = engine or _native.default_engine()
    names, header, tables = _tables(tsv_list, engine)
    _write(engine, tables, names, out_file, header, union)
    for t in tables:
        t.close()


def _write(engine, tables, names, out_file, header, union):
    if union:
        matrix = engine.merge_tables(tables)
        matrix.write_tsv(out_file, header, names, transposed=False)
        matrix.close()
    else:
        engine.merge_tables_reference(tables, out_file, header, names)


def merge_tsv_T(tsv_list: dict, out_file: os.PathLike, engine=None):
    """Transposed: one row per sample, one column per k-mer.  The reference orders the columns by the iteration order
    of a Python ``set`` (unspecified); here they are sorted."""
    engine = engine or _native.default_engine()
    names, _, tables = _tables(tsv_list, engine)
    matrix = engine.merge_tables(tables)
    matrix.write_tsv(out_file, "sample", names, transposed=True)
    matrix.close()
    for t in tables:
        t.close()


def merge_tables(tables: dict, out_file: os.PathLike, header: str = "k-mer", engine=None, union: bool = False):
    """merge_tsv for tables that are still on the device ({sample name: Table}): no TSV is re-read."""
    engine = engine or _native.default_engine()
    names = sorted(tables.keys())
    _write(engine, [tables[n] for n in names], names, out_file, header, union)
