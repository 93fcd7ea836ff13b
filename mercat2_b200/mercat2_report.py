"""merge_tsv / merge_tsv_T of the reference's lib/mercat2_report.py (:98-160, :164-194) on the engine: the per-sample
tables are parsed (or taken as they are, still on the device), their k-mers united and sorted, and the sample x k-mer
matrix is formatted on the GPU.  Same names and arguments as the reference (``tsv_list``: {sample name: TSV path})."""
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


def merge_tsv(tsv_list: dict, out_file: os.PathLike, engine=None):
    """One row per k-mer of the sorted union, one column per sample (sorted by name), '0' where a sample lacks it."""
    engine = engine or _native.default_engine()
    names, header, tables = _tables(tsv_list, engine)
    matrix = engine.merge_tables(tables)
    matrix.write_tsv(out_file, header, names, transposed=False)
    matrix.close()
    for t in tables:
        t.close()


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


def merge_tables(tables: dict, out_file: os.PathLike, header: str = "k-mer", engine=None):
    """merge_tsv for tables that are still on the device ({sample name: Table}): no TSV is re-read."""
    engine = engine or _native.default_engine()
    names = sorted(tables.keys())
    matrix = engine.merge_tables([tables[n] for n in names])
    matrix.write_tsv(out_file, header, names, transposed=False)
    matrix.close()
