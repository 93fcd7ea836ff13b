"""The reference's per-sample driver glue (bin/mercat2.py:86-137) on top of the engine.

``chunk_files`` -> a chunk size instead of piece files (virtual chunking), ``countKmers`` /
``run_mercat2`` -> one engine sample per input sample: every file is split at the reference's piece
boundaries on device, each piece is counted and filtered with ``-c`` on its own, the filtered
tables are summed, the sorted TSV is written."""
from __future__ import annotations

import os
from pathlib import Path

from . import _native


def chunk_trigger(filename, chunk_size_mb: int) -> int:
    """bin/mercat2.py:101 (+ ``-s 0`` at :314/:417): chunk only if the ON-DISK size (gz size for
    .gz files) reaches ``chunk_size_mb`` MiB.  Returns the chunk size in bytes, 0 = do not chunk."""
    if chunk_size_mb and chunk_size_mb > 0 and os.stat(filename).st_size >= chunk_size_mb * 1024 * 1024:
        return chunk_size_mb * 1024 * 1024
    return 0


def count_files(files, kmer: int, min_count: int, chunk_size_mb: int = 0, engine=None):
    """-> (Table, number of pieces counted) for one sample given its original file(s)."""
    engine = engine or _native.default_engine()
    sample = engine.sample(kmer, min_count)
    pieces = 0
    for file in files:
        # the engine reads the file itself: reader thread -> pinned buffers (inflate for '.gz') -> device window
        pieces += sample.add_file(Path(file), chunk_trigger(file, chunk_size_mb), gunzip=Path(file).suffix == ".gz")
    return sample.finish(), pieces


def countKmers(file, kmer, min_count):
    """bin/mercat2.py:112-114"""
    from .mercat2_kmers import find_kmers
    return find_kmers(Path(file), kmer, min_count)


def run_mercat2(basename: str, files: list, out_file, kmer, min_count, num_cores=None, chunk_size_mb: int = 0,
                engine=None, quiet: bool = False):
    """bin/mercat2.py:115-137.  ``files`` are the sample's files: either the reference's piece files
    (then ``chunk_size_mb`` stays 0) or the original file with ``chunk_size_mb`` = ``-s``.
    Returns ``(basename, out_file)`` or ``(basename, None)`` when no k-mer survives (no file written)."""
    table, _ = count_files(files, kmer, min_count, chunk_size_mb, engine)
    rows = table.rows
    if rows:
        if not quiet:
            print(f"Significant k-mers: {rows}")
        table.write_tsv(out_file, basename)
        return basename, out_file
    if not quiet:
        print("No significant k-mers found")
    return basename, None


def sample_metric_rows(file, chunk_size_mb: int = 0, engine=None):
    """Rows of report/metrics-protein.tsv for one protein sample the way the reference's driver produces them:
    ``samples[type][name] = [orig] + chunk_files(orig)`` (bin/mercat2.py:319-326, 422-429) and plot_sample_metrics
    walks EVERY file of that list (lib/mercat2_figures.py:157-186) -- so each sequence appears once for the original
    file (sorted by length, descending) and once more for the piece that holds it (each piece sorted on its own; the
    original again when the file was not chunked).  Pieces are the engine's virtual pieces: no files are written."""
    from . import mercat2_metrics
    from .mercat2_kmers import read_text_bytes
    engine = engine or _native.default_engine()
    data = read_text_bytes(Path(file))
    rows = mercat2_metrics.file_metrics_text(data, engine)
    yield from rows
    chunk = chunk_trigger(file, chunk_size_mb)
    if not chunk:
        yield from rows
        return
    offsets = list(engine.chunk_offsets(data, chunk)) + [len(data)]
    for a, b in zip(offsets[:-1], offsets[1:]):
        yield from mercat2_metrics.file_metrics_text(data[a:b], engine)


# ---------------------------------------------------------------------------------------------------
# many samples (bin/mercat2.py:411-448: one run_mercat2 task per sample)
# ---------------------------------------------------------------------------------------------------
BATCH_BYTES = 96 << 20          # text of one batched pass (about one ordinary 100 MB piece)
BATCH_SAMPLES = 256
# A sample joins a batched pass only below this much text.  Measured on a B200 (profiles/r02_workloads.jsonl, 64 proteomes,
# k=5): 1.6 MB samples 11 ms batched against 59 ms one by one, 15.6 MB samples 85 ms batched against 42 ms one by one (a
# sample of that size pays for its own dense table); the lines cross at 6-8 MB.
BATCH_TEXT_MAX = 6 << 20


def run_samples(samples: dict, out_dir, kmer: int, min_count: int, chunk_size_mb: int = 0, engine=None, quiet: bool = False) -> dict:
    """Count every sample of ``samples`` ({basename: file}) and write ``<out_dir>/<basename>_counts.tsv`` like the
    reference's sample loop.  Samples that are ONE piece (on-disk size below the ``-s`` trigger) and small are counted
    several at a time in a single pass of the engine (``Engine.count_batch``: keys carry the sample index) -- a
    1.6 M-residue proteome alone is all launch latency on a B200; the others go through ``run_mercat2`` one by one.
    Returns {basename: tsv path or None}."""
    from .mercat2_kmers import read_text_bytes
    engine = engine or _native.default_engine()
    results, batch, batch_bytes = {}, [], 0

    def flush():
        nonlocal batch, batch_bytes
        if not batch:
            return
        tables = engine.count_batch([text for _, text in batch], kmer, min_count)
        for (base, _), table in zip(batch, tables):
            out_file = os.path.join(out_dir, f"{base}_counts.tsv")
            rows = table.rows
            if rows:
                if not quiet:
                    print(f"Significant k-mers: {rows}")
                table.write_tsv(out_file, base)
                results[base] = out_file
            else:
                if not quiet:
                    print("No significant k-mers found")
                results[base] = None
            table.close()
        batch, batch_bytes = [], 0

    for base, file in samples.items():
        small = not chunk_trigger(file, chunk_size_mb) and os.stat(file).st_size <= BATCH_TEXT_MAX
        if not small:
            flush()
            results[base] = run_mercat2(base, [file], os.path.join(out_dir, f"{base}_counts.tsv"), kmer, min_count,
                                        chunk_size_mb=chunk_size_mb, engine=engine, quiet=quiet)[1]
            continue
        text = read_text_bytes(Path(file))
        if len(text) > BATCH_TEXT_MAX:                         # (a small file that inflates to a large text)
            small = False
        if not small:
            flush()
            results[base] = run_mercat2(base, [file], os.path.join(out_dir, f"{base}_counts.tsv"), kmer, min_count,
                                        chunk_size_mb=chunk_size_mb, engine=engine, quiet=quiet)[1]
            continue
        if batch and (batch_bytes + len(text) > BATCH_BYTES or len(batch) >= BATCH_SAMPLES):
            flush()
        batch.append((base, text))
        batch_bytes += len(text)
    flush()
    return {base: results[base] for base in samples}
