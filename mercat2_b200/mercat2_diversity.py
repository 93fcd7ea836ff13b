"""Inputs of the reference's diversity / summary steps, straight from device tables.

The reference re-reads every per-sample TSV to rebuild the count vector it hands to scikit-bio
(lib/mercat2_diversity.py:13-27) and re-reads the combined TSV to pick the five k-mers with the largest mean count
(lib/mercat2_figures.py:41-65).  Both only need counts.  Here the vector and its abundance spectrum come from the table
that is still on the GPU, and the top rows from the sample x k-mer matrix; the ARITHMETIC of the metrics stays with
scikit-bio exactly like in the reference (it is an un-vendored dependency: re-implementing it would be unpinnable).
"""
from __future__ import annotations

from . import _native

ALPHA_METRICS = ["shannon", "simpson", "simpson_e", "goods_coverage", "fisher_alpha", "dominance", "chao1", "chao1_ci", "ace"]


def alpha_inputs(table) -> dict:
    """{'counts': uint64[rows] (the list lib/mercat2_diversity.py:23-27 builds), 'spectrum': device-side reductions}"""
    return {"counts": table.counts_array(), "spectrum": table.count_spectrum()}


def compute_alpha_diversity(basename: str, table, out_file):
    """lib/mercat2_diversity.py:13-52 with the count vector taken from the device table instead of the TSV.  Needs
    scikit-bio (the reference's own dependency); raises ImportError when it is not installed."""
    from skbio.diversity import alpha as skbio_alpha
    counts = table.counts_array().astype("int64").tolist()
    results = {}
    for func in ALPHA_METRICS:
        try:
            results[func] = getattr(skbio_alpha, func)(counts)
        except Exception:
            results[func] = "NA"
    with open(out_file, "w") as writer:
        print("Metric", basename, sep="\t", file=writer)
        for func in ALPHA_METRICS:
            value = results[func]
            if not isinstance(value, str):
                try:
                    value = round(value, 2)
                except Exception:
                    value = [round(x, 2) for x in value]
            print(func, value, sep="\t", file=writer)


def top_kmers(tables: dict, top: int = 5, engine=None):
    """The rows lib/mercat2_figures.py:41-65 keeps for its summary plot: (k-mer, [count per sample, names sorted]) of the
    `top` k-mers with the largest mean count over the samples."""
    engine = engine or _native.default_engine()
    names = sorted(tables.keys())
    matrix = engine.merge_tables([tables[n] for n in names])
    try:
        rows = matrix.top_rows(top)
        kmers, counts = matrix.arrays()
        return names, [(bytes(kmers[r]).decode(), [int(x) for x in counts[r]]) for r in rows]
    finally:
        matrix.close()
