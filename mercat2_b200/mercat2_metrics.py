"""Drop-in for the reference's lib/mercat2_metrics.py -- pI (ProMoST), molecular weight and
hydropathy computed by the device kernels (csrc/metrics.cuh); ``round(x, 2)`` is applied here,
exactly where the reference applies it.  ``sequence_metrics`` is the batched form."""
from __future__ import annotations

from . import _native


def sequence_metrics(seqs, engine=None):
    """[(pI or None, MW, hydro), ...] for raw sequences, rounded like the reference.  Raises
    ``KeyError``/``IndexError`` where the reference does (first residue unknown / empty)."""
    engine = engine or _native.default_engine()
    res = engine.sequence_metrics(seqs)
    out = []
    for i, seq in enumerate(seqs):
        status = int(res["status"][i])
        if status == 255:
            raise IndexError("string index out of range")          # seq[0] on an empty sequence
        if status == 2:
            raise KeyError(seq[0])
        mw, hydro = round(float(res["mw"][i]), 2), round(float(res["hydro"][i]), 2)
        if status == 1:
            print(seq[-1] + " not found!")
            out.append((None, mw, hydro))
        else:
            out.append((round(float(res["pi"][i]), 2), mw, hydro))
    return out


def predict_isoelectric_point_ProMoST(seq):
    """lib/mercat2_metrics.py:57-101"""
    return sequence_metrics([seq])[0][0]


def calculate_MW(seq):
    """lib/mercat2_metrics.py:158-163 (no first/last-residue requirements)"""
    res = _native.default_engine().sequence_metrics([seq])
    return round(float(res["mw"][0]), 2)


def calculate_hydro(seq):
    """lib/mercat2_metrics.py:166-170"""
    res = _native.default_engine().sequence_metrics([seq])
    return round(float(res["hydro"][0]), 2)


def file_metrics(file, engine=None):
    """Rows ``(header, first_word, length, pI, MW, hydro)`` for every non-empty record of a protein
    FASTA file, sorted by length descending -- the table part of plot_sample_metrics
    (lib/mercat2_figures.py:157-186).  A later record with the same header replaces the earlier one
    (``DataFrame.at[name, ...]``)."""
    from .mercat2_kmers import read_text_bytes
    return file_metrics_text(read_text_bytes(file), engine)


def file_metrics_text(data: bytes, engine=None):
    """``file_metrics`` for a protein FASTA text already in memory."""
    engine = engine or _native.default_engine()
    res = engine.protein_metrics(data)
    rows = {}
    for i in range(len(res["length"])):
        off, ln = int(res["header_off"][i]), int(res["header_len"][i])
        name = data[off:off + ln].decode("utf-8", errors="replace")      # headers may carry UTF-8 (the reference reads text)
        status = int(res["status"][i])
        if status == 2:
            raise KeyError(name)
        pi = None if status == 1 else round(float(res["pi"][i]), 2)
        rows[name] = (name, name.split()[0], float(res["length"][i]), pi,
                      round(float(res["mw"][i]), 2), round(float(res["hydro"][i]), 2))
    return sorted(rows.values(), key=lambda r: -r[2])
