"""Pre-processing that feeds the hot path (the reference's lib/mercat2_fasta.py, the parts in scope): ``removeN``
(+ ``-toupper``) and the FASTQ -> FASTA conversion.  Both produce the same ``clean/<base>_clean.fna.gz`` /
``clean/<base>.fna.gz`` artefacts as the reference, whose on-disk size drives the chunk trigger (bin/mercat2.py:101).
``fq2fa`` converts on the device (``fq2fa_device`` hands back the FASTA text still in HBM, ready to be counted);
``removeN`` is plain Python on the host like in the reference (its device version is the next step: DESIGN.md)."""
from __future__ import annotations

import gzip
import os
import re
import textwrap
from pathlib import Path

_N_RUN = re.compile(r"(N+)")


def split_sequenceN(header: str, sequence: str):
    """lib/mercat2_fasta.py:21-49: cut ``sequence`` at runs of upper-case 'N'; piece i (1-based) gets
    the header ``>{first word}_{i} {rest}`` and is wrapped at 80 columns."""
    n_lengths = [len(m.group(1)) for m in _N_RUN.finditer(sequence)]
    words = header.split()
    base, info = words[0], " ".join(words[1:])
    lines = []
    for i, piece in enumerate(_N_RUN.sub("\n", sequence).split("\n"), 1):
        lines.append(f">{base}_{i} {info}")
        lines += textwrap.wrap(piece, 80)
    return lines, n_lengths


def _records(reader):
    """(header without '>', stripped sequence lines) for every record; text before the first header
    is ignored (lib/mercat2_fasta.py:77-91)."""
    name, lines = None, []
    for raw in reader:
        text = raw.strip()
        if text.startswith(">"):
            if name is not None:
                yield name, lines
            name, lines = text[1:], []
        elif name is not None:
            lines.append(text)
    if name is not None:
        yield name, lines


def removeN(fasta, outpath, toupper: bool):
    """lib/mercat2_fasta.py:53-119 -> (path of clean/<base>_clean.fna.gz, {'GC Content': pct})."""
    os.makedirs(outpath, exist_ok=True)
    fasta = Path(fasta)
    basename = fasta.stem.split(".")[0]
    out_fasta = Path(outpath, f"{basename}_clean.fna.gz")
    gc_count = total = 0
    reader = gzip.open(fasta, "rt") if fasta.suffix == ".gz" else open(fasta, "r")
    with reader, gzip.open(out_fasta, "wt") as writer:
        for name, seq_lines in _records(reader):
            sequence = "".join(seq_lines)
            if "N" in sequence:
                out_lines, _ = split_sequenceN(name, sequence)
                for line in out_lines:
                    if line.startswith(">"):
                        print(line, file=writer)
                    else:
                        print(line.upper() if toupper else line, file=writer)
                    gc_count += line.count("G") + line.count("C")
                    total += len(line)
            else:
                print(">", name, sep="", file=writer)
                for line in seq_lines:
                    print(line.upper() if toupper else line, file=writer)
                gc_count += sequence.count("G") + sequence.count("C")
                total += len(sequence)
    return out_fasta.absolute(), {"GC Content": 100.0 * gc_count / total}


def fq2fa_host(fq_file, outpath, f_name: str) -> str:
    """lib/mercat2_fasta.py:175-198 (``sed -n '1~4s/^@/>/p;2~4p'``) in plain Python: of every 4 lines keep the first
    with its leading '@' turned into '>' (dropped if it does not start with '@') and the second.  Test helper / CPU
    hosts; the product path is ``fq2fa``."""
    os.makedirs(outpath, exist_ok=True)
    fna_file = os.path.join(outpath, f_name + ".fna.gz")
    opener = gzip.open if str(fq_file).endswith(".gz") else open
    with opener(fq_file, "rb") as reader, gzip.open(fna_file, "wb") as writer:
        for i, line in enumerate(reader):
            if i % 4 == 0:
                if line.startswith(b"@"):
                    writer.write(b">" + line[1:])
            elif i % 4 == 1:
                writer.write(line)
    return os.path.abspath(fna_file)


def fq2fa_device(fq_file, engine=None):
    """FASTQ file -> FASTA text on the device (``_native.DeviceText``: pass it to the counting calls as it is)."""
    from . import _native
    engine = engine or _native.default_engine()
    opener = gzip.open if str(fq_file).endswith(".gz") else open
    with opener(fq_file, "rb") as reader:
        data = reader.read()
    return engine.fastq_to_fasta(data)


def fq2fa(fq_file, outpath, f_name: str, engine=None) -> str:
    """lib/mercat2_fasta.py:175-198: writes ``<outpath>/<f_name>.fna.gz`` (same bytes as the reference's sed pipeline)
    from the text converted on the device; returns its path."""
    os.makedirs(outpath, exist_ok=True)
    fna_file = os.path.join(outpath, f_name + ".fna.gz")
    text = fq2fa_device(fq_file, engine)
    try:
        with gzip.open(fna_file, "wb") as writer:
            writer.write(text.to_bytes())
    finally:
        text.close()
    return os.path.abspath(fna_file)
