"""Drop-in for the reference's lib/mercat2_Chunker.py.

``Chunker(path, dest, chunksize, delim='>')`` computes the reference's split points on the GPU
(virtual chunking: ``.offsets`` holds the byte offset where every piece starts).  For API
compatibility the pieces are also written to ``dest`` with the reference's names and ``.files``
lists them -- pass ``dest=None`` to skip writing (the counting pipeline never needs the files: it
hands the offsets' byte ranges to the engine).
"""
from __future__ import annotations

import glob
import os
import re
from pathlib import Path

from . import _native
from .mercat2_kmers import read_text_bytes

# unit spelling -> power of 1024 (the four symbol families of lib/mercat2_Chunker.py:110-117, plus 'k' as an alias of 'K')
_POWER = {unit: power
          for family in ("B K M G T P E Z Y", "byte kilo mega giga tera peta exa zetta iotta",
                         "Bi Ki Mi Gi Ti Pi Ei Zi Yi", "byte kibi mebi gibi tebi pebi exbi zebi yobi")
          for power, unit in enumerate(family.split())}
_POWER["k"] = 1
_SIZE = re.compile(r"([0-9.]*)(.*)", re.DOTALL)


def human2bytes(s: str) -> int:
    """'100M' -> 104857600, '0.5kilo' -> 512, '1 Gi' -> 2**30: a leading run of digits and dots, then a unit
    (same grammar and errors as lib/mercat2_Chunker.py:82-139: ValueError for a malformed number or an unknown unit)."""
    number, unit = _SIZE.fullmatch(s).groups()
    value = float(number)
    try:
        return int(value * (1 << (10 * _POWER[unit.strip()])))
    except KeyError:
        raise ValueError("can't interpret %r" % s) from None


def piece_name(path, index: int) -> str:
    """``<stem before first dot>.<%05d><suffixes without the last>`` (lib/mercat2_Chunker.py:25-26,41)."""
    p = Path(path)
    return "%s.%05d%s" % (p.stem.split(".")[0], index, "".join(p.suffixes[:-1]))


def split_offsets(data, chunk_bytes: int, engine=None) -> list:
    """Byte offsets where the reference's Chunker would start each piece of ``data`` (device)."""
    engine = engine or _native.default_engine()
    return engine.chunk_offsets(data, chunk_bytes)


class Chunker:
    def __init__(self, path, dest, chunksize="1000M", delim=None, lines=None):
        if lines is not None or delim != ">":
            raise NotImplementedError("mercat2_b200 Chunker supports delimiter mode with delim='>' only "
                                      "(the only mode MerCat2 uses: bin/mercat2.py:103)")
        self.path = str(path)
        self.dest = dest
        self.chunksize = human2bytes(chunksize)
        self.delim = delim
        self.lines = lines
        self.fn = os.path.basename(self.path)
        self.name = Path(path).stem.split(".")[0]
        self.ext = "".join(Path(path).suffixes[:-1])
        data = read_text_bytes(Path(self.path))
        self.offsets = split_offsets(data, self.chunksize) if self.chunksize > 0 else [0]
        self.nbytes = len(data)
        self.files = []
        if dest is not None:
            os.makedirs(dest, exist_ok=True)
            bounds = self.offsets + [len(data)]
            for i in range(len(self.offsets)):
                piece = data[bounds[i]:bounds[i + 1]]
                # the reference re-writes text-mode lines: "\r\n" and lone "\r" come out as "\n"
                piece = piece.replace(b"\r\n", b"\n").replace(b"\r", b"\n")
                with open(os.path.join(dest, piece_name(path, i)), "wb") as out:
                    out.write(piece)
            self.files = glob.glob(os.path.join(dest, "*"))
