#!/usr/bin/env python
"""Generate tests/golden/ from the REFERENCE's own code -- TEST INFRASTRUCTURE ONLY.

Run in the build container (needs the read-only checkout at /root/reference):

    python oracle/make_golden.py

It imports the reference's unmodified modules (lib/mercat2_kmers.py, lib/mercat2_Chunker.py,
lib/mercat2_metrics.py, lib/mercat2_fasta.py) and records their answers on
  * a hand-written edge-case corpus (tests/golden/edge_cases.json),
  * BASELINE.json configs 1-3 (tests/golden/configs.json + expected TSVs),
  * the reference's committed result tree results/2023-11-29 (k=5, c=10, -s 1 / -s 10):
    TSV digests and Chunker piece sizes (tests/golden/reference_results.json),
  * protein metrics of DJ_pro (tests/golden/metrics_DJ_pro.tsv.gz),
and copies the (BSD-licensed) input DATA files the GPU box needs into tests/golden/data/.
No reference SOURCE is copied.  The GPU box has no /root/reference: tests read only the
files written here.
"""
from __future__ import annotations

import base64
import gzip
import hashlib
import io
import json
import os
import shutil
import sys
import tempfile
import types
from pathlib import Path

REF = Path("/root/reference")
ROOT = Path(__file__).resolve().parents[1]
GOLD = ROOT / "tests" / "golden"


def import_reference():
    """Import the reference's hot-path modules under their real package name."""
    pkg = types.ModuleType("mercat2_lib")
    pkg.__path__ = [str(REF / "lib")]
    sys.modules["mercat2_lib"] = pkg
    sys.modules.setdefault("pyrodigal", types.ModuleType("pyrodigal"))   # ORF caller: out of scope
    import importlib
    mods = {}
    for name in ("mercat2_kmers", "mercat2_Chunker", "mercat2_metrics", "mercat2_fasta"):
        mods[name] = importlib.import_module("mercat2_lib." + name)
    return mods


def md5(data: bytes) -> str:
    return hashlib.md5(data).hexdigest()


def tsv_of(basename: str, table: dict) -> bytes:
    out = io.StringIO()
    print("k-mer", f"{basename}_Count", sep="\t", file=out)
    for kmer, count in sorted(table.items()):
        print(kmer, count, sep="\t", file=out)
    return out.getvalue().encode()


EDGE_TEXTS = {
    # name: (text, [(k, c), ...])
    "simple_two_records": (">a\nACGTACGT\n>b\nTTTTACGT\n", [(1, 1), (3, 1), (3, 2), (8, 1), (9, 1)]),
    "text_before_first_header": ("ACGTAC\nGTAC\n>r1\nGGGCCC\n", [(3, 1), (4, 1)]),
    "blank_lines_and_crlf": (">r1\r\nACGT\r\n\r\nACGT\r\n>r2\r\nAC\r\nGT\r\n", [(3, 1), (5, 1)]),
    "lone_cr_newlines": (">r1\rACGTAC\rGTTT\r>r2\rACG\r", [(3, 1), (4, 1)]),
    "leading_ws_before_header": ("  \t>r1 desc\nACGT\n   >r2\nACGA\n", [(2, 1), (3, 1)]),
    "star_mid_sequence_joins": (">p1\nMK*LV*\n>p2\nMKLV*\nMK*\n", [(3, 1), (2, 2)]),
    "star_first_on_line": (">p\n*>AB\n*\n**A*\n", [(2, 1), (3, 1)]),
    "internal_space_kept": (">r\nAC GT\n  ACG T  \n", [(3, 1), (4, 1)]),
    "ws_then_star_tail": (">r\nAB *\nAB* \nAB\t*\t\nC D*E\n", [(2, 1), (3, 1)]),
    "tab_vt_ff_inside": (">r\nA\tC\x0bG\x0cT\n\x1cAC\x1d\x1e\x1fGT \n", [(2, 1), (3, 1)]),
    "record_shorter_than_k": (">a\nAC\n>b\nACGTA\n>c\nA\n", [(3, 1), (5, 1), (6, 1)]),
    "lowercase_distinct": (">r\nacgtACGTacgt\n", [(2, 1), (4, 1)]),
    "n_kept": (">h\nacgNACG\n>i\nNNNNACGTNNNN\n", [(3, 1), (4, 2)]),
    "bare_gt_header": (">\nACGT\n>\n>\nACGA\n", [(2, 1), (4, 1)]),
    "gt_inside_sequence_line": (">r\nAC>GT\nA>C\n", [(2, 1), (3, 1)]),
    "header_only": (">only a header\n", [(1, 1), (3, 1)]),
    "empty_file": ("", [(1, 1), (3, 1)]),
    "no_trailing_newline": (">r\nACGTACG", [(3, 1), (7, 1)]),
    "trailing_ws_no_newline": (">r\nACGT  ", [(2, 1), (4, 1)]),
    "only_whitespace_lines": (" \n\t\n>r\n  \nAC\n \nGT\n", [(2, 1), (4, 1)]),
    "digits_and_punct": (">r\nA1C2-G.T\n", [(2, 1), (3, 1)]),
    "protein_alphabet": (">p\nMKVLAAGIVGLLLAQWERTYIPASDFGHKLCVNMX*\n>q\nMKVLAAGIVBZUO*\n", [(3, 1), (5, 1)]),
    "repeats_filter": (">r\n" + "ACGT" * 50 + "\n>s\n" + "AAAAAAAAAA" * 7 + "\n", [(3, 10), (4, 10), (4, 49), (4, 50), (12, 10)]),
    "long_k": (">r\n" + "ACGTTGCAAGCTTAGC" * 9 + "\n>s\n" + "ACGTTGCAAGCTTAGC" * 9 + "N\n",
               [(31, 1), (32, 1), (33, 2), (40, 1), (64, 1), (65, 1)]),
    "multi_line_wrapped": (">g\n" + "\n".join(["ACGTTGCA" * 10] * 7 + ["ACG"]) + "\n", [(5, 1), (12, 2), (31, 1)]),
    "header_with_gt_in_desc": (">r1 a>b >c\nACGT\n>r2 >\nACGT\n", [(4, 1)]),
    "line_is_only_stars": (">p\nAB\n***\nCD\n", [(2, 1), (4, 1)]),
    # header lines may carry any UTF-8 text (the reference reads text mode and drops them, lib/mercat2_kmers.py:52)
    "utf8_in_header": (">r1 \u03b2-lactamase \u00b5 na\u00efve\nACGTACGT\n>r2 \u2603\nACGA\nTT\n", [(3, 1), (4, 1)]),
}


def build_edge_cases(ref):
    cases = []
    with tempfile.TemporaryDirectory() as tmp:
        for name, (text, ks) in EDGE_TEXTS.items():
            path = Path(tmp, name + ".fa")
            with open(path, "w", newline="", encoding="utf-8") as out:      # keep \r bytes as written
                out.write(text)
            for k, c in ks:
                expected = ref["mercat2_kmers"].find_kmers(path, k, c)
                cases.append({"name": name, "text_b64": base64.b64encode(text.encode()).decode(),
                              "k": k, "min_count": c, "expected": expected})
    return cases


def clean_text_of(ref, gz_in: Path, tmp: str, toupper: bool) -> bytes:
    out, _ = ref["mercat2_fasta"].removeN(gz_in, Path(tmp, "clean_up" if toupper else "clean"), toupper)
    return gzip.open(out, "rb").read()


def main():
    ref = import_reference()
    GOLD.mkdir(parents=True, exist_ok=True)
    data = GOLD / "data"
    (data / "fna_gz").mkdir(parents=True, exist_ok=True)
    (data / "faa_gz").mkdir(parents=True, exist_ok=True)
    exp = GOLD / "expected"
    exp.mkdir(exist_ok=True)

    json.dump(build_edge_cases(ref), open(GOLD / "edge_cases.json", "w"), indent=0, sort_keys=True)

    configs = {"nucleotide_k3_c10": {}, "protein_k3_c10": {}, "test_r1_k12": {}, "k5_c10_unchunked": {},
               "removeN": {}}
    with tempfile.TemporaryDirectory() as tmp:
        # ---- config 1: data/5-genomes-fna_gz, removeN, k=3 c=10 (and k=5 for the golden tree)
        for src in sorted((REF / "data/5-genomes-fna_gz").glob("*.fna.gz")):
            base = src.name.split(".")[0]
            shutil.copyfile(src, data / "fna_gz" / src.name)
            cleaned = clean_text_of(ref, src, tmp, False)
            configs["removeN"][base] = {"clean_md5": md5(cleaned), "clean_bytes": len(cleaned)}
            clean_path = Path(tmp, "clean", f"{base}_clean.fna.gz")
            for k, key in ((3, "nucleotide_k3_c10"), (5, "k5_c10_unchunked")):
                table = ref["mercat2_kmers"].find_kmers(clean_path, k, 10)
                tsv = tsv_of(base, table)
                configs[key][base] = {"rows": len(table), "total": sum(table.values()), "tsv_md5": md5(tsv)}
                if k == 3:
                    (exp / f"{base}_k3_c10.tsv").write_bytes(tsv)
        # removeN fixture with N runs and lower case, both -toupper settings
        scaf = REF / "data/Scaffolds_with-NNN.fna"
        with open(scaf, "rb") as fin, gzip.open(data / "Scaffolds_with-NNN.fna.gz", "wb", 9) as fout:
            shutil.copyfileobj(fin, fout)
        for toupper in (False, True):
            cleaned = clean_text_of(ref, scaf, tmp, toupper)
            key = "Scaffolds_toupper" if toupper else "Scaffolds"
            configs["removeN"][key] = {"clean_md5": md5(cleaned), "clean_bytes": len(cleaned)}
            p = Path(tmp, key + ".fna")
            p.write_bytes(cleaned)
            for k, c in ((4, 10), (12, 3)):
                table = ref["mercat2_kmers"].find_kmers(p, k, c)
                configs["removeN"][f"{key}_k{k}_c{c}"] = {"rows": len(table), "total": sum(table.values()),
                                                           "tsv_md5": md5(tsv_of(key, table))}

        # ---- config 2: data/5-genomes-faa protein k=3 c=10 (+k=5)
        for src in sorted((REF / "data/5-genomes-faa_gz").glob("*.faa.gz")):
            base = src.name.split(".")[0]
            shutil.copyfile(src, data / "faa_gz" / src.name)
            plain = REF / "data/5-genomes-faa" / f"{base}.faa"
            assert gzip.open(src, "rb").read() == plain.read_bytes()
            for k, key in ((3, "protein_k3_c10"), (5, "k5_c10_unchunked")):
                table = ref["mercat2_kmers"].find_kmers(plain, k, 10)
                tsv = tsv_of(base, table)
                configs[key][base] = {"rows": len(table), "total": sum(table.values()), "tsv_md5": md5(tsv)}
                if k == 3:
                    with gzip.open(exp / f"{base}_k3_c10.tsv.gz", "wb", 9) as out:
                        out.write(tsv)

        # ---- config 3: data/Test_R1.fastq -skipclean, k=12
        shutil.copyfile(REF / "data/Test_R1.fastq.gz", data / "Test_R1.fastq.gz")
        fna = ref["mercat2_fasta"].fq2fa(str(REF / "data/Test_R1.fastq"), os.path.join(tmp, "fq"), "Test_R1")
        fna_text = gzip.open(fna, "rb").read()
        configs["test_r1_k12"]["fasta_md5"] = md5(fna_text)
        for c in (1, 2, 10):
            table = ref["mercat2_kmers"].find_kmers(Path(fna), 12, c)
            tsv = tsv_of("Test_R1", table)
            configs["test_r1_k12"][f"c{c}"] = {"rows": len(table), "total": sum(table.values()),
                                               "with_N": sum("N" in w for w in table), "tsv_md5": md5(tsv)}
            if c == 2:
                (exp / "Test_R1_k12_c2.tsv").write_bytes(tsv)
    json.dump(configs, open(GOLD / "configs.json", "w"), indent=1, sort_keys=True)

    # ---- the reference's committed result tree (k=5, c=10, -s 1 and -s 10)
    tree = REF / "results/2023-11-29"
    results = {}
    for run in ("faa-5genomes-1", "faa-5genomes-10", "faa-5genomes_gz-1", "faa-5genomes_gz-10",
                "fna-5genomes_gz-1", "fna-5genomes_gz-10"):
        entry = {"tsv": {}, "chunks": {}}
        for kind in ("tsv_protein", "tsv_nucleotide"):
            for tsv in sorted((tree / run / kind).glob("*_counts.tsv")):
                blob = tsv.read_bytes()
                entry["tsv"][tsv.name.replace("_counts.tsv", "")] = {
                    "kind": kind, "rows": blob.count(b"\n") - 1, "tsv_md5": md5(blob)}
        for kind in ("chunks_protein", "chunks_nucleotide"):
            base_dir = tree / run / kind
            if base_dir.is_dir():
                for sample in sorted(base_dir.iterdir()):
                    entry["chunks"][sample.name] = [
                        {"name": f.name, "bytes": f.stat().st_size, "md5": md5(f.read_bytes())}
                        for f in sorted(sample.iterdir())]
        results[run] = entry
    json.dump(results, open(GOLD / "reference_results.json", "w"), indent=1, sort_keys=True)

    # ---- protein metrics golden (every sequence appears twice there; keep unique rows)
    rows = (tree / "DJ_gz-1/report/metrics-protein.tsv").read_text().splitlines()
    unique = [rows[0]] + sorted(set(rows[1:]))
    with gzip.open(GOLD / "metrics_DJ_pro.tsv.gz", "wb", 9) as out:
        out.write(("\n".join(unique) + "\n").encode())
    # and what the reference's metric functions say today on a few odd sequences
    odd = ["M", "MK", "ACDEFGHIKLMNPQRSTVWY", "UUUU", "BZXJO", "MKKKKKKKKK", "DDDDDDDDDE", "MAB", "M*K", "MK L"]
    m = ref["mercat2_metrics"]
    json.dump([{"seq": s, "pI": m.predict_isoelectric_point_ProMoST(s), "MW": m.calculate_MW(s),
                "hydro": m.calculate_hydro(s)} for s in odd], open(GOLD / "metrics_odd.json", "w"), indent=0)
    print("golden written to", GOLD)


if __name__ == "__main__":
    main()
