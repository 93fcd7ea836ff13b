"""The oracle (oracle/mercat2_oracle.py) against vectors produced by the reference's own
code (oracle/make_golden.py) and against the reference's committed result tree."""
import gzip
import hashlib
import json
import os

import pytest

from oracle import mercat2_oracle as orc
from conftest import GOLDEN, read_maybe_gz


def md5(b):
    return hashlib.md5(b).hexdigest()


def test_edge_cases(edge_cases, tmp_path):
    assert len(edge_cases) > 50
    for case in edge_cases:
        path = tmp_path / "case.fa"
        path.write_bytes(case["text"])
        got = orc.find_kmers(path, case["k"], case["min_count"])
        assert got == case["expected"], (case["name"], case["k"], case["min_count"])
        assert orc.find_kmers_text(case["text"].decode(), case["k"], case["min_count"]) == case["expected"]


def test_calculate_kmer_count():
    assert orc.calculate_kmer_count("ACGTACG", 3) == {"ACG": 2, "CGT": 1, "GTA": 1, "TAC": 1}
    assert orc.calculate_kmer_count("AC", 3) == {}


@pytest.mark.parametrize("base", ["DJ_pro", "GIC31_pro", "RW1_pro", "RW2_pro", "Rleg_pro"])
def test_protein_k3_c10(base, golden_configs):
    src = GOLDEN / "data/faa_gz" / f"{base}.faa.gz"
    table = orc.find_kmers(src, 3, 10)
    want = golden_configs["protein_k3_c10"][base]
    assert len(table) == want["rows"] and sum(table.values()) == want["total"]
    tsv = orc.tsv_bytes(base, table)
    assert md5(tsv) == want["tsv_md5"]
    assert tsv == gzip.open(GOLDEN / "expected" / f"{base}_k3_c10.tsv.gz", "rb").read()


def test_test_r1_k12(golden_configs, tmp_path):
    # fastq -> fasta is host plumbing (lib/mercat2_fasta.py:175-198): lines 1 mod 4 ('@'->'>') and 2 mod 4
    lines = gzip.open(GOLDEN / "data/Test_R1.fastq.gz", "rt").read().split("\n")
    fasta = "".join((">" + l[1:] if i % 4 == 0 and l.startswith("@") else l) + "\n"
                    for i, l in enumerate(lines[:-1]) if i % 4 < 2)
    assert md5(fasta.encode()) == golden_configs["test_r1_k12"]["fasta_md5"]
    path = tmp_path / "Test_R1.fna"
    path.write_text(fasta)
    for c in (1, 2, 10):
        table = orc.find_kmers(path, 12, c)
        want = golden_configs["test_r1_k12"][f"c{c}"]
        assert len(table) == want["rows"] and sum(table.values()) == want["total"]
        assert md5(orc.tsv_bytes("Test_R1", table)) == want["tsv_md5"]
        assert sum("N" in w for w in table) == want["with_N"]


def test_chunked_protein_against_committed_tree(reference_results, tmp_path):
    """results/2023-11-29/faa-5genomes-1: k=5 c=10 -s 1 -- pins per-chunk filtering and the
    Chunker's split points (piece sizes and digests)."""
    run = reference_results["faa-5genomes-1"]
    for base, want in run["tsv"].items():
        src = tmp_path / f"{base}.faa"
        src.write_bytes(read_maybe_gz(GOLDEN / "data/faa_gz" / f"{base}.faa.gz"))
        work = tmp_path / ("w_" + base)
        files = orc.chunk_files(src, 1, work / "chunks")
        if base in run["chunks"]:
            got = [{"name": os.path.basename(f), "bytes": os.path.getsize(f),
                    "md5": md5(open(f, "rb").read())} for f in sorted(files)]
            assert got == run["chunks"][base]
        else:
            assert files == [str(src)]
        _, out = orc.run_sample(base, files, tmp_path / f"{base}.tsv", 5, 10)
        blob = open(out, "rb").read()
        assert blob.count(b"\n") - 1 == want["rows"]
        assert md5(blob) == want["tsv_md5"], base


def test_unchunked_protein_against_committed_tree(reference_results, golden_configs):
    run = reference_results["faa-5genomes-10"]
    for base, want in run["tsv"].items():
        assert golden_configs["k5_c10_unchunked"][base]["tsv_md5"] == want["tsv_md5"]
    base = "RW1_pro"
    table = orc.find_kmers(GOLDEN / "data/faa_gz" / f"{base}.faa.gz", 5, 10)
    assert md5(orc.tsv_bytes(base, table)) == run["tsv"][base]["tsv_md5"]


def test_human2bytes_and_names():
    assert orc.human2bytes("0 B") == 0 and orc.human2bytes("1 K") == 1024
    assert orc.human2bytes("1M") == 1 << 20 and orc.human2bytes("1 Gi") == 1 << 30
    assert orc.human2bytes("0.5kilo") == 512 and orc.human2bytes("1 k") == 1024
    with pytest.raises(ValueError):
        orc.human2bytes("12 foo")
    assert orc.chunk_piece_name("x/c.fna", 0) == "c.00000"
    assert orc.chunk_piece_name("x/c_clean.fna.gz", 2) == "c_clean.00002.fna"


def test_metrics_against_committed_table():
    rows = gzip.open(GOLDEN / "metrics_DJ_pro.tsv.gz", "rt").read().splitlines()[1:]
    want = {}
    for row in rows:
        header, name, length, pi, mw, hydro = row.split("\t")
        want[header] = (name, float(length), float(pi), float(mw), float(hydro))
    got = orc.sample_metrics(GOLDEN / "data/faa_gz/DJ_pro.faa.gz")
    assert len(got) == len(want) == 4852
    lengths = [r[2] for r in got]
    assert lengths == sorted(lengths, reverse=True)
    for header, name, length, pi, mw, hydro in got:
        assert want[header] == (name, length, pi, mw, hydro), header


def test_metrics_odd_sequences(capsys):
    for item in json.load(open(GOLDEN / "metrics_odd.json")):
        assert orc.predict_isoelectric_point_ProMoST(item["seq"]) == item["pI"], item
        assert orc.calculate_MW(item["seq"]) == item["MW"]
        assert orc.calculate_hydro(item["seq"]) == item["hydro"]


def _merge_inputs(tmp_path, label):
    import gzip as _gzip
    exp = GOLDEN / "expected"
    tsv_list = {}
    for f in sorted(exp.glob("*_k3_c10.tsv*")):
        pro = "_pro_" in f.name
        if pro != (label == "protein_k3"):
            continue
        name = f.name.split("_k3_")[0]
        dst = tmp_path / f"{name}_counts.tsv"
        dst.write_bytes(_gzip.open(f, "rb").read() if f.suffix == ".gz" else f.read_bytes())
        tsv_list[name] = str(dst)
    return tsv_list


@pytest.mark.parametrize("label", ["nucleotide_k3", "protein_k3"])
def test_oracle_merge_tsv_vs_reference(tmp_path, label):
    """merge_tsv restatement against the reference's own output on the committed per-sample tables
    (oracle/make_golden_merge.py); merge_tsv_T: same cells (the reference's column order is a set's)"""
    import gzip as _gzip
    tsv_list = _merge_inputs(tmp_path, label)
    assert len(tsv_list) == 5
    out = tmp_path / "combined.tsv"
    orc.merge_tsv(tsv_list, out)
    assert out.read_bytes() == _gzip.open(GOLDEN / f"merged_{label}.tsv.gz", "rb").read()
    out_t = tmp_path / "combined_T.tsv"
    orc.merge_tsv_T(tsv_list, out_t)

    def cells(blob):
        lines = blob.decode().splitlines()
        cols = lines[0].split("\t")[1:]
        return {(row.split("\t")[0], c): v for row in lines[1:] for c, v in zip(cols, row.split("\t")[1:])}
    assert cells(out_t.read_bytes()) == cells(_gzip.open(GOLDEN / f"merged_{label}_T.tsv.gz", "rb").read())
