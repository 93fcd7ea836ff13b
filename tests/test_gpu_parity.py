"""GPU parity: the CUDA path through the C ABI against the oracle and the reference's golden vectors.
Bit-exact for counts (integer/byte work); pI / MW / hydropathy compared after the reference's own
round(x, 2) and additionally within 1e-6 relative on the unrounded values."""
import gzip
import hashlib
import json
import os
import random

import numpy as np
import pytest

from conftest import GOLDEN, read_maybe_gz
from oracle import mercat2_oracle as orc

pytestmark = pytest.mark.gpu


def md5(b):
    return hashlib.md5(b).hexdigest()


@pytest.fixture(scope="module")
def engine():
    import mercat2_b200
    eng = mercat2_b200.Engine(0)
    yield eng
    eng.close()


def reset(engine):
    engine.set_option("force_path", 0)
    engine.set_option("force_encoding", -1)
    engine.set_option("batch_symbols", 1 << 28)
    engine.set_option("dense_max_bins", 1 << 24)
    engine.set_option("smem_max_bins", 32768)
    engine.set_option("sparse_algo", 0)
    engine.set_option("fast_nt", 1)
    engine.set_option("hash_bucket_keys", 3500)
    engine.set_option("parse_single", 0)
    engine.set_option("count_mode", -1)
    engine.set_option("row_merge", 0)
    engine.set_option("grid_waves", 0)
    engine.set_option("bucket_growth", 1)


def diff_msg(got, want):
    missing = sorted(set(want) - set(got))[:5]
    extra = sorted(set(got) - set(want))[:5]
    wrong = [(k, got[k], want[k]) for k in sorted(set(got) & set(want)) if got[k] != want[k]][:5]
    return f"rows got={len(got)} want={len(want)} missing={missing} extra={extra} wrong={wrong}"


def check(engine, text, k, c, want, label):
    got = engine.count_text(text, k, c).to_dict()
    assert got == want, f"{label}: {diff_msg(got, want)}"


# ---- 1. the edge-case corpus, every path --------------------------------------------------------
def test_edge_cases_auto(engine, edge_cases):
    reset(engine)
    for case in edge_cases:
        check(engine, case["text"], case["k"], case["min_count"], case["expected"],
              f"{case['name']} k={case['k']} c={case['min_count']}")


@pytest.mark.parametrize("path", [2, 3])
def test_edge_cases_forced_paths(engine, edge_cases, path):
    reset(engine)
    engine.set_option("force_path", path)
    try:
        for case in edge_cases:
            check(engine, case["text"], case["k"], case["min_count"], case["expected"],
                  f"path={path} {case['name']} k={case['k']} c={case['min_count']}")
    finally:
        reset(engine)


@pytest.mark.parametrize("algo,fast_nt", [(1, 0), (2, 0), (2, 1)])
def test_edge_cases_sparse_algorithms(engine, edge_cases, algo, fast_nt):
    """force the sparse path with the radix-sort (1) and the range-partition + shared-memory-table (2) counting
    kernels, the latter with and without the SWAR/packed nucleotide lane"""
    reset(engine)
    engine.set_option("force_path", 2)
    engine.set_option("sparse_algo", algo)
    engine.set_option("fast_nt", fast_nt)
    try:
        for case in edge_cases:
            check(engine, case["text"], case["k"], case["min_count"], case["expected"],
                  f"algo={algo} {case['name']} k={case['k']} c={case['min_count']}")
    finally:
        reset(engine)


@pytest.mark.parametrize("enc", [0, 1, 2])
def test_edge_cases_forced_encodings(engine, edge_cases, enc):
    reset(engine)
    engine.set_option("force_encoding", enc)
    try:
        for case in edge_cases:
            check(engine, case["text"], case["k"], case["min_count"], case["expected"],
                  f"enc={enc} {case['name']} k={case['k']} c={case['min_count']}")
    finally:
        reset(engine)


def test_misaligned_device_and_host_buffers(engine, edge_cases):
    import torch
    reset(engine)
    case = next(c for c in edge_cases if c["name"] == "multi_line_wrapped" and c["k"] == 12)
    for shift in (1, 3, 7, 15, 16, 17):
        host = np.frombuffer(b"#" * shift + case["text"], dtype=np.uint8)
        got = engine.count_text(host[shift:], case["k"], case["min_count"]).to_dict()
        assert got == case["expected"], f"host shift {shift}"
        dev = torch.from_numpy(host.copy()).cuda()
        got = engine.count_text(dev[shift:], case["k"], case["min_count"]).to_dict()
        assert got == case["expected"], f"device shift {shift}"


def test_non_ascii_is_rejected(engine):
    import mercat2_b200
    with pytest.raises(mercat2_b200.NonAsciiError):
        engine.count_text(">r\nAC\xc3\xa9GT\n".encode("latin-1"), 2, 1)
    with pytest.raises(mercat2_b200.Mc2Error):
        engine.count_text(b">r\nACGT\n", 0, 1)


def test_calculate_kmer_count(engine):
    from mercat2_b200 import mercat2_kmers
    reset(engine)
    rng = random.Random(5)
    for seq, k in [("ACGTACGTAC", 3), ("AC", 3), ("MKV*LA>G\nX", 2), ("acgtNNACGT" * 30, 12),
                   ("".join(rng.choice("ACGT") for _ in range(5000)), 31),
                   ("".join(rng.choice("ACDEFGHIKLMNPQRSTVWY") for _ in range(3000)), 4)]:
        assert mercat2_kmers.calculateKmerCount(seq, k) == orc.calculate_kmer_count(seq, k), (seq[:20], k)


# ---- 2. BASELINE configs --------------------------------------------------------------------------
@pytest.mark.parametrize("base", ["DJ_pro", "GIC31_pro", "RW1_pro", "RW2_pro", "Rleg_pro"])
def test_config2_protein_k3_c10(engine, base, golden_configs):
    reset(engine)
    text = read_maybe_gz(GOLDEN / "data/faa_gz" / f"{base}.faa.gz")
    table = engine.count_text(text, 3, 10)
    want = golden_configs["protein_k3_c10"][base]
    assert table.rows == want["rows"] and table.total == want["total"]
    tsv = table.tsv_bytes(base)
    assert tsv == gzip.open(GOLDEN / "expected" / f"{base}_k3_c10.tsv.gz", "rb").read()
    assert md5(tsv) == want["tsv_md5"]


@pytest.mark.parametrize("base", ["DJ", "GIC31", "RW1", "RW2", "Rleg"])
def test_config1_nucleotide_k3_c10(engine, base, golden_configs, tmp_path):
    from mercat2_b200 import mercat2_fasta, mercat2_kmers
    reset(engine)
    clean, _ = mercat2_fasta.removeN(GOLDEN / "data/fna_gz" / f"{base}.fna.gz", tmp_path / "clean", False)
    text = mercat2_kmers.read_text_bytes(clean)
    assert md5(text) == golden_configs["removeN"][base]["clean_md5"]
    table = engine.count_text(text, 3, 10)
    want = golden_configs["nucleotide_k3_c10"][base]
    assert table.rows == want["rows"] and table.total == want["total"]
    tsv = table.tsv_bytes(base)
    assert tsv == (GOLDEN / "expected" / f"{base}_k3_c10.tsv").read_bytes()
    # the same file through the other counting paths
    for path in (2, 3):
        engine.set_option("force_path", path)
        try:
            if path == 3 and base not in ("RW1",):
                continue
            assert md5(engine.count_text(text, 3, 10).tsv_bytes(base)) == want["tsv_md5"], f"path {path}"
        finally:
            reset(engine)


def test_config3_test_r1_k12(engine, golden_configs, tmp_path):
    from mercat2_b200 import mercat2_fasta, mercat2_kmers
    reset(engine)
    fna = mercat2_fasta.fq2fa(str(GOLDEN / "data/Test_R1.fastq.gz"), str(tmp_path / "clean"), "Test_R1")
    text = mercat2_kmers.read_text_bytes(fna)
    for force in (0, 2, 3):
        engine.set_option("force_path", force)
        try:
            for c in (1, 2, 10):
                table = engine.count_text(text, 12, c)
                want = golden_configs["test_r1_k12"][f"c{c}"]
                assert table.rows == want["rows"], (force, c)
                assert table.total == want["total"], (force, c)
                assert md5(table.tsv_bytes("Test_R1")) == want["tsv_md5"], (force, c)
                if c == 2:
                    assert table.tsv_bytes("Test_R1") == (GOLDEN / "expected/Test_R1_k12_c2.tsv").read_bytes()
        finally:
            reset(engine)
    # drop-in entry point + "no TSV when nothing survives"
    d = mercat2_kmers.find_kmers(__import__("pathlib").Path(fna), 12, 1)
    assert len(d) == golden_configs["test_r1_k12"]["c1"]["rows"] and sum("N" in w for w in d) == 19
    assert engine.count_text(text, 12, 10).write_tsv(tmp_path / "none.tsv", "Test_R1") is False
    assert not (tmp_path / "none.tsv").exists()


# ---- 3. the reference's committed result tree: k=5 c=10, -s 1 and -s 10 ------------------------------
@pytest.mark.parametrize("base", ["DJ_pro", "GIC31_pro", "RW1_pro", "RW2_pro", "Rleg_pro"])
def test_protein_k5_chunked_and_unchunked(engine, base, reference_results, tmp_path):
    from mercat2_b200 import pipeline
    reset(engine)
    text = read_maybe_gz(GOLDEN / "data/faa_gz" / f"{base}.faa.gz")
    src = tmp_path / f"{base}.faa"
    src.write_bytes(text)
    # -s 10: no file reaches 10 MiB -> one piece
    want = reference_results["faa-5genomes-10"]["tsv"][base]
    name, out = pipeline.run_mercat2(base, [src], tmp_path / "u.tsv", 5, 10, chunk_size_mb=10, engine=engine, quiet=True)
    blob = open(out, "rb").read()
    assert blob.count(b"\n") - 1 == want["rows"] and md5(blob) == want["tsv_md5"]
    # -s 1: virtual pieces must equal the reference's piece files, and the per-piece filter applies
    run = reference_results["faa-5genomes-1"]
    chunk_bytes = pipeline.chunk_trigger(src, 1)
    offsets = engine.chunk_offsets(text, chunk_bytes) if chunk_bytes else [0]
    sizes = [b - a for a, b in zip(offsets, offsets[1:] + [len(text)])]
    want_sizes = [p["bytes"] for p in run["chunks"].get(base, [{"bytes": len(text)}])]
    assert sizes == want_sizes
    for (a, n), piece in zip(zip(offsets, sizes), run["chunks"].get(base, [])):
        assert md5(text[a:a + n]) == piece["md5"]
    name, out = pipeline.run_mercat2(base, [src], tmp_path / "c.tsv", 5, 10, chunk_size_mb=1, engine=engine, quiet=True)
    blob = open(out, "rb").read()
    want = run["tsv"][base]
    assert blob.count(b"\n") - 1 == want["rows"] and md5(blob) == want["tsv_md5"]


@pytest.mark.parametrize("base", ["RW1", "Rleg", "DJ"])
def test_nucleotide_k5_gz_chunked(engine, base, reference_results, tmp_path):
    from mercat2_b200 import mercat2_fasta, pipeline
    reset(engine)
    clean, _ = mercat2_fasta.removeN(GOLDEN / "data/fna_gz" / f"{base}.fna.gz", tmp_path / "clean", False)
    for s in (1, 10):
        want = reference_results[f"fna-5genomes_gz-{s}"]["tsv"][base]
        name, out = pipeline.run_mercat2(base, [clean], tmp_path / f"{s}.tsv", 5, 10, chunk_size_mb=s, engine=engine, quiet=True)
        blob = open(out, "rb").read()
        assert blob.count(b"\n") - 1 == want["rows"] and md5(blob) == want["tsv_md5"], s
    if base == "Rleg":
        text = gzip.open(clean, "rb").read()
        pieces = reference_results["fna-5genomes_gz-1"]["chunks"]["Rleg"]
        offsets = engine.chunk_offsets(text, 1 << 20)
        sizes = [b - a for a, b in zip(offsets, offsets[1:] + [len(text)])]
        # the committed tree lacks the first piece (Rleg_clean.00000.fna, a large blob): compare the rest
        assert [p["name"] for p in pieces] == ["Rleg_clean.%05d.fna" % i for i in (1, 2, 3)]
        assert sizes[1:] == [p["bytes"] for p in pieces] and sum(sizes) == len(text)
        import hashlib
        for a, n, piece in zip(offsets[1:], sizes[1:], pieces):
            assert hashlib.md5(text[a:a + n]).hexdigest() == piece["md5"]


# ---- 4. synthetic data against the oracle -----------------------------------------------------------
def synth_reads(n_reads, read_len, seed, n_rate=0.0, lower_rate=0.0, genome_len=200000):
    rng = np.random.default_rng(seed)
    genome = rng.integers(0, 4, genome_len, dtype=np.uint8)
    starts = rng.integers(0, genome_len - read_len, n_reads)
    lut = np.frombuffer(b"ACGT", dtype=np.uint8)
    lines = []
    for i, s in enumerate(starts):
        seq = lut[genome[s:s + read_len]].copy()
        if n_rate:
            seq[rng.random(read_len) < n_rate] = ord("N")
        if lower_rate and rng.random() < lower_rate:
            seq = np.frombuffer(seq.tobytes().lower(), dtype=np.uint8)
        lines.append(b">r%010d\n" % i + seq.tobytes() + b"\n")
    return b"".join(lines)


@pytest.mark.parametrize("k,c", [(31, 1), (31, 2), (32, 2), (21, 3), (13, 2), (12, 2), (33, 2), (40, 1)])
def test_synthetic_reads_vs_oracle(engine, k, c):
    reset(engine)
    text = synth_reads(6000, 150, seed=k * 7 + c, n_rate=0.002, lower_rate=0.01)
    want = orc.find_kmers_text(text.decode(), k, c)
    check(engine, text, k, c, want, f"reads k={k} c={c}")
    if k <= 32:
        engine.set_option("batch_symbols", 65536)          # several batches per chunk
        try:
            check(engine, text, k, c, want, f"reads batched k={k} c={c}")
        finally:
            reset(engine)
        for algo, bucket_keys, fast_nt in ((1, 7000, 0), (2, 7000, 0), (2, 64, 0), (2, 1000000, 0),
                                           (2, 7000, 1), (2, 64, 1), (2, 1000000, 1)):
            # 64 keys/bucket: many level-1 buckets; 10^6: every table overflows -> sort fallback
            engine.set_option("force_path", 2)
            engine.set_option("sparse_algo", algo)
            engine.set_option("hash_bucket_keys", bucket_keys)
            engine.set_option("fast_nt", fast_nt)
            try:
                check(engine, text, k, c, want, f"reads algo={algo} bucket_keys={bucket_keys} fast_nt={fast_nt} k={k} c={c}")
            finally:
                reset(engine)


@pytest.mark.parametrize("k", [8, 9, 12, 17, 33])
def test_wide_path_thresholds(engine, k):
    """Regression: the wide path's run-length threshold kernels follow a data-dependent byte-compare
    loop; block collectives after it need an explicit warp sync (see common.cuh)."""
    reset(engine)
    text = synth_reads(6000, 150, seed=238, n_rate=0.002, lower_rate=0.01)
    engine.set_option("force_path", 3)
    try:
        for c in (1, 2, 3, 7):
            want = orc.find_kmers_text(text.decode(), k, c)
            check(engine, text, k, c, want, f"wide k={k} c={c}")
    finally:
        reset(engine)


@pytest.mark.parametrize("k,c", [(3, 10), (5, 2), (6, 2), (12, 2), (13, 1)])
def test_synthetic_protein_vs_oracle(engine, k, c):
    reset(engine)
    rng = np.random.default_rng(100 + k)
    letters = np.frombuffer(b"ACDEFGHIKLMNPQRSTVWYXBZUO", dtype=np.uint8)
    recs = []
    for i in range(800):
        n = int(rng.integers(20, 900))
        seq = letters[rng.integers(0, 20 if i % 50 else 25, n)].tobytes()
        body = b"\n".join(seq[j:j + 60] for j in range(0, n, 60))
        recs.append(b">p%d some protein\n" % i + body + b"*\n")
    text = b"".join(recs)
    want = orc.find_kmers_text(text.decode(), k, c)
    check(engine, text, k, c, want, f"protein k={k} c={c}")
    for path in (2, 3):
        engine.set_option("force_path", path)
        engine.set_option("batch_symbols", 32768)
        try:
            check(engine, text, k, c, want, f"protein path={path} k={k} c={c}")
        finally:
            reset(engine)


@pytest.mark.parametrize("line_end", [b"\n", b"\r\n", b"\r"])
def test_fast_lane_wrapped_genome(engine, line_end):
    """multi-line records (windows span line ends), every line-end flavour, odd line widths, long headers,
    a record whose lines put '>' / 'N' / lower case inside the sequence, through the packed lane"""
    reset(engine)
    rng = np.random.default_rng(77)
    lut = np.frombuffer(b"ACGT", dtype=np.uint8)
    recs = []
    for i, width in enumerate((60, 61, 70, 80, 17, 16, 15, 1, 33, 4096, 5000)):
        n = int(rng.integers(3000, 9000))
        seq = bytearray(lut[rng.integers(0, 4, n)].tobytes())
        if i % 3 == 1:
            for pos in rng.integers(0, n, 6):
                seq[pos] = ord("N")
        if i % 4 == 2:
            seq[100:130] = seq[100:130].lower()
        if i == 5:
            seq[500] = ord(">")
        body = line_end.join(bytes(seq[j:j + width]) for j in range(0, n, width))
        recs.append(b">contig_%d some description with > inside" % i + b"x" * (i * 37) + line_end + body + line_end)
    text = b"".join(recs)
    for k, c in ((31, 2), (32, 2), (13, 3), (2, 2), (1, 2)):
        want = orc.find_kmers_text(text.decode(), k, c)
        for fast, single in ((1, 1), (1, 0), (0, 1)):          # packed lane with the one-pass and the two-pass parse, general lane
            engine.set_option("force_path", 2)
            engine.set_option("fast_nt", fast)
            engine.set_option("parse_single", single)
            try:
                check(engine, text, k, c, want, f"genome le={line_end!r} k={k} c={c} fast={fast} single={single}")
            finally:
                reset(engine)


def test_virtual_chunking_vs_oracle(engine, tmp_path):
    """Chunker split points + per-piece filter + sum on synthetic reads with CRLF line ends."""
    reset(engine)
    text = synth_reads(20000, 100, seed=11, genome_len=3000).replace(b"\n", b"\r\n")
    src = tmp_path / "reads.fna"
    src.write_bytes(text)
    for chunk in (1 << 20, 300000, 100000):
        pieces = orc.chunker_pieces(open(src, "r"), chunk)
        sizes = []
        want = orc.merge_counts(orc.find_kmers_text("".join(p), 9, 5) for p in pieces)
        # raw byte size of each piece: every line regains its "\r"
        raw_sizes = [sum(len(l) + 1 for l in p) for p in pieces]
        offsets = engine.chunk_offsets(text, chunk)
        got_sizes = [b - a for a, b in zip(offsets, offsets[1:] + [len(text)])]
        assert got_sizes == raw_sizes, chunk
        table, offs2 = engine.count_sample(text, 9, 5, chunk)
        assert offs2 == offsets
        got = table.to_dict()
        assert got == want, f"chunk={chunk}: {diff_msg(got, want)}"


def test_streaming_upload_equals_resident(engine):
    """host text (pageable and pinned) goes through the pipelined upload + windowed boundary search; the device-
    resident text through the whole-text boundary search: same pieces, same table"""
    import torch
    import bench
    reset(engine)
    dev = torch.device("cuda", 0)
    genomes = bench.make_genomes(dev, 0.002)
    text = bench.make_reads_text(dev, genomes, 500_000, 0)          # 82 MB, ~4x coverage: plenty of survivors
    chunk = 16 << 20
    t_dev, offs_dev = engine.count_sample(text, 25, 3, chunk)
    k_dev, c_dev = t_dev.arrays()
    assert len(offs_dev) == 5 and len(c_dev) > 1000
    host = text.cpu()
    for name, buf in (("pageable", host.numpy()), ("pinned", host.pin_memory())):
        t_h, offs_h = engine.count_sample(buf, 25, 3, chunk)
        k_h, c_h = t_h.arrays()
        assert offs_h == offs_dev, name
        assert (k_h == k_dev).all() and (c_h == c_dev).all(), name
    # and against the oracle on a prefix small enough for it
    small = host.numpy()[:164 * 20000].tobytes()
    want = orc.merge_counts(orc.find_kmers_text("".join(p), 25, 2)
                            for p in orc.chunker_pieces(small.decode().splitlines(True), 1 << 20))
    got = engine.count_sample(small, 25, 2, 1 << 20)[0].to_dict()
    assert got == want, diff_msg(got, want)


def test_large_dense_and_sparse_properties(engine):
    """Size-independent properties at a larger size: sum of counts == number of windows; forcing
    another path gives the same table; c filter is monotone."""
    reset(engine)
    n_reads, L = 200000, 150
    text = synth_reads(n_reads, L, seed=3, genome_len=2_000_000)
    for k in (4, 12, 31):
        table = engine.count_text(text, k, 1)
        assert table.total == n_reads * (L - k + 1), k
        kmers, counts = table.arrays()
        order = np.lexsort(kmers.T[::-1])
        assert (order == np.arange(len(order))).all(), "rows must be sorted by k-mer text"
        engine.set_option("force_path", 2)
        try:
            t2 = engine.count_text(text, k, 1)
            k2, c2 = t2.arrays()
            assert (k2 == kmers).all() and (c2 == counts).all()
        finally:
            reset(engine)
        for algo in (1, 2):
            engine.set_option("sparse_algo", algo)
            engine.set_option("force_path", 2)
            try:
                t3 = engine.count_text(text, k, 3)
                k3, c3 = t3.arrays()
                keep = counts >= 3
                assert (k3 == kmers[keep]).all() and (c3 == counts[keep]).all(), (k, algo)
            finally:
                reset(engine)


# ---- 5. protein metrics -----------------------------------------------------------------------------
def test_metrics_golden_table(engine):
    from mercat2_b200 import mercat2_metrics
    rows = gzip.open(GOLDEN / "metrics_DJ_pro.tsv.gz", "rt").read().splitlines()[1:]
    want = {}
    for row in rows:
        header, name, length, pi, mw, hydro = row.split("\t")
        want[header] = (name, float(length), float(pi), float(mw), float(hydro))
    got = mercat2_metrics.file_metrics(GOLDEN / "data/faa_gz/DJ_pro.faa.gz", engine)
    assert len(got) == len(want) == 4852
    lengths = [r[2] for r in got]
    assert lengths == sorted(lengths, reverse=True)
    bad = [(h, (n, l, pi, mw, hy), want[h]) for h, n, l, pi, mw, hy in got if want[h] != (n, l, pi, mw, hy)]
    assert not bad, bad[:3]


def test_metrics_unrounded_vs_oracle(engine):
    text = read_maybe_gz(GOLDEN / "data/faa_gz/RW1_pro.faa.gz")
    res = engine.protein_metrics(text)
    recs = [(n, s) for n, s in orc.protein_records(text.decode().splitlines(True)) if s]
    assert len(recs) == len(res["length"])
    for i, (name, seq) in enumerate(recs):
        off, ln = int(res["header_off"][i]), int(res["header_len"][i])
        assert text[off:off + ln].decode() == name
        assert int(res["length"][i]) == len(seq)
        assert res["pi"][i] == orc.isoelectric_point_unrounded(seq)          # same bisection path
        assert abs(res["mw"][i] - orc.molecular_weight_unrounded(seq)) <= 1e-6 * abs(res["mw"][i])
        assert abs(res["hydro"][i] - orc.hydropathy_unrounded(seq)) <= 1e-6 * max(1.0, abs(res["hydro"][i]))


def test_metrics_scalar_entry_points(engine, capsys):
    from mercat2_b200 import mercat2_metrics as mm
    for item in json.load(open(GOLDEN / "metrics_odd.json")):
        assert mm.predict_isoelectric_point_ProMoST(item["seq"]) == item["pI"], item
        assert mm.calculate_MW(item["seq"]) == item["MW"], item
        assert mm.calculate_hydro(item["seq"]) == item["hydro"], item
    with pytest.raises(KeyError):
        mm.predict_isoelectric_point_ProMoST("*MK")
    with pytest.raises(IndexError):
        mm.predict_isoelectric_point_ProMoST("")
    odd = ">a x\n  MK*L*  \n*\nAB *\n>b\n\n>c\nMKV\r\n"
    got = mm.file_metrics_text(odd.encode(), engine)
    want = []
    for name, seq in orc.protein_records(odd.splitlines(True)):
        if seq:
            want.append((name, name.split()[0], float(len(seq)), orc.predict_isoelectric_point_ProMoST(seq),
                         orc.calculate_MW(seq), orc.calculate_hydro(seq)))
    assert sorted(got) == sorted(want)


# ---- 6. the command line and the multi-rank merge ------------------------------------------------------
def test_cli_end_to_end(golden_configs, reference_results, tmp_path):
    """bin/mercat2.py -i <5 proteomes> -k 5 -c 10 -s 1 reproduces the reference's committed TSVs; and the
    nucleotide flow (removeN -> clean/*.fna.gz -> count) reproduces config 1."""
    import subprocess, sys
    from conftest import ROOT
    inputs = []
    for base in ("RW1_pro", "GIC31_pro"):
        src = tmp_path / f"{base}.faa"
        src.write_bytes(read_maybe_gz(GOLDEN / "data/faa_gz" / f"{base}.faa.gz"))
        inputs.append(str(src))
    out = tmp_path / "out"
    proc = subprocess.run([sys.executable, str(ROOT / "bin/mercat2.py"), "-i", *inputs, "-k", "5", "-c", "10", "-s", "1",
                           "-o", str(out)], capture_output=True, text=True)
    assert proc.returncode == 0, proc.stderr[-2000:]
    assert "Significant k-mers:" in proc.stdout
    for base in ("RW1_pro", "GIC31_pro"):
        blob = (out / "tsv_protein" / f"{base}_counts.tsv").read_bytes()
        assert md5(blob) == reference_results["faa-5genomes-1"]["tsv"][base]["tsv_md5"]
    assert (out / "report" / "metrics-protein.tsv").read_text().count("\n") > 300
    out2 = tmp_path / "out2"
    proc = subprocess.run([sys.executable, str(ROOT / "bin/mercat2.py"), "-i", str(GOLDEN / "data/fna_gz/RW1.fna.gz"),
                           "-k", "3", "-o", str(out2)], capture_output=True, text=True)
    assert proc.returncode == 0, proc.stderr[-2000:]
    assert md5((out2 / "tsv_nucleotide" / "RW1_counts.tsv").read_bytes()) == golden_configs["nucleotide_k3_c10"]["RW1"]["tsv_md5"]
    assert (out2 / "clean" / "RW1_clean.fna.gz").exists()


def test_engine_reducer_merges_rank_tables(engine):
    """two 'ranks' count disjoint sets of pieces; the engine reducer (device sort + reduce-by-key over added
    rows) must give the reference's merged table"""
    from mercat2_b200 import distributed as mcd
    reset(engine)
    text = synth_reads(4000, 150, seed=9, n_rate=0.003, lower_rate=0.02, genome_len=60000)
    reads = [b">" + r for r in text.split(b">") if r]
    halves = [b"".join(reads[0::2]), b"".join(reads[1::2])]
    for k, c in ((21, 2), (33, 2), (4, 5)):
        tables = [engine.count_text(h, k, c).arrays() for h in halves]
        mk, mc = mcd.engine_reducer(engine)([t[0] for t in tables], [t[1] for t in tables], k)
        got = {bytes(r).decode(): int(n) for r, n in zip(mk, mc)}
        want = orc.merge_counts(orc.find_kmers_text(h.decode(), k, c) for h in halves)
        assert got == want, diff_msg(got, want)
        assert mcd.tsv_bytes("s", mk, mc) == orc.tsv_bytes("s", want)


# ---- device-resident exchange: NCCL process group of size 1 on this GPU (the 2-GPU run is tools/sharded_check.py) ----
def test_sharded_sample_device_exchange(engine, tmp_path):
    """count_sample_sharded: local pieces -> per-sample merge on the device -> TSV written by byte ranges.  With one
    rank the all-to-all is a self copy, but every device call of the exchange (row pointers, splitter cuts, rebuild
    from packed + literal rows, TSV body, dense reduce) runs."""
    import torch
    import torch.distributed as dist
    from mercat2_b200 import distributed as mcd
    reset(engine)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", str(29500 + os.getpid() % 2000))
    created = not dist.is_initialized()
    if created:
        dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
    try:
        dev = torch.device("cuda", 0)
        text = synth_reads(3000, 150, seed=21, n_rate=0.003, lower_rate=0.02, genome_len=50000)
        reads = [b">" + r for r in text.split(b">") if r]
        pieces = [b"".join(reads[i::3]) for i in range(3)]
        for k, c in ((21, 2), (4, 5), (33, 2), (12, 1)):
            out = tmp_path / f"s_{k}_{c}.tsv"
            part = mcd.count_sample_sharded(engine, pieces, k, c, dist, dev, out_path=out, basename="s")
            want = orc.merge_counts(orc.find_kmers_text(p.decode(), k, c) for p in pieces)
            assert part is not None and part.to_dict() == want, diff_msg(part.to_dict(), want)
            if want:
                assert out.read_bytes() == orc.tsv_bytes("s", want)
            else:
                assert not out.exists()
        # one piece split by position, keys exchanged before the filter (whole-piece counts)
        whole = b"".join(reads)
        assert b"".join(mcd.split_at_headers(whole, 3)) == whole
        for k, c in ((21, 2), (31, 3), (12, 4)):
            out = tmp_path / f"p_{k}_{c}.tsv"
            part = mcd.count_piece_position_sharded(engine, whole, k, c, dist, dev, out_path=out, basename="s")
            want = orc.find_kmers_text(whole.decode(), k, c)
            assert part.to_dict() == want, diff_msg(part.to_dict(), want)
            assert out.read_bytes() == orc.tsv_bytes("s", want)
        engine.set_option("hash_bucket_keys", 4)          # received keys exceed one hash batch: local level-0 partition
        try:
            part = mcd.count_piece_position_sharded(engine, whole, 21, 2, dist, dev)
            want = orc.find_kmers_text(whole.decode(), 21, 2)
            assert part.to_dict() == want, diff_msg(part.to_dict(), want)
        finally:
            reset(engine)
    finally:
        if created:
            dist.destroy_process_group()


@pytest.mark.parametrize("k,c", [(21, 2), (31, 3), (9, 2)])
def test_big_chunk_level0_partition(engine, k, c):
    """a chunk with more windows than one hash batch holds: parsed in spans cut at header lines, keys partitioned
    once into level-0 groups, each group counted by the two-level pipeline; the -c filter applies to the whole chunk"""
    reset(engine)
    text = synth_reads(5000, 150, seed=31, n_rate=0.002, lower_rate=0.0, genome_len=30000)
    want = orc.find_kmers_text(text.decode(), k, c)
    engine.set_option("hash_bucket_keys", 4)          # one batch = 384 * 128 * 4 = 196 608 keys
    engine.set_option("span_bytes", 64 << 10)
    engine.set_option("force_path", 2)
    try:
        launches0 = engine.stat("launches")
        check(engine, text, k, c, want, f"big chunk k={k} c={c}")
        engine.set_option("profile", 2)
        engine.count_text(text, k, c).close()
        prof = engine.profile()
        engine.set_option("profile", 0)
        assert "hk_scatter1_kernel" in prof and prof["hk_scatter1_kernel"]["launches"] >= 2, sorted(prof)
        assert engine.stat("launches") > launches0
        # the same chunk with the level-0 path switched off (sort fallback) gives the same table
        engine.set_option("big_chunks", 0)
        check(engine, text, k, c, want, f"sort fallback k={k} c={c}")
    finally:
        engine.set_option("big_chunks", 1)
        engine.set_option("span_bytes", 1 << 30)
        reset(engine)


def test_streaming_file_reader(engine, tmp_path):
    """mc2_sample_add_file (reader thread -> pinned buffers -> device window) against add_text of the same bytes:
    plain and gzip files, chunked and unchunked, an empty file, a missing file"""
    import mercat2_b200
    reset(engine)
    text = synth_reads(30000, 150, seed=41, n_rate=0.001, lower_rate=0.0, genome_len=200000)      # ~4.9 MB
    plain = tmp_path / "reads.fna"
    plain.write_bytes(text)
    gz = tmp_path / "reads.fna.gz"
    with gzip.open(gz, "wb", compresslevel=1) as f:
        f.write(text)
    for k, c, chunk in ((21, 2, 0), (21, 2, 1 << 20), (5, 3, 1 << 20), (31, 1, 700000)):
        s0 = engine.sample(k, c)
        pieces0 = s0.add_text(text, chunk)
        want = s0.finish().to_dict()
        for path in (plain, gz):
            s1 = engine.sample(k, c)
            pieces1 = s1.add_file(path, chunk)
            assert s1.text_bytes == len(text)
            got = s1.finish().to_dict()
            assert pieces1 == pieces0, (path.name, k, c, chunk, pieces1, pieces0)
            assert got == want, f"{path.name} k={k} c={c} chunk={chunk}: {diff_msg(got, want)}"
    # small pieces: the same files now span ~75 pieces, the device window is compacted and (unchunked) grown
    engine.set_option("file_piece_bytes", 64 << 10)
    try:
        for k, c, chunk in ((21, 2, 0), (21, 2, 300000), (31, 1, 1 << 20)):
            s0 = engine.sample(k, c)
            pieces0 = s0.add_text(text, chunk)
            want = s0.finish().to_dict()
            for path in (plain, gz):
                s1 = engine.sample(k, c)
                pieces1 = s1.add_file(path, chunk)
                got = s1.finish().to_dict()
                assert pieces1 == pieces0 and s1.text_bytes == len(text)
                assert got == want, f"small pieces {path.name} k={k} c={c} chunk={chunk}: {diff_msg(got, want)}"
    finally:
        engine.set_option("file_piece_bytes", 32 << 20)
    empty = tmp_path / "empty.fna"
    empty.write_bytes(b"")
    s2 = engine.sample(5, 1)
    assert s2.add_file(empty, 0) == 1 and s2.finish().rows == 0
    s3 = engine.sample(5, 1)
    with pytest.raises(mercat2_b200.Mc2Error):
        s3.add_file(tmp_path / "missing.fna", 0)
    bad = tmp_path / "broken.fna.gz"
    bad.write_bytes(gz.read_bytes()[: gz.stat().st_size // 2])
    s4 = engine.sample(5, 1)
    with pytest.raises(mercat2_b200.Mc2Error):
        s4.add_file(bad, 0)


def test_inputs_beyond_4_gib():
    """one chunk of more than 2^32 bytes on each lane (packed dense with the 64-bit fold, level-0 partition, general
    parser): sum of counts = number of windows, all 4^3 / 20^3 rows present and sorted (tools/big_sanity.py)"""
    import importlib.util
    import torch
    if torch.cuda.mem_get_info()[0] < (60 << 30):
        pytest.skip("needs 60 GB of free device memory")
    spec = importlib.util.spec_from_file_location("big_sanity", os.path.join(os.path.dirname(__file__), "..", "tools", "big_sanity.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert mod.main() == 0


def test_sample_metric_rows_orig_and_pieces(engine, tmp_path):
    """metrics-protein.tsv lists every file of [orig] + chunk_files(orig): the reference's committed table holds the
    4852 proteins of DJ_pro twice, each block sorted by length; with chunking the second block is the pieces"""
    from mercat2_b200 import pipeline, mercat2_metrics
    src = GOLDEN / "data/faa_gz/DJ_pro.faa.gz"
    once = mercat2_metrics.file_metrics(src, engine)
    rows = list(pipeline.sample_metric_rows(src, 100, engine))              # below the trigger: [orig, orig]
    assert len(rows) == 2 * 4852 and rows[:4852] == once and rows[4852:] == once
    plain = tmp_path / "DJ_pro.faa"
    plain.write_bytes(read_maybe_gz(src))
    rows = list(pipeline.sample_metric_rows(plain, 1, engine))              # 1 MiB pieces
    assert rows[:4852] == once
    tail = rows[4852:]
    assert sorted(tail) == sorted(once)
    lengths = [r[2] for r in tail]
    runs = 1 + sum(1 for a, b in zip(lengths, lengths[1:]) if b > a)
    pieces = len(engine.chunk_offsets(plain.read_bytes(), 1 << 20))
    assert pieces >= 2 and runs == pieces


@pytest.mark.parametrize("label", ["nucleotide_k3", "protein_k3"])
def test_merge_tsv_vs_reference(engine, tmp_path, label):
    """merge_tsv / merge_tsv_T on the engine (device TSV parse, union, matrix, text) against the reference's own
    merged tables; and the same matrix from tables that never left the device"""
    from mercat2_b200 import mercat2_report
    exp = GOLDEN / "expected"
    tsv_list, sources = {}, {}
    for f in sorted(exp.glob("*_k3_c10.tsv*")):
        if ("_pro_" in f.name) != (label == "protein_k3"):
            continue
        name = f.name.split("_k3_")[0]
        dst = tmp_path / f"{name}_counts.tsv"
        dst.write_bytes(read_maybe_gz(f))
        tsv_list[name] = str(dst)
    # default: the reference's own table byte for byte (its k-way merge repeats / reorders labels when the samples' k-mer
    # sets differ: protein k=3); union=True: the sorted union, equal to the reference's table when all samples hold the
    # same k-mers (nucleotide k=3)
    ref = gzip.open(GOLDEN / f"merged_{label}.tsv.gz", "rb").read()
    union = tmp_path / "union.tsv"
    orc.merge_tsv_union(tsv_list, union)
    want = union.read_bytes()
    assert (want == ref) == (label == "nucleotide_k3")
    out = tmp_path / "combined.tsv"
    mercat2_report.merge_tsv(tsv_list, out, engine)
    assert out.read_bytes() == ref
    mercat2_report.merge_tsv(tsv_list, out, engine, union=True)
    assert out.read_bytes() == want
    out_t = tmp_path / "combined_T.tsv"
    mercat2_report.merge_tsv_T(tsv_list, out_t, engine)
    ref_t = tmp_path / "ref_T.tsv"
    orc.merge_tsv_T(tsv_list, ref_t)                      # (sorted columns, like the engine)
    assert out_t.read_bytes() == ref_t.read_bytes()
    # resident tables: count the samples again and merge without touching the TSVs
    reset(engine)
    kind = "faa_gz" if label == "protein_k3" else "fna_gz"
    tables = {}
    for name in tsv_list:
        src = next((GOLDEN / "data" / kind).glob(name + ".*"))
        tables[name] = engine.count_text(read_maybe_gz(src), 3, 10)
    if label == "protein_k3":                              # (the nucleotide goldens were counted on removeN-cleaned files)
        out2 = tmp_path / "combined_resident.tsv"
        mercat2_report.merge_tables(tables, out2, "k-mer", engine)
        assert out2.read_bytes() == ref
        mercat2_report.merge_tables(tables, out2, "k-mer", engine, union=True)
        assert out2.read_bytes() == want
    # a k-mer missing from some samples, literal-byte rows, a file without trailing newline
    a = tmp_path / "a.tsv"; a.write_bytes(b"k-mer\ta_Count\nAAC\t5\nACN\t7\nTTT\t1")
    b = tmp_path / "b.tsv"; b.write_bytes(b"k-mer\tb_Count\nAAC\t2\nGGG\t123456789012\n")
    out3 = tmp_path / "ab.tsv"
    mercat2_report.merge_tsv({"b": str(b), "a": str(a)}, out3, engine, union=True)
    assert out3.read_bytes() == b"k-mer\ta\tb\nAAC\t5\t2\nACN\t7\t0\nGGG\t0\t123456789012\nTTT\t1\t0\n"
    mercat2_report.merge_tsv({"b": str(b), "a": str(a)}, out3, engine)          # the reference's walk: label = next key of the files that advanced
    ref3 = tmp_path / "ab_ref.tsv"
    orc.merge_tsv({"b": str(b), "a": str(a)}, ref3)
    assert out3.read_bytes() == ref3.read_bytes()


@pytest.mark.parametrize("option,value", [("count_mode", 0), ("count_mode", 1), ("parse_single", 1), ("prefetch_pass", 0)])
def test_experiment_variants_stay_exact(engine, option, value):
    """the kernel variants kept behind engine options (measured, not the default -- DESIGN.md §4) count the same tables"""
    reset(engine)
    text = synth_reads(6000, 150, seed=55, n_rate=0.002, lower_rate=0.01, genome_len=40000)
    pieces = 3
    want = None
    for setting in (None, value):
        if setting is not None:
            engine.set_option(option, setting)
        try:
            table, offs = engine.count_sample(text, 25, 2, len(text) // pieces)
            got = table.to_dict()
        finally:
            defaults = {"count_mode": -1, "parse_single": 0, "prefetch_pass": 1}
            engine.set_option(option, defaults[option])
        if want is None:
            want = got
            ref = orc.merge_counts(orc.find_kmers_text(text[a:b].decode(), 25, 2)
                                   for a, b in zip(offs, list(offs[1:]) + [len(text)]))
            assert got == ref, diff_msg(got, ref)
        else:
            assert got == want, f"{option}={value}: {diff_msg(got, want)}"


def test_bench_json_contract():
    """bench.py prints ONE JSON line with the keys the driver reads (both arms), on a small shard"""
    import subprocess
    import sys
    root = os.path.join(os.path.dirname(__file__), "..")
    common = ["--reads-per-gpu", "700000", "--steps", "2", "--warmup", "3", "--cpu-reads", "20000"]
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), *common], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
                "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert key in line, key
    assert line["n_gpus"] == 1 and line["steps"] == 2 and line["warmup"] == 3 and line["value"] > 0 and line["gpu_launches"] > 0
    assert "workload" in line["config"] and line["vs_baseline"] is None and line["dtype"] == "u64"
    for key in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert key in line["roofline"], key
    for key in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert key in line["e2e"], key
    assert line["e2e"]["h2d_bytes_per_step"] == 700000 * 164 and line["e2e"]["value"] > 0
    # the headline is config 4 as written: one global table (-s 0 -c 2) with survivors, the table read back, checked first
    assert "-c 2 -s 0" in line["config"]["workload"] and line["config"]["surviving_rows"] > 0
    assert line["e2e"]["d2h_bytes_per_step"] == 12 * line["config"]["surviving_rows"]      # uint64 key + uint32 count per row
    assert "12 B per row" in line["e2e"]["rows_format"]
    assert line["check"]["ok"] and line["check"]["oracle_md5"] == line["check"]["engine_md5"] and line["check"]["rows"] > 0
    assert set(("parse_ms", "partition_ms", "count_ms", "emit_ms")) <= set(line["phases"]) and line["limiting_phase"] in line["phases"]
    assert "-c 10 -s 100" in line["secondary"]["workload"] and line["secondary"]["value"] > 0
    cfg5 = line["secondary"]["cfg5"]                                   # BASELINE config 5: 64 proteomes, one table + metrics each
    assert "error" not in cfg5, cfg5
    assert cfg5["samples"] == 64 and cfg5["proteins"] == 64 * 5000 and cfg5["rows"] > 0 and cfg5["samples_per_s"] > 0
    for key in ("value", "unit", "cores", "kind", "sample"):
        assert key in line["cpu_baseline"], key
    ref = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--cpu-reads", "20000"], capture_output=True, text=True, timeout=600)
    assert ref.returncode == 0, ref.stderr[-2000:]
    rline = json.loads(ref.stdout.strip().splitlines()[-1])
    assert rline["impl"] == "reference" and rline["metric"] == line["metric"] and rline["unit"] == line["unit"]
    assert rline["cpu_baseline"]["kind"] in ("port", "reference") and rline["e2e"]["h2d_bytes_per_step"] == 0


def test_c_example_end_to_end(tmp_path):
    """the C client of examples/count_file.c produces the reference's TSV for config 1 (k=3, -c 10) from the .gz file"""
    import shutil
    import subprocess
    from pathlib import Path
    root = Path(__file__).resolve().parents[1]
    gcc = shutil.which("gcc")
    lib = root / "mercat2_b200" / "libmercat2_b200.so"
    if not gcc:
        pytest.skip("no gcc")
    exe = tmp_path / "count_file"
    subprocess.run([gcc, "-std=c99", f"-I{root / 'include'}", str(root / "examples" / "count_file.c"), f"-L{lib.parent}",
                    "-lmercat2_b200", f"-Wl,-rpath,{lib.parent}", "-o", str(exe)], check=True)
    src = GOLDEN / "data/fna_gz/DJ.fna.gz"
    out = tmp_path / "DJ_counts.tsv"
    run = subprocess.run([str(exe), str(src), "3", "10", "0", str(out)], capture_output=True, text=True)
    assert run.returncode == 0, run.stderr
    want = orc.tsv_bytes("sample", orc.find_kmers(src, 3, 10))
    assert out.read_bytes() == want and "Significant k-mers: 64" in run.stdout


# ---- full-piece known answers on the range path (no oracle: numpy computes every count) ---------------------------
def numpy_piece(n_reads, genome_len, seed, dup_every=0, dup_copies=3):
    """(FASTA text bytes, codes uint8[n, 150]) of 150-bp error-free reads; every dup_every-th read is repeated
    dup_copies times in a row (planted multiplicities)."""
    rng = np.random.default_rng(seed)
    genome = rng.integers(0, 4, genome_len, dtype=np.uint8)
    starts = rng.integers(0, genome_len - 150, n_reads)
    if dup_every:
        src = np.arange(n_reads)
        for j in range(1, dup_copies):
            src[j::dup_every] = src[0::dup_every][: len(src[j::dup_every])]
        starts = starts[src]
    codes = genome[starts[:, None] + np.arange(150)[None, :]]
    rec = np.empty((n_reads, 164), dtype=np.uint8)
    rec[:, 0] = ord(">")
    rec[:, 1] = ord("r")
    ids = np.arange(n_reads, dtype=np.int64)
    for j in range(10):
        rec[:, 2 + j] = (ids // 10 ** (9 - j)) % 10 + 48
    rec[:, 12] = 10
    rec[:, 13:163] = np.frombuffer(b"ACGT", dtype=np.uint8)[codes]
    rec[:, 163] = 10
    return rec.reshape(-1), codes


def numpy_table(codes, k, c):
    """sorted (keys, counts) of all k-mers with count >= c -- keys are the engine's order-preserving 2-bit codes"""
    n, L = codes.shape
    w = L - k + 1
    keys = np.zeros((n, w), dtype=np.uint64)
    for j in range(k):
        keys <<= np.uint64(2)
        keys |= codes[:, j:j + w]
    uniq, cnt = np.unique(keys.reshape(-1), return_counts=True)
    keep = cnt >= c
    return uniq[keep], cnt[keep].astype(np.uint64)


@pytest.mark.parametrize("genome_len,dup_every,c", [(200_000_000, 5, 2), (2_000_000, 0, 2), (200_000_000, 5, 1), (20_000_000, 7, 3)])
def test_full_piece_known_answer(engine, genome_len, dup_every, c):
    """one 100 MiB piece (640 k reads, 76.8 M windows) through the default path: packed lane -> two-level range partition
    -> shared-memory tables; the table (every key, every count) must equal numpy's sort/unique of the same windows.
    Sparse data with planted repeats (bitmap pre-filter + exact second pass), 48x coverage (duplicate-rich mode),
    min_count 1 (every distinct key is a row)."""
    import torch
    reset(engine)
    n_reads = 640_000
    text, codes = numpy_piece(n_reads, genome_len, seed=genome_len % 1000 + dup_every, dup_every=dup_every)
    want_k, want_c = numpy_table(codes, 31, c)
    assert len(want_k) > 1000
    dev = torch.from_numpy(text).cuda()
    for mode in (-1, 0, 1) if c >= 2 else (-1,):
        engine.set_option("count_mode", mode)
        try:
            ovf0 = engine.stat("overflow_buckets")
            table = engine.count_text(dev, 31, c)
            info = table.info()
            got_k, got_c = table.packed_arrays()
            table.close()
        finally:
            reset(engine)
        assert info["wide_rows"] == 0 and len(got_k) == len(want_k), (mode, len(got_k), len(want_k))
        assert np.array_equal(got_k, want_k), f"mode {mode}: keys differ at row {int(np.argmax(got_k != want_k))}"
        assert np.array_equal(got_c, want_c), f"mode {mode}: {int((got_c != want_c).sum())} counts differ"
        if mode == -1 and genome_len >= 20_000_000:
            assert engine.stat("overflow_buckets") == ovf0          # the common case never needs the sort fallback


def test_duplicate_rich_sample_grows_buckets(engine):
    """36x coverage in 8 pieces: from the third piece on the sub-buckets are sized by the measured keys per distinct key
    (the default size would need more than 192 level-1 bins here), persistent kernels launch 4 waves of CTAs (the setting
    of the multi-GPU exchange); every piece is filtered on its own and the sums must equal numpy's"""
    import torch
    reset(engine)
    text, codes = numpy_piece(120_000, 400_000, seed=31)
    dev = torch.from_numpy(text).cuda()
    results = {}
    for growth in (1, 0):
        engine.set_option("hash_bucket_keys", 40)
        engine.set_option("grid_waves", 4)
        engine.set_option("bucket_growth", growth)
        try:
            table, offs = engine.count_sample(dev, 31, 2, len(text) // 8)
            results[growth] = table.packed_arrays()
            table.close()
        finally:
            reset(engine)
    bounds = [o // 164 for o in offs] + [120_000]
    parts = [numpy_table(codes[a:b], 31, 2) for a, b in zip(bounds[:-1], bounds[1:])]
    uk, inv = np.unique(np.concatenate([p[0] for p in parts]), return_inverse=True)
    uc = np.zeros(len(uk), dtype=np.uint64)
    np.add.at(uc, inv, np.concatenate([p[1] for p in parts]))
    assert len(offs) >= 8 and len(uk) > 100_000
    for growth, (got_k, got_c) in results.items():
        assert np.array_equal(got_k, uk) and np.array_equal(got_c, uc), growth


def test_full_piece_determinism(engine):
    """the same 100 MiB piece counted 20 times at -c 2 (both counting modes): identical digests every time -- the
    stress harness of tools/stress_hash.sh as a test (a lost count under any interleaving changes the digest)"""
    import torch
    reset(engine)
    text, codes = numpy_piece(640_000, 300_000_000, seed=99, dup_every=3, dup_copies=2)
    want_k, want_c = numpy_table(codes, 31, 2)
    want = hashlib.md5(want_k.tobytes() + want_c.tobytes()).hexdigest()
    dev = torch.from_numpy(text).cuda()
    for mode in (1, 0):
        engine.set_option("count_mode", mode)
        try:
            for rep in range(20 if mode == 1 else 5):
                table = engine.count_text(dev, 31, 2)
                got_k, got_c = table.packed_arrays()
                table.close()
                assert hashlib.md5(got_k.tobytes() + got_c.tobytes()).hexdigest() == want, f"mode {mode} repetition {rep}"
        finally:
            reset(engine)


def test_oversized_buckets_take_rounds(engine):
    """sub-buckets far beyond one register round (4096 keys): the counting kernel walks them in rounds from global
    memory in both passes; repeated 10 times -- the variant of round 1 that did this lost counts in a timing-dependent
    way, so this is its regression test"""
    import torch
    reset(engine)
    text, codes = numpy_piece(60_000, 500_000, seed=5)           # 7.2 M windows, ~14x coverage: few distinct keys per bucket
    dev = torch.from_numpy(text).cuda()
    for c in (1, 2, 4):
        want_k, want_c = numpy_table(codes, 31, c)
        for bucket_keys, mode in ((20000, 0), (20000, 1), (60000, 0), (9000, 1)):
            engine.set_option("hash_bucket_keys", bucket_keys)
            engine.set_option("count_mode", mode)
            try:
                for rep in range(10):
                    table = engine.count_text(dev, 31, c)
                    got_k, got_c = table.packed_arrays()
                    table.close()
                    assert np.array_equal(got_k, want_k) and np.array_equal(got_c, want_c), (c, bucket_keys, mode, rep)
            finally:
                reset(engine)


def test_add_rows_sums_with_counted_text(engine):
    """mc2_sample_add_rows on a sample that also counted text itself (the dict merge of bin/mercat2.py:121-127 for a
    table counted on another rank): a k-mer present on both sides must come out as ONE summed row -- on the sparse,
    the dense (nucleotide and protein) and the literal-byte paths, rows added before or after the text"""
    reset(engine)
    text = synth_reads(3000, 150, seed=77, n_rate=0.003, lower_rate=0.02, genome_len=40000)
    reads = [b">" + r for r in text.split(b">") if r]
    halves = [b"".join(reads[0::2]), b"".join(reads[1::2])]
    prot = (b">p1\nMKVLAAGIVGLLLAQWERTYIPASDFGHKLCVNMX*\n>p2\nMKVLAAGIVBZUOMKVLAAG*\n" * 40, b">q\nMKVLAAGIVGLLWERTYIPASMKVLAAGXB*\n" * 30)
    for data, k, c in ((halves, 21, 2), (halves, 4, 5), (halves, 33, 2), (halves, 12, 1), (prot, 3, 2), (prot, 6, 1)):
        want = orc.merge_counts(orc.find_kmers_text(h.decode(), k, c) for h in data)
        other = engine.count_text(data[1], k, c)
        ok, oc = other.arrays()
        other.close()
        for rows_first in (False, True):
            sample = engine.sample(k, c)
            if rows_first:
                sample.add_rows(ok, oc)
            sample.add_text(data[0])
            if not rows_first:
                sample.add_rows(ok, oc)
            got = sample.finish().to_dict()
            assert got == want, f"k={k} c={c} rows_first={rows_first}: {diff_msg(got, want)}"


# ---- BASELINE config 5: many protein samples, one table each -------------------------------------------------------
def test_count_batch_equals_per_sample(engine):
    """Engine.count_batch (all samples in ONE pass, the sample index rides above the k-mer code) against the oracle on
    every sample: S5 proteomes, samples without a trailing newline / starting without a header / empty / headers only,
    nucleotide samples, and a batch that must fall back (lower-case residues -> literal-byte rows)"""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
    from tools import synth_s5
    reset(engine)
    prot = [synth_s5.sample_text(j, 120) for j in range(9)]
    odd = [b"MKVLAAGIVGLLLAQWERTY", b"", b">only a header\n", b">p\nMKV*LAAG\nIVGLL*\n>q\nMKVLAAGIVG", prot[0][:-1], b"\n\n>x\n" + prot[1]]
    for texts, k, c in ((prot, 5, 2), (prot, 5, 10), (prot + odd, 3, 2), (odd + prot[:2], 5, 1), (prot[:1], 5, 2), (prot, 12, 1)):
        tables = engine.count_batch(texts, k, c)
        assert len(tables) == len(texts)
        for j, (text, table) in enumerate(zip(texts, tables)):
            want = orc.find_kmers_text(text.decode(), k, c)
            got = table.to_dict()
            assert got == want, f"sample {j} k={k} c={c}: {diff_msg(got, want)}"
            if want:
                assert table.tsv_bytes("s") == orc.tsv_bytes("s", want)
    reads = [synth_reads(400, 150, seed=300 + j, n_rate=0.0, lower_rate=0.0, genome_len=9000) for j in range(5)]
    for texts, k, c in ((reads, 21, 2), (reads, 15, 1)):
        for j, table in enumerate(engine.count_batch(texts, k, c)):
            want = orc.find_kmers_text(texts[j].decode(), k, c)
            assert table.to_dict() == want, f"reads sample {j} k={k}: {diff_msg(table.to_dict(), want)}"
    mixed = [prot[0], prot[1].lower(), b">n\nACGTNNACGTacgt\n", prot[2]]                       # not batchable: per-sample fallback
    for j, table in enumerate(engine.count_batch(mixed, 4, 2)):
        want = orc.find_kmers_text(mixed[j].decode(), 4, 2)
        assert table.to_dict() == want, f"mixed sample {j}: {diff_msg(table.to_dict(), want)}"
    import torch
    dev = [torch.from_numpy(np.frombuffer(t, dtype=np.uint8).copy()).cuda() for t in prot[:4]]      # device-resident texts
    for j, table in enumerate(engine.count_batch(dev, 5, 2)):
        assert table.to_dict() == orc.find_kmers_text(prot[j].decode(), 5, 2)


def test_cfg5_sample_set_pipeline(engine, tmp_path):
    """64 synthetic proteomes (S5 at a size the oracle finishes in seconds), k=5 -c 10 and -c 2, through
    pipeline.run_samples (batched passes): every per-sample TSV byte-identical to the oracle's, no file for a sample
    without survivors; LPT sharding covers every sample exactly once"""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
    from tools import synth_s5
    from mercat2_b200 import distributed as mcd, pipeline
    reset(engine)
    files = {}
    for j in range(64):
        path = tmp_path / f"s{j:02d}.faa"
        path.write_bytes(synth_s5.sample_text(j, 150 if j % 7 else 600))
        files[f"s{j:02d}"] = str(path)
    for c in (10, 2):
        out = tmp_path / f"tsv_{c}"
        out.mkdir()
        res = pipeline.run_samples(files, str(out), 5, c, chunk_size_mb=100, engine=engine, quiet=True)
        assert list(res) == list(files)
        for base, path in files.items():
            want = orc.find_kmers(__import__("pathlib").Path(path), 5, c)
            if want:
                assert res[base] and open(res[base], "rb").read() == orc.tsv_bytes(base, want), base
            else:
                assert res[base] is None and not (out / f"{base}_counts.tsv").exists(), base
    sizes = [os.stat(f).st_size for f in files.values()]
    shards = mcd.shard_lpt(sizes, 8)
    assert sorted(i for sh in shards for i in sh) == list(range(64))
    loads = [sum(sizes[i] for i in sh) for sh in shards]
    assert max(loads) <= 1.25 * (sum(sizes) / 8)


def test_diversity_inputs_and_top_kmers(engine):
    """row N4: the count vector and its device-side abundance spectrum against numpy on the oracle's table; the top-5
    k-mers by mean count against the reference's streaming selection restated on the merged table"""
    from mercat2_b200 import mercat2_diversity
    reset(engine)
    texts = {f"s{j}": synth_reads(2500, 150, seed=400 + j, n_rate=0.002, lower_rate=0.01, genome_len=30000 + 7000 * j) for j in range(4)}
    tables = {name: engine.count_text(text, 9, 2) for name, text in texts.items()}
    for name, text in texts.items():
        want = orc.find_kmers_text(text.decode(), 9, 2)
        vec = np.array([want[key] for key in sorted(want)], dtype=np.uint64)
        inputs = mercat2_diversity.alpha_inputs(tables[name])
        assert np.array_equal(inputs["counts"], vec)
        sp = inputs["spectrum"]
        assert sp["observed"] == len(vec) and sp["total"] == int(vec.sum()) and sp["max"] == int(vec.max())
        assert sp["sum_squares"] == sum(int(x) ** 2 for x in vec.tolist())
        assert sp["seen_exactly"] == {i: int((vec == i).sum()) for i in range(1, 11)}
    names, top = mercat2_diversity.top_kmers(tables, 5, engine)
    dicts = {name: orc.find_kmers_text(texts[name].decode(), 9, 2) for name in names}
    union = sorted(set().union(*[set(d) for d in dicts.values()]))
    keep = []                                              # lib/mercat2_figures.py:50-65, on the sorted-union table
    for kmer in union:
        row = [dicts[n].get(kmer, 0) for n in names]
        if len(keep) < 5:
            keep.append((kmer, row))
        else:
            keep.sort(key=lambda kr: sum(kr[1]) / len(kr[1]))
            if sum(row) / len(row) > sum(keep[0][1]) / len(keep[0][1]):
                keep[0] = (kmer, row)
    assert sorted(top) == sorted(keep)
    assert [sum(r) for _, r in top] == sorted((sum(r) for _, r in top), reverse=True)
    for t in tables.values():
        t.close()


def test_big_chunk_async_groups_and_row_streaming(engine):
    """a very large chunk with min_count >= 2: the level-0 groups are enqueued without a host round trip, their rows
    collect in the consumed front of the key array, overflowed sub-buckets are copied aside by the device; and
    count_text_rows streams the rows of finished groups to (pinned) host memory.  All variants must give numpy's table."""
    import torch
    reset(engine)
    text, codes = numpy_piece(60_000, 3_000_000, seed=17, dup_every=4, dup_copies=2)        # 7.2 M windows
    dev = torch.from_numpy(text).cuda()
    for c in (2, 3):
        want_k, want_c = numpy_table(codes, 31, c)
        for opts in ({"batch_symbols": 1 << 20}, {"batch_symbols": 1 << 20, "group_sync": 1}, {"batch_symbols": 1 << 20, "count_mode": 0},
                     {"batch_symbols": 1 << 20, "hash_bucket_keys": 1_000_000, "count_mode": 0},       # tables overflow -> device-side collection -> sort path
                     {"batch_symbols": 1 << 19, "hash_bucket_keys": 1_000_000, "count_mode": 1}):
            for name, value in opts.items():
                engine.set_option(name, value)
            try:
                table = engine.count_text(dev, 31, c)
                got_k, got_c = table.packed_arrays()
                table.close()
                rows = torch.empty((len(want_k) + 8, 2), dtype=torch.int64, pin_memory=True)
                n = engine.count_text_rows(dev, 31, c, rows.data_ptr(), rows.shape[0])
                keys64 = torch.empty(len(want_k) + 8, dtype=torch.int64, pin_memory=True)
                counts32 = torch.empty(len(want_k) + 8, dtype=torch.int32, pin_memory=True)
                n2 = engine.count_text_rows_split(dev, 31, c, keys64.data_ptr(), counts32.data_ptr(), len(keys64))
            finally:
                engine.set_option("group_sync", 0)
                reset(engine)
            assert np.array_equal(got_k, want_k) and np.array_equal(got_c, want_c), (c, opts)
            got = rows[:n].numpy().view(np.uint64)
            assert n == len(want_k) and np.array_equal(got[:, 0], want_k) and np.array_equal(got[:, 1], want_c), (c, opts, n)
            assert n2 == len(want_k) and np.array_equal(keys64[:n2].numpy().view(np.uint64), want_k), (c, opts, n2)       # 12-byte rows
            assert np.array_equal(counts32[:n2].numpy().view(np.uint32).astype(np.uint64), want_c), (c, opts)
    # a text with k-mers outside ACGT cannot be delivered as packed rows; small texts take the ordinary path
    import mercat2_b200
    rows = torch.empty((1024, 2), dtype=torch.int64, pin_memory=True)
    with pytest.raises(mercat2_b200.Mc2Error):
        engine.count_text_rows(b">r\nACGTACGTACGTACGTACGTACGTACGTaACGTACGTACGTACGT\n", 3, 1, rows.data_ptr(), 1024)      # 'a': a literal-byte row
    n = engine.count_text_rows(b">r\nACGTACGTACGT\n", 3, 2, rows.data_ptr(), 1024)
    want = orc.find_kmers_text(">r\nACGTACGTACGT\n", 3, 2)
    assert n == len(want)
    with pytest.raises(mercat2_b200.Mc2Error):
        engine.count_text_rows(dev, 31, 2, rows.data_ptr(), 16)                                  # buffer too small
    k64 = torch.empty(1024, dtype=torch.int64, pin_memory=True)
    c32 = torch.empty(1024, dtype=torch.int32, pin_memory=True)
    n = engine.count_text_rows_split(b">r\nACGTACGTACGT\n", 3, 2, k64.data_ptr(), c32.data_ptr(), 1024)                # the ordinary (small) path
    assert n == len(want) and [c32[i].item() for i in range(n)] == [want[kmer] for kmer in sorted(want)]          # rows are in k-mer order
    assert k64[:n].tolist() == sorted(k64[:n].tolist())
    with pytest.raises(mercat2_b200.Mc2Error):
        engine.count_text_rows_split(dev, 31, 2, k64.data_ptr(), c32.data_ptr(), 16)                                  # buffer too small


def test_fastq_to_fasta_on_device(engine, golden_configs, tmp_path):
    """row N2: fq2fa on the device against the reference's sed pipeline (golden digest of Test_R1) and against the host
    restatement on awkward inputs; the converted text is counted straight from HBM"""
    from mercat2_b200 import mercat2_fasta
    reset(engine)
    out = mercat2_fasta.fq2fa(str(GOLDEN / "data/Test_R1.fastq.gz"), str(tmp_path / "clean"), "Test_R1", engine)
    assert md5(gzip.open(out, "rb").read()) == golden_configs["test_r1_k12"]["fasta_md5"]
    text = mercat2_fasta.fq2fa_device(str(GOLDEN / "data/Test_R1.fastq.gz"), engine)
    for c in (1, 2):
        want = orc.find_kmers_text(gzip.open(out, "rb").read().decode(), 12, c)
        assert engine.count_text(text, 12, c).to_dict() == want
    text.close()
    cases = [b"", b"@r1\nACGT\n+\nIIII\n", b"@r1\nACGT\n+\nIIII", b"@r1\r\nACGT\r\n+\r\nIIII\r\n@r2\nAC", b"r1 no at\nACGT\n+\nIIII\n@r2\nGG\n+\nII\n",
              b"@r1\n\n+\n\n@r2\nTTTT\n+\nIIII\n", b"\n\n\n\n@x\nA\n", b"@only"]
    rng = random.Random(3)
    big = b"".join(b"@read%d extra\n%s\n+\n%s\n" % (i, bytes(rng.choice(b"ACGTN") for _ in range(rng.randrange(1, 300))), b"I" * 5) for i in range(5000))
    for j, data in enumerate(cases + [big, big[:-1]]):
        src = tmp_path / f"case{j}.fastq"
        src.write_bytes(data)
        ref = mercat2_fasta.fq2fa_host(str(src), str(tmp_path / "ref"), f"case{j}")
        dev = engine.fastq_to_fasta(data)
        assert dev.to_bytes() == gzip.open(ref, "rb").read(), f"case {j}"
        dev.close()


def test_row_merge_through_range_partition(engine):
    """the dict merge of many filtered tables (bin/mercat2.py:121-127) on the opt-in range path (row_merge=1): a sample of
    many pieces with survivors (parts with overlapping key ranges) and a table rebuilt from unsorted rows with repeated keys
    must equal numpy's reduce-by-key, and so must the default sort-based merge"""
    import torch
    reset(engine)
    engine.set_option("row_merge", 1)
    text, codes = numpy_piece(120_000, 4_000_000, seed=23)                 # 14.4 M windows, ~3.6x coverage
    dev = torch.from_numpy(text).cuda()
    pieces = 12
    merges0 = engine.stat("row_merges")
    table, offs = engine.count_sample(dev, 31, 2, len(text) // pieces)
    got_k, got_c = table.packed_arrays()
    table.close()
    assert engine.stat("row_merges") > merges0                         # (not the sort fallback)
    engine.set_option("hash_bucket_keys", 40)                          # rows beyond one pass: parts cut at common splitters, ranges summed one by one
    engine.set_option("row_merge", 1)
    try:
        merges0 = engine.stat("row_merges")
        table, _ = engine.count_sample(dev, 31, 2, len(text) // pieces)
        cut_k, cut_c = table.packed_arrays()
        table.close()
        assert engine.stat("row_merges") >= merges0 + 2
    finally:
        reset(engine)
    merges0 = engine.stat("row_merges")
    table, _ = engine.count_sample(dev, 31, 2, len(text) // pieces)    # the default: sort + segmented sum
    srt_k, srt_c = table.packed_arrays()
    table.close()
    assert engine.stat("row_merges") == merges0
    bounds = [o // 164 for o in offs] + [120_000]
    parts = [numpy_table(codes[a:b], 31, 2) for a, b in zip(bounds[:-1], bounds[1:])]
    all_k = np.concatenate([p[0] for p in parts])
    all_c = np.concatenate([p[1] for p in parts])
    uk, inv = np.unique(all_k, return_inverse=True)
    uc = np.zeros(len(uk), dtype=np.uint64)
    np.add.at(uc, inv, all_c)
    assert len(offs) >= pieces and len(all_k) > 500_000
    assert np.array_equal(got_k, uk) and np.array_equal(got_c, uc)
    assert np.array_equal(cut_k, uk) and np.array_equal(cut_c, uc)
    assert np.array_equal(srt_k, uk) and np.array_equal(srt_c, uc)
    # unsorted rows with repeats and huge counts (64-bit sums), straight into table_from_rows
    rng = np.random.default_rng(4)
    keys = rng.integers(0, 1 << 62, 300_000, dtype=np.int64).astype(np.uint64)
    keys = np.concatenate([keys, keys[:100_000], keys[:50_000], np.array([(1 << 62) - 1] * 3, dtype=np.uint64)])
    cnts = rng.integers(1, 1 << 40, len(keys), dtype=np.int64).astype(np.uint64)
    perm = rng.permutation(len(keys))
    keys, cnts = keys[perm], cnts[perm]
    wk, winv = np.unique(keys, return_inverse=True)
    wc = np.zeros(len(wk), dtype=np.uint64)
    np.add.at(wc, winv, cnts)
    dk, dc = torch.from_numpy(keys.view(np.int64)).cuda(), torch.from_numpy(cnts.view(np.int64)).cuda()
    torch.cuda.synchronize()
    for algo in (0, 1):
        engine.set_option("row_merge", 1 - algo)
        try:
            t = engine.table_from_rows(31, 0, 0, dk.data_ptr(), dc.data_ptr(), len(keys), True)
            tk, tc = t.packed_arrays()
            t.close()
        finally:
            reset(engine)
        assert np.array_equal(tk, wk) and np.array_equal(tc, wc), algo
