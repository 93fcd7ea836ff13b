"""CPU-only checks of the host side: the C-ABI library loads and exports every symbol the header
declares, the host helpers (human2bytes, piece names, removeN, fastq->fasta) match golden vectors
made by the reference, and the product refuses to run without a GPU (no CPU fallback)."""
import gzip
import hashlib
import re
import sys

import pytest

from conftest import GOLDEN, ROOT


def md5(b):
    return hashlib.md5(b).hexdigest()


def test_library_exports_every_declared_symbol():
    from mercat2_b200 import _native
    header = (ROOT / "include" / "mercat2_b200.h").read_text()
    declared = set(re.findall(r"\b(mc2_[a-z0-9_]+)\s*\(", header))
    bound = {name for name, _, _ in _native.SYMBOLS}
    assert declared == bound, (declared - bound, bound - declared)
    lib = _native.load_library()
    for name in declared:
        assert hasattr(lib, name), name
    assert b"sm_100a" in lib.mc2_version()


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from mercat2_b200 import _native, mercat2_kmers
    with pytest.raises(_native.Mc2Error):
        _native.Engine(0)
    with pytest.raises(_native.Mc2Error):
        mercat2_kmers.calculateKmerCount("ACGTACGT", 3)


def test_product_does_not_import_oracle():
    for path in (ROOT / "mercat2_b200").rglob("*.py"):
        text = path.read_text()
        assert "import oracle" not in text and "from oracle" not in text, path


def test_human2bytes_and_piece_names():
    from mercat2_b200.mercat2_Chunker import human2bytes, piece_name
    assert human2bytes("0 B") == 0 and human2bytes("1 K") == 1024 and human2bytes("1M") == 1 << 20
    assert human2bytes("1 Gi") == 1 << 30 and human2bytes("1 tera") == 1 << 40
    assert human2bytes("0.5kilo") == 512 and human2bytes("0.1  byte") == 0 and human2bytes("1 k") == 1024
    with pytest.raises(ValueError):
        human2bytes("12 foo")
    assert piece_name("x/c.fna", 0) == "c.00000"
    assert piece_name("x/c_clean.fna.gz", 2) == "c_clean.00002.fna"
    assert piece_name("DJ_pro.faa", 1) == "DJ_pro.00001"


@pytest.mark.parametrize("base", ["RW1", "GIC31"])
def test_removeN_matches_reference(base, golden_configs, tmp_path):
    from mercat2_b200.mercat2_fasta import removeN
    out, stats = removeN(GOLDEN / "data/fna_gz" / f"{base}.fna.gz", tmp_path / "clean", False)
    assert out.name == f"{base}_clean.fna.gz"
    blob = gzip.open(out, "rb").read()
    want = golden_configs["removeN"][base]
    assert len(blob) == want["clean_bytes"] and md5(blob) == want["clean_md5"]
    assert 0.0 < stats["GC Content"] < 100.0


@pytest.mark.parametrize("toupper", [False, True])
def test_removeN_scaffolds_with_N_runs(toupper, golden_configs, tmp_path):
    from mercat2_b200.mercat2_fasta import removeN
    src = tmp_path / "Scaffolds_with-NNN.fna"
    src.write_bytes(gzip.open(GOLDEN / "data/Scaffolds_with-NNN.fna.gz", "rb").read())
    out, _ = removeN(src, tmp_path / "clean", toupper)
    blob = gzip.open(out, "rb").read()
    want = golden_configs["removeN"]["Scaffolds_toupper" if toupper else "Scaffolds"]
    assert len(blob) == want["clean_bytes"] and md5(blob) == want["clean_md5"]


def test_fq2fa_matches_reference(golden_configs, tmp_path):
    from mercat2_b200.mercat2_fasta import fq2fa_host as fq2fa          # (the host restatement: the CPU suite has no device)
    out = fq2fa(str(GOLDEN / "data/Test_R1.fastq.gz"), str(tmp_path / "clean"), "Test_R1")
    assert out.endswith("Test_R1.fna.gz")
    assert md5(gzip.open(out, "rb").read()) == golden_configs["test_r1_k12"]["fasta_md5"]


def test_chunk_trigger(tmp_path):
    from mercat2_b200.pipeline import chunk_trigger
    f = tmp_path / "x.fa"
    f.write_bytes(b"A" * (1 << 20))
    assert chunk_trigger(f, 1) == 1 << 20 and chunk_trigger(f, 2) == 0 and chunk_trigger(f, 0) == 0


def test_cli_flags_and_classification():
    from mercat2_b200 import cli
    args, _ = cli.parseargs(["-i", __file__, "-k", "5"])
    assert args.c == 10 and args.s == 100 and args.o == "mercat_results" and not args.toupper and not args.skipclean
    args, _ = cli.parseargs(["-i", __file__, "-k", "31", "-c", "2", "-s", "1", "-toupper", "-skipclean", "-replace"])
    assert (args.k, args.c, args.s, args.toupper, args.skipclean, args.replace) == (31, 2, 1, True, True, True)
    assert cli.classify("x/DJ.fna.gz") == ("nucleotide", "DJ")
    assert cli.classify("x/DJ_pro.faa") == ("protein", "DJ_pro")
    assert cli.classify("x/Test_R1.fastq") == ("fastq", "Test_R1")
    assert cli.classify("x/a.b.fasta") == ("nucleotide", "a.b")
    assert cli.classify("x/readme.txt")[0] is None


def test_distributed_host_helpers():
    """host-side pieces of the multi-GPU paths: header-aligned position split, packed-key text, literal-row sums"""
    import numpy as np
    from mercat2_b200 import distributed as mcd
    from oracle import mercat2_oracle as orc
    text = b">a x\nACGT\nAC\n>b\nGG\n>c y z\nTTTT\nAA\n>d\nC\n"
    for world in (1, 2, 3, 5, 9):
        parts = mcd.split_at_headers(text, world)
        assert len(parts) == world and b"".join(parts) == text
        assert all((not p) or p[:1] == b">" for p in parts)
        # splitting at header lines loses no window: per-part counts (unfiltered) sum to the whole
        whole = orc.find_kmers_text(text.decode(), 2, 1)
        summed = orc.merge_counts(orc.find_kmers_text(p.decode(), 2, 1) for p in parts if p)
        assert summed == whole
    # packed key -> text, all layouts (csrc/tsv.cuh::tsv_decode)
    assert mcd.decode_key(0b00011011, 4, mcd.ENC_NT2, mcd.KEY_CODE) == b"ACGT"
    assert mcd.decode_key((12 << 10) | (0 << 5) | 25, 3, mcd.ENC_AA5, mcd.KEY_CODE) == b"MAZ"
    assert mcd.decode_key((ord("a") << 8) | ord("N"), 2, mcd.ENC_BYTE, mcd.KEY_CODE) == b"aN"
    assert mcd.decode_key(2 * 26 * 26 + 0 * 26 + 25, 3, mcd.ENC_AA5, mcd.KEY_DENSE_AA) == b"CAZ"
    # literal-byte rows summed by text
    k = 3
    a = (np.frombuffer(b"ANTNNNANT", np.uint8).reshape(-1, k), np.array([2, 5, 1], np.uint64))
    b = (np.frombuffer(b"NNNacg", np.uint8).reshape(-1, k), np.array([7, 4], np.uint64))
    sk, sc = mcd._sum_rows([a[0], b[0]], [a[1], b[1]], k)
    got = {bytes(r): int(c) for r, c in zip(sk, sc)}
    assert got == {b"ANT": 3, b"NNN": 12, b"acg": 4}
    assert [bytes(r) for r in sk] == sorted(got)


def test_bind_to_gpu_numa_is_safe_without_a_gpu():
    """the NUMA helper must never raise (no NVML / no device -> False) and must not change affinity then"""
    import os
    import mercat2_b200
    before = os.sched_getaffinity(0)
    ok = mercat2_b200.bind_to_gpu_numa(0)
    assert ok in (True, False)
    if not ok:
        assert os.sched_getaffinity(0) == before
    else:
        os.sched_setaffinity(0, before)


def test_public_header_is_plain_c():
    """the drop-in boundary must be bindable from C (cgo / ctypes / JNI stubs): the header compiles as C99 and as C++"""
    import shutil
    import subprocess
    from pathlib import Path
    header = Path(__file__).resolve().parents[1] / "include" / "mercat2_b200.h"
    gcc = shutil.which("gcc")
    if not gcc:
        pytest.skip("no gcc")
    subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-x", "c", str(header)], check=True)
    gxx = shutil.which("g++")
    if gxx:
        subprocess.run([gxx, "-std=c++11", "-fsyntax-only", "-x", "c++", str(header)], check=True)


def test_c_example_links_against_the_abi(tmp_path):
    """examples/count_file.c (a plain C99 client) compiles and links against the shared library; without a GPU it
    must fail loudly at mc2_engine_create (no CPU fallback)"""
    import shutil
    import subprocess
    from pathlib import Path
    import torch
    root = Path(__file__).resolve().parents[1]
    gcc = shutil.which("gcc")
    lib = root / "mercat2_b200" / "libmercat2_b200.so"
    if not gcc or not lib.exists():
        pytest.skip("needs gcc and the built library")
    exe = tmp_path / "count_file"
    subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", f"-I{root / 'include'}", str(root / "examples" / "count_file.c"),
                    f"-L{lib.parent}", "-lmercat2_b200", f"-Wl,-rpath,{lib.parent}", "-o", str(exe)], check=True)
    if not torch.cuda.is_available():
        run = subprocess.run([str(exe), str(root / "tests/golden/data/fna_gz/DJ.fna.gz"), "3", "10", "0", str(tmp_path / "x.tsv")],
                             capture_output=True, text=True)
        assert run.returncode != 0 and "no CPU fallback" in run.stderr and not (tmp_path / "x.tsv").exists()


def test_bench_accounting_helpers():
    """the roofline accounting of bench.py: algorithmic bytes per kernel and the phase split"""
    sys.path.insert(0, str(ROOT))
    import bench
    reads = 1_000_000
    text_bytes = reads * bench.REC_BYTES
    windows = reads * (bench.READ_LEN - 31 + 1)
    rows = windows // 10
    assert bench.algorithmic_bytes_per_step("hk_scatter1_kernel", windows, text_bytes, rows) == 16.0 * windows      # key read once, written once
    assert bench.algorithmic_bytes_per_step("rc_count_kernel<0>", windows, text_bytes, rows) == 8.0 * windows + 16.0 * rows
    assert bench.algorithmic_bytes_per_step("fn_parse_kernel<2>", windows, text_bytes, rows) == text_bytes
    packed = bench.algorithmic_bytes_per_step("fn_hist_kernel", windows, text_bytes, rows)
    assert 0.37 * reads * bench.READ_LEN < packed < 0.39 * reads * (bench.READ_LEN + 1)                                # 3 bits per symbol
    assert bench.algorithmic_bytes_per_step("not_a_kernel", windows, text_bytes, rows) is None
    profile = {"fn_parse_kernel<1>": {"us": 2000.0, "launches": 2}, "hk_scatter1_kernel": {"us": 4000.0, "launches": 4},
               "rc_count_kernel<0>": {"us": 6000.0, "launches": 4}, "rc_gather_kernel": {"us": 500.0, "launches": 4},
               "mystery_kernel": {"us": 100.0, "launches": 1}}
    ph = bench.phases_from_profile(profile, 2)
    assert ph == {"parse_ms": 1.0, "partition_ms": 2.0, "count_ms": 3.0, "emit_ms": 0.25, "other_ms": 0.05}
    assert "-s 0" in bench.workload_name(31, 2, 0) and "one global table" in bench.workload_name(31, 2, 0)
    assert "per-chunk filter" in bench.workload_name(31, 10, 100)
