#!/usr/bin/env python
"""One compact pass over every path the headline bench does not touch, for ncu (launch list + full captures of the
dense / general-lane / literal-byte / chunker / TSV / metrics / merge / text-prep kernels):

    python tools/profile_paths.py            # ~2 s of GPU work after a warm-up pass
"""
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

import bench  # noqa: E402
import mercat2_b200  # noqa: E402
from tools import bench_workloads as wl  # noqa: E402
from tools import synth_s5  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    engine = mercat2_b200.Engine(0)
    genome, _ = wl.genome_text(dev, 40_000_000, 1)
    genome_n, _ = wl.genome_text(dev, 8_000_000, 2, with_n=True)
    genomes = bench.make_genomes(dev, 0.05)
    reads = bench.make_reads_text(dev, genomes, 1_000_000, 0)
    prot, _ = wl.protein_text(dev, 10_000, 1000)
    crlf = torch.frombuffer(bytearray(bytes(genome[:20_000_000].cpu().numpy().tobytes()).replace(b"\n", b"\r\n")), dtype=torch.uint8).to(dev)
    fastq = b"".join(b"@r%d\n%s\n+\n%s\n" % (i, b"ACGT" * 30, b"I" * 120) for i in range(200_000))
    s5 = [torch.frombuffer(bytearray(synth_s5.sample_text(j, 1000)), dtype=torch.uint8).to(dev) for j in range(16)]
    tmp = ("/dev/shm" if os.path.isdir("/dev/shm") else "/tmp") + "/mc2_profile.tsv"
    for rep in range(int(os.environ.get("MC2_PROFILE_PASSES", "2"))):   # pass 0 warms up (module load, pool growth); ncu full captures use 1
        for text, k, c, s in ((genome, 3, 10, 0), (genome, 12, 10, 0), (reads, 12, 10, 100 << 20), (genome_n, 40, 1, 0), (genome_n, 31, 1, 0),
                              (prot, 5, 10, 0), (prot, 3, 10, 0), (prot, 8, 2, 0), (reads, 31, 2, 50 << 20)):
            table, _ = engine.count_sample(text, k, c, s)
            if table.rows and k != 40:
                table.write_tsv(tmp, "s")
                os.unlink(tmp)
            table.close()
        engine.set_option("force_path", 1)                 # protein k=5 through the dense 26^5 table (a small sample would go sparse)
        engine.count_text(prot, 5, 10).close()
        engine.set_option("force_path", 0)
        engine.set_option("sparse_algo", 1)                # the radix-sort fallback (rs_*, rle_*)
        engine.count_text(reads[:40_000_000], 31, 2).close()
        engine.set_option("sparse_algo", 0)
        engine.chunk_offsets(genome, 4 << 20)
        engine.chunk_offsets(crlf, 4 << 20)
        engine.protein_metrics(prot)
        for t in engine.count_batch(s5, 5, 10):
            t.close()
        tabs = [engine.count_text(x, 5, 2) for x in s5[:4]]
        m = engine.merge_tables(tabs)
        m.top_rows(5)
        m.write_tsv(tmp, "k-mer", ["a", "b", "c", "d"])
        os.unlink(tmp)
        m.close()
        tabs[0].count_spectrum()
        for t in tabs:
            t.close()
        engine.fastq_to_fasta(fastq).close()
        torch.cuda.synchronize()
    print("profile pass done; launches:", engine.stat("launches"))


if __name__ == "__main__":
    main()
