#!/usr/bin/env python
"""Throughput of the paths the headline bench does not touch (S1-like genome k=3, S3' reads k=12,
S5 proteins k=5, c=1 sparse, the wide path, TSV emission, protein metrics).  One line per workload:
input symbols/s with the text resident in HBM, the counting path the engine chose, rows of the result.

    python tools/bench_workloads.py [--reps 3]

Not a bench.py replacement: numbers explain where the non-headline paths stand."""
from __future__ import annotations

import argparse
import json
import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

import bench  # noqa: E402
import mercat2_b200  # noqa: E402

TSV_PATH = ("/dev/shm" if os.path.isdir("/dev/shm") else "/tmp") + "/mc2_bench_counts.tsv"
AA = b"ACDEFGHIKLMNPQRSTVWY"
AA_FREQ = [9.9, 0.9, 5.4, 5.9, 3.7, 8.1, 2.2, 5.1, 3.9, 10.6, 2.4, 3.0, 4.9, 3.8, 6.6, 5.6, 5.2, 7.3, 1.3, 2.6]


def protein_text(device, n_proteins, seed):
    """FASTA of n_proteins proteins, lengths 30 + geometric(mean 300) capped at 5000, 60-column lines, trailing '*'."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    lens = (30 + torch.empty(n_proteins).exponential_(1 / 300.0, generator=g)).clamp(max=5000).long()
    probs = torch.tensor(AA_FREQ)
    total = int(lens.sum())
    res = torch.tensor(list(AA), dtype=torch.uint8)[torch.multinomial(probs, total, replacement=True, generator=g)].numpy().tobytes()
    parts, o = [], 0
    for i, L in enumerate(lens.tolist()):
        seq = res[o:o + L] + b"*"
        o += L
        parts.append(b">p%07d hypothetical protein\n" % i)
        parts.extend(seq[j:j + 60] + b"\n" for j in range(0, len(seq), 60))
    data = b"".join(parts)
    return torch.frombuffer(bytearray(data), dtype=torch.uint8).to(device), total


def genome_text(device, n_bases, seed, with_n=False):
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    codes = torch.randint(0, 4, (n_bases,), dtype=torch.uint8, device=device, generator=g)
    lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=device)
    seq = lut[codes.long()]
    if with_n:                                    # one N per ~2000 bases: windows over it take the literal-byte path
        idx = torch.randint(0, n_bases, (n_bases // 2000,), device=device, generator=g)
        seq[idx] = ord("N")
    cols = 80
    rows = n_bases // cols
    body = torch.empty((rows, cols + 1), dtype=torch.uint8, device=device)
    body[:, :cols] = seq[:rows * cols].view(rows, cols)
    body[:, cols] = 10
    head = torch.tensor(list(b">contig_1 synthetic\n"), dtype=torch.uint8, device=device)
    return torch.cat([head, body.reshape(-1)]), rows * cols


def run(engine, name, text, symbols, k, c, chunk_bytes, reps, tsv=False):
    torch.cuda.synchronize()
    best, rows = None, 0
    for _ in range(reps + 1):
        t0 = time.perf_counter()
        table, offsets = engine.count_sample(text, k, c, chunk_bytes)
        rows = table.rows
        nbytes = None
        if tsv and rows:
            t1 = time.perf_counter()
            table.write_tsv(TSV_PATH, "s")
            tsv_s = time.perf_counter() - t1
            nbytes = os.path.getsize(TSV_PATH)
            os.unlink(TSV_PATH)
        table.close()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    engine.set_option("profile", 2)              # one more pass with per-kernel CUDA-event timing
    table, _ = engine.count_sample(text, k, c, chunk_bytes)
    table.close()
    prof = engine.profile()
    engine.set_option("profile", 0)
    kern = {n.replace("_kernel", ""): round(v["us"] / 1e3, 3) for n, v in sorted(prof.items(), key=lambda kv: -kv[1]["us"])[:7]}
    out = {"workload": name, "k": k, "min_count": c, "symbols": symbols, "text_bytes": int(text.numel()),
           "pieces": len(offsets), "rows": rows, "seconds": round(best, 5), "symbols_per_s": round(symbols / best),
           "kernel_ms": kern}
    if tsv and rows:
        out["tsv_bytes"] = nbytes
        out["tsv_seconds_last"] = round(tsv_s, 5)
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--scale", type=float, default=1.0)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    engine = mercat2_b200.Engine(0)
    sc = args.scale

    text, n = genome_text(dev, int(200e6 * sc), 1)
    run(engine, "genome 200 Mbp, k=3 (dense, shared-memory histogram)", text, n, 3, 10, 0, args.reps)
    run(engine, "genome 200 Mbp, k=12 (dense 4^12 table)", text, n, 12, 10, 0, args.reps)
    run(engine, "genome 200 Mbp, k=31 -c 1 (sparse, every k-mer kept) + TSV", text, n, 31, 1, 0, args.reps, tsv=True)
    del text
    text, n = genome_text(dev, int(100e6 * sc), 2, with_n=True)
    run(engine, "genome 100 Mbp with N every ~2 kbp, k=31 -c 1 (sparse + literal-byte rows) + TSV", text, n, 31, 1, 0, args.reps, tsv=True)
    del text

    genomes = bench.make_genomes(dev, 0.05)
    reads = bench.make_reads_text(dev, genomes, int(10e6 * sc), 0)
    run(engine, "S3': 10 M reads x 150 bp, k=12 -c 10 -s 100", reads, int(10e6 * sc) * 150, 12, 10, 100 << 20, args.reps)
    run(engine, "10 M reads x 150 bp, k=31 -c 2 -s 100 (survivors) + TSV", reads, int(10e6 * sc) * 150, 31, 2, 100 << 20, args.reps, tsv=True)
    del reads, genomes
    for scale, label in ((0.025, "24x"), (0.005, "120x")):          # heavily duplicated keys: isolate-genome coverage per piece
        genomes = bench.make_genomes(dev, scale)
        reads = bench.make_reads_text(dev, genomes, int(4e6 * sc), 0)
        ov0 = engine.stat("overflow_buckets")
        run(engine, f"4 M reads x 150 bp at {label} coverage, k=31 -c 10 -s 100", reads, int(4e6 * sc) * 150, 31, 10, 100 << 20, args.reps, tsv=True)
        print(json.dumps({"overflow_buckets": engine.stat("overflow_buckets") - ov0}), flush=True)
        del reads, genomes

    # a FILE through the engine's own reader (page cache -> pinned buffers -> device), chunked like the CLI would
    genomes = bench.make_genomes(dev, 0.05)
    reads = bench.make_reads_text(dev, genomes, int(6e6 * sc), 0)
    fpath = TSV_PATH + ".reads.fna"
    with open(fpath, "wb") as f:
        f.write(reads.cpu().numpy().tobytes())
    nbytes = os.path.getsize(fpath)
    del reads, genomes
    for _ in range(2):
        t0 = time.perf_counter()
        smp = engine.sample(31, 10)
        pieces = smp.add_file(fpath, 100 << 20)
        tbl = smp.finish()
        dt = time.perf_counter() - t0
        tbl.close()
    print(json.dumps({"workload": "file reader: 6 M reads FASTA file (page cache), k=31 -c 10 -s 100", "file_bytes": nbytes, "pieces": pieces,
                      "seconds": round(dt, 4), "symbols_per_s": round(6e6 * sc * 150 / dt), "file_GB_per_s": round(nbytes / dt / 1e9, 2)}), flush=True)
    os.unlink(fpath)

    # BASELINE config 5: 64 synthetic proteomes (S5), k=5 -c 10, one table each -- all in ONE batched pass vs one pass per sample
    from tools import synth_s5
    for proteins, label in ((5000, "S5"), (50000, "S5 x10")):
        proteins = max(50, int(proteins * sc))
        texts = [torch.frombuffer(bytearray(synth_s5.sample_text(j, proteins)), dtype=torch.uint8).to(dev) for j in range(64)]
        residues = sum(int(t.numel()) for t in texts)              # (text bytes: residues + headers + newlines)
        torch.cuda.synchronize()
        for mode in ("batched", "one pass per sample"):
            best = None
            for _ in range(args.reps):
                t0 = time.perf_counter()
                if mode == "batched":
                    tables = []
                    at = 0
                    while at < len(texts):                           # batches of <= 96 MB of text
                        size, end = 0, at
                        while end < len(texts) and (end == at or size + int(texts[end].numel()) <= (96 << 20)):
                            size += int(texts[end].numel())
                            end += 1
                        tables += engine.count_batch(texts[at:end], 5, 10)
                        at = end
                else:
                    tables = [engine.count_text(t, 5, 10) for t in texts]
                rows = sum(t.rows for t in tables)
                for t in tables:
                    t.close()
                dt = time.perf_counter() - t0
                best = dt if best is None else min(best, dt)
            print(json.dumps({"workload": f"cfg5 {label}: 64 proteomes x {proteins} proteins, k=5 -c 10, {mode}", "text_bytes": residues,
                              "rows": rows, "seconds": round(best, 5), "samples_per_s": round(64 / best, 1), "text_bytes_per_s": round(residues / best)}), flush=True)
        del texts

    text, n = protein_text(dev, int(50000 * sc), 1000)
    run(engine, "S5 x10: 50 k proteins, k=5 -c 10 (dense 26^5)", text, n, 5, 10, 0, args.reps, tsv=True)
    run(engine, "50 k proteins, k=3 -c 10 (dense, shared memory)", text, n, 3, 10, 0, args.reps, tsv=True)
    run(engine, "50 k proteins, k=8 -c 2 (sparse, 5-bit codes)", text, n, 8, 2, 0, args.reps, tsv=True)
    torch.cuda.synchronize()
    for _ in range(args.reps):
        t0 = time.perf_counter()
        m = engine.protein_metrics(text) if hasattr(engine, "protein_metrics") else None
        dt = time.perf_counter() - t0
    if m is not None:
        print(json.dumps({"workload": "protein metrics (pI/MW/hydro) of 50 k proteins", "residues": n, "seconds": round(dt, 5),
                          "residues_per_s": round(n / dt)}), flush=True)


if __name__ == "__main__":
    main()
