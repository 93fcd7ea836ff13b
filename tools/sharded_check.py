#!/usr/bin/env python
"""N-GPU check of the device-resident sample merge (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/sharded_check.py

Pieces of one sample are dealt round-robin to the ranks; the merged TSV (written by byte ranges, one per rank) must
equal the oracle's run over all pieces.  Also times the exchange on a larger synthetic table."""
import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import mercat2_b200  # noqa: E402
from mercat2_b200 import distributed as mcd  # noqa: E402
from oracle import mercat2_oracle as orc  # noqa: E402


def synth(n_reads, seed, genome_len=40000, n_rate=0.003):
    rng = np.random.default_rng(seed)
    genome = rng.integers(0, 4, genome_len)
    out = []
    for i in range(n_reads):
        s = int(rng.integers(0, genome_len - 150))
        seq = bytearray(b"ACGT"[c] for c in genome[s:s + 150])
        for j in np.nonzero(rng.random(150) < n_rate)[0]:
            seq[j] = ord("N")
        out.append(b">r%d desc\n" % i + bytes(seq) + b"\n")
    return out


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    mcd.init_nccl(dev, dist)
    engine = mercat2_b200.Engine(local)
    reads = synth(6000, 5)
    pieces = [b"".join(reads[i::8]) for i in range(8)]
    mine = pieces[rank::world]
    ok = True
    tmp = Path(os.environ.get("TMPDIR", "/tmp"))
    for k, c in ((21, 2), (31, 1), (4, 5), (12, 3), (33, 2)):
        out = tmp / f"sharded_{k}_{c}.tsv"
        if rank == 0 and out.exists():
            out.unlink()
        dist.barrier()
        mcd.count_sample_sharded(engine, mine, k, c, dist, dev, out_path=out, basename="s")
        dist.barrier()
        if rank == 0:
            want = orc.merge_counts(orc.find_kmers_text(p.decode(), k, c) for p in pieces)
            got = out.read_bytes() if out.exists() else b""
            exp = orc.tsv_bytes("s", want) if want else b""
            good = got == exp
            ok &= good
            print(f"k={k} c={c}: rows {len(want)} {'OK' if good else 'MISMATCH'}", flush=True)
    # one piece split by position: keys are exchanged before the filter
    whole = b"".join(reads)
    mine_text = mcd.split_at_headers(whole, world)[rank]
    for k, c in ((21, 2), (31, 3), (12, 4)):
        out = tmp / f"position_{k}_{c}.tsv"
        if rank == 0 and out.exists():
            out.unlink()
        dist.barrier()
        mcd.count_piece_position_sharded(engine, mine_text, k, c, dist, dev, out_path=out, basename="s")
        dist.barrier()
        if rank == 0:
            want = orc.find_kmers_text(whole.decode(), k, c)
            got = out.read_bytes() if out.exists() else b""
            exp = orc.tsv_bytes("s", want) if want else b""
            good = got == exp
            ok &= good
            print(f"position-sharded k={k} c={c}: rows {len(want)} {'OK' if good else 'MISMATCH'}", flush=True)
    # throughput of the position-sharded path on synthetic reads (per rank: 4 M reads = 0.6 Gbp)
    import bench
    genomes = bench.make_genomes(dev, 0.05)
    big = bench.make_reads_text(dev, genomes, 4_000_000, rank * 4_000_000)
    torch.cuda.synchronize()
    for it in range(4):
        dist.barrier()
        t0 = time.perf_counter()
        part = mcd.count_piece_position_sharded(engine, big, 31, 2, dist, dev)
        torch.cuda.synchronize()
        dist.barrier()
        dt = time.perf_counter() - t0
        rows = part.rows
        part.close()
    if rank == 0:
        print(f"position-sharded -c 2: {world} x 0.6 Gbp in {dt * 1e3:.1f} ms = {world * 0.6 / dt:.2f} Gbases/s, rows on rank 0: {rows}", flush=True)
    del big, genomes
    # timing: exchange of a large table of random 62-bit keys
    n = 20_000_000
    g = torch.Generator(device=dev)
    g.manual_seed(100 + rank)
    keys = torch.randint(0, 1 << 62, (n,), dtype=torch.int64, device=dev, generator=g)
    cnts = torch.ones(n, dtype=torch.int64, device=dev)
    torch.cuda.synchronize()
    table = engine.table_from_rows(31, 0, 0, keys.data_ptr(), cnts.data_ptr(), n, True)
    for it in range(3):
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        part = mcd.merge_table_device(engine, table, dist, dev)
        torch.cuda.synchronize()
        dist.barrier()
        dt = time.perf_counter() - t0
        rows = part.rows
        part.close()
    tot = torch.tensor([rows], dtype=torch.int64, device=dev)
    dist.all_reduce(tot)
    if rank == 0:
        print(f"exchange of {world} x {n} rows: {dt * 1e3:.1f} ms, merged rows {int(tot)} ({world * n * 16 / dt / 1e9:.1f} GB/s of rows)", flush=True)
        print("SHARDED CHECK", "PASS" if ok else "FAIL", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
