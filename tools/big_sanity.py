#!/usr/bin/env python
"""Size-independent checks on inputs beyond 4 GiB (one chunk): sum of counts = number of windows, rows sorted.
   python tools/big_sanity.py"""
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import mercat2_b200  # noqa: E402


def fasta(dev, n_records, rec_len, alphabet, seed, cols=80):
    """n_records records of rec_len symbols each, `cols` per line, header '>s%09d\n' (11 bytes)."""
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    lut = torch.tensor(list(alphabet), dtype=torch.uint8, device=dev)
    lines = rec_len // cols
    rec_bytes = 11 + lines * (cols + 1)
    out = torch.empty(n_records * rec_bytes, dtype=torch.uint8, device=dev)
    view = out.view(n_records, rec_bytes)
    pow10 = torch.tensor([10 ** (8 - j) for j in range(9)], device=dev, dtype=torch.int64)
    B = 4096
    for b0 in range(0, n_records, B):
        nb = min(B, n_records - b0)
        rows = view[b0:b0 + nb]
        ids = torch.arange(b0, b0 + nb, device=dev, dtype=torch.int64)
        rows[:, 0] = ord(">")
        rows[:, 1] = ord("s")
        rows[:, 2:11 - 0 - 0][:, :9] = ((ids[:, None] // pow10[None, :]) % 10 + 48).to(torch.uint8)
        rows[:, 10] = 10
        body = rows[:, 11:].view(nb, lines, cols + 1)
        codes = torch.randint(0, len(alphabet), (nb, lines, cols), device=dev, generator=g)
        body[:, :, :cols] = lut[codes]
        body[:, :, cols] = 10
    return out, lines * cols


def main():
    dev = torch.device("cuda:0")
    engine = mercat2_b200.Engine(0)
    ok = True
    cases = [("nucleotide 4.6 Gbp, k=3 (dense, 64-bit fold)", b"ACGT", 46000, 100000, 3, 10),
             ("nucleotide 4.6 Gbp, k=31 -c 5 (level-0 partition)", b"ACGT", 46000, 100000, 31, 5),
             ("protein 4.4 G residues, k=3 (general parser, dense)", b"ACDEFGHIKLMNPQRSTVWY", 44000, 100000, 3, 10)]
    for name, alpha, nrec, rlen, k, c in cases:
        text, per = fasta(dev, nrec, rlen, alpha, 7)
        windows = nrec * (per - k + 1)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        table, offs = engine.count_sample(text, k, c, 0)
        dt = time.perf_counter() - t0
        rows, total = table.rows, table.total
        good = True
        if k == 3:
            good = total == windows and rows == len(alpha) ** 3
            kmers, _ = table.arrays()
            keys = [bytes(r) for r in kmers]
            good = good and keys == sorted(keys)
        else:
            good = rows == 0 and total == 0          # random 31-mers: nothing repeats 5 times
        ok &= good
        print(f"{name}: {text.numel() / 2**30:.2f} GiB, {dt:.2f} s, rows {rows}, sum {total} (windows {windows}) {'OK' if good else 'FAIL'}", flush=True)
        table.close()
        del text
        torch.cuda.empty_cache()
    print("BIG SANITY", "PASS" if ok else "FAIL")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
