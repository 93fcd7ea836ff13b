"""world_size-2 gloo test of the multi-rank table merge (host logic only: the reducer used here is a numpy
stand-in owned by the test; on GPUs the reducer is the engine, see test_gpu_parity.py)."""
import os
import sys
import tempfile

import numpy as np
import pytest

from conftest import ROOT
from oracle import mercat2_oracle as orc


def numpy_reducer(parts_k, parts_c, k):
    kk = np.concatenate([np.asarray(p, np.uint8).reshape(-1, k) for p in parts_k]) if parts_k else np.zeros((0, k), np.uint8)
    cc = np.concatenate([np.asarray(p, np.uint64) for p in parts_c]) if parts_c else np.zeros(0, np.uint64)
    if len(cc) == 0:
        return kk, cc
    keys = kk.view(f"S{k}").reshape(-1)
    uniq, inv = np.unique(keys, return_inverse=True)
    sums = np.zeros(len(uniq), np.uint64)
    np.add.at(sums, inv, cc)
    return np.frombuffer(uniq.tobytes(), np.uint8).reshape(-1, k), sums


def _worker(rank, world, port, tmp):
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    import torch.distributed as dist
    from mercat2_b200 import distributed as mcd
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    k, c = 7, 2
    rng = np.random.default_rng(5)
    genome = "".join(rng.choice(list("ACGT"), 4000))
    pieces = [">r%d\n%s\n" % (i, genome[s:s + 120]) for i, s in enumerate(rng.integers(0, 3880, 600))]
    mine = "".join(mcd.shard_round_robin(pieces, rank, world))            # whole pieces per rank
    local = orc.find_kmers_text(mine, k, c)                               # each rank filters its own piece
    kk = np.frombuffer("".join(local).encode(), np.uint8).reshape(-1, k)
    cc = np.array(list(local.values()), np.uint64)
    merged = mcd.merge_tables(kk, cc, k, numpy_reducer, dist)
    gathered = [None] * world
    dist.all_gather_object(gathered, local)
    if rank == 0:
        want = orc.merge_counts(gathered)
        mk, mc = merged
        got = {bytes(r).decode(): int(n) for r, n in zip(mk, mc)}
        assert got == want
        assert [bytes(r) for r in mk] == sorted(bytes(r) for r in mk)      # rank order == sorted order
        assert mcd.tsv_bytes("s", mk, mc) == orc.tsv_bytes("s", want)
        open(os.path.join(tmp, "ok"), "w").write("ok")
    else:
        assert merged is None
    dist.destroy_process_group()


def test_merge_tables_world2():
    import torch.multiprocessing as mp
    with tempfile.TemporaryDirectory() as tmp:
        port = 29500 + os.getpid() % 2000
        mp.spawn(_worker, args=(2, port, tmp), nprocs=2, join=True)
        assert os.path.exists(os.path.join(tmp, "ok"))


def test_sharding_helpers():
    from mercat2_b200 import distributed as mcd
    assert mcd.shard_round_robin(list(range(7)), 1, 3) == [1, 4]
    parts = mcd.shard_lpt([9, 1, 8, 2, 7, 3], 3)
    assert sorted(i for p in parts for i in p) == list(range(6))
    loads = [sum([9, 1, 8, 2, 7, 3][i] for i in p) for p in parts]
    assert max(loads) - min(loads) <= 1


def _encoding_worker(rank, world, port, tmp):
    sys.path.insert(0, str(ROOT))
    import torch.distributed as dist
    from mercat2_b200 import distributed as mcd
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # rank 0 sees a soft-masked (lower-case) stretch, rank 1 plain ACGT: alone they would pick different encodings
    first = [b">a\n" + b"acgtnacgtn" * 50 + b"\n", b">b\n" + b"ACGT" * 4000 + b"\n"][rank]
    alone = mcd.ENC_NT2 if (lambda n, a, u: n == 0 or a * 10 >= n * 9)(*mcd.alphabet_stats(first)) else None
    agreed = mcd.agree_encoding(first, dist)
    got = [None] * world
    dist.all_gather_object(got, (alone, agreed))
    if rank == 0:
        assert got[0][0] is None and got[1][0] == mcd.ENC_NT2           # the ranks disagree on their own ...
        assert got[0][1] == got[1][1] == mcd.ENC_NT2                    # ... and agree on the summed statistics
        open(os.path.join(tmp, "ok"), "w").write("ok")
    dist.destroy_process_group()


def test_ranks_agree_on_encoding_world2():
    import torch.multiprocessing as mp
    with tempfile.TemporaryDirectory() as tmp:
        port = 31500 + os.getpid() % 2000
        mp.spawn(_encoding_worker, args=(2, port, tmp), nprocs=2, join=True)
        assert os.path.exists(os.path.join(tmp, "ok"))


def test_position_sharding_host_helpers():
    from mercat2_b200 import distributed as mcd
    text = b">a\nACGT\n>b\nGGCC\n>c\nTTAA\n>d\nAC\n"
    for world in (1, 2, 3, 4, 7):
        parts = mcd.split_at_headers(text, world)
        assert b"".join(parts) == text and len(parts) == world
        assert all(p == b"" or p.startswith(b">") for p in parts)        # every part starts at a header line
    assert mcd.split_at_headers(b">a", 4) == [b"", b">a", b"", b""]      # (tiny text: no cut at byte 1)
    assert mcd._prefix_text(0, 16) == b"A" * 16 and mcd._prefix_text(0xFFFFFFFF, 16) == b"T" * 16
    assert mcd._prefix_text(0b00011011 << 24, 4) == b"ACGT"
    assert mcd.alphabet_stats(b">h ACGT\nACGTNN\nacgt*\n") == (10, 4, 6)
    assert mcd.decode_key(0b00011011, 4, mcd.ENC_NT2, mcd.KEY_CODE) == b"ACGT"
