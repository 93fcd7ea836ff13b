"""Multi-GPU glue: one process per GPU, whole samples / whole pieces per rank (no data-path collective), and
one exchange step when several ranks hold parts of the SAME sample: their already-filtered tables are
partitioned by k-mer range, exchanged all-to-all, summed per range on each GPU and concatenated in rank order
(= sorted order).  This replaces the serial dict merge of ``run_mercat2`` (bin/mercat2.py:121-127) across
processes.  The exchange runs over ``torch.distributed`` (NCCL between GPUs; gloo works for the CPU tests)."""
from __future__ import annotations

import numpy as np


def shard_round_robin(items, rank: int, world: int):
    """Whole pieces / samples per rank (exact: a piece is filtered on its own, lib/mercat2_kmers.py:73-78)."""
    return [x for i, x in enumerate(items) if i % world == rank]


def shard_lpt(sizes, world: int):
    """Longest-processing-time assignment of samples to ranks -> list of index lists."""
    loads = [0] * world
    out = [[] for _ in range(world)]
    for idx in sorted(range(len(sizes)), key=lambda i: -sizes[i]):
        r = loads.index(min(loads))
        out[r].append(idx)
        loads[r] += sizes[idx]
    return out


def _as_keys(kmers: np.ndarray, k: int) -> np.ndarray:
    """uint8[rows, k] -> fixed-width byte strings (numpy compares them like Python bytes)."""
    return np.ascontiguousarray(kmers, dtype=np.uint8).reshape(-1, k).view(f"S{k}").reshape(-1)


def choose_splitters(keys: np.ndarray, world: int, dist) -> list:
    """world-1 k-mers that split the union of all ranks' (sorted) keys into ranges of similar size."""
    take = min(len(keys), 256)
    sample = keys[np.linspace(0, len(keys) - 1, take).astype(np.int64)].tolist() if take else []
    gathered = [None] * world
    dist.all_gather_object(gathered, sample)
    pool = sorted(x for part in gathered for x in part)
    if not pool:
        return []
    return [pool[(i * len(pool)) // world] for i in range(1, world)]


def exchange_rows(send_k: list, send_c: list, k: int, dist, device=None):
    """all-to-all of per-destination (kmers uint8[n,k], counts uint64[n]) -> lists indexed by source rank."""
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    counts_out = [len(c) for c in send_c]
    counts_in = [None] * world
    all_counts = [None] * world
    dist.all_gather_object(all_counts, counts_out)
    counts_in = [all_counts[src][rank] for src in range(world)]
    use_a2a = dist.get_backend() == "nccl"
    recv_k, recv_c = [], []
    if use_a2a:
        dev = device or torch.device("cuda", torch.cuda.current_device())
        sk = torch.from_numpy(np.concatenate([np.asarray(x, np.uint8).reshape(-1, k) for x in send_k])).to(dev)
        sc = torch.from_numpy(np.concatenate([np.asarray(x, np.uint64) for x in send_c]).view(np.int64)).to(dev)
        rk = torch.empty((sum(counts_in), k), dtype=torch.uint8, device=dev)
        rc = torch.empty(sum(counts_in), dtype=torch.int64, device=dev)
        dist.all_to_all_single(rk, sk, [n for n in counts_in], [n for n in counts_out])
        dist.all_to_all_single(rc, sc, counts_in, counts_out)
        rk, rc = rk.cpu().numpy(), rc.cpu().numpy().view(np.uint64)
        at = 0
        for n in counts_in:
            recv_k.append(rk[at:at + n])
            recv_c.append(rc[at:at + n])
            at += n
    else:                                    # gloo: pairwise sends (CPU tests)
        reqs, bufs = [], []
        for src in range(world):
            bk = torch.empty((counts_in[src], k), dtype=torch.uint8)
            bc = torch.empty(counts_in[src], dtype=torch.int64)
            bufs.append((bk, bc))
            if src != rank and counts_in[src]:
                reqs += [dist.irecv(bk, src=src), dist.irecv(bc, src=src)]
        for dst in range(world):
            if dst != rank and counts_out[dst]:
                tk = torch.from_numpy(np.ascontiguousarray(send_k[dst], np.uint8).reshape(-1, k))
                tc = torch.from_numpy(np.ascontiguousarray(send_c[dst], np.uint64).view(np.int64))
                reqs += [dist.isend(tk, dst=dst), dist.isend(tc, dst=dst)]
        for r in reqs:
            r.wait()
        for src in range(world):
            if src == rank:
                recv_k.append(np.asarray(send_k[rank], np.uint8).reshape(-1, k))
                recv_c.append(np.asarray(send_c[rank], np.uint64))
            else:
                recv_k.append(bufs[src][0].numpy())
                recv_c.append(bufs[src][1].numpy().view(np.uint64))
    return recv_k, recv_c


def engine_reducer(engine):
    """Sum equal k-mers on the GPU (device sort + reduce-by-key through the engine)."""
    def reduce(parts_k, parts_c, k):
        sample = engine.sample(k, 1)
        for pk, pc in zip(parts_k, parts_c):
            if len(pc):
                sample.add_rows(pk, pc)
        return sample.finish().arrays()
    return reduce


def merge_tables(kmers: np.ndarray, counts: np.ndarray, k: int, reducer, dist=None, root: int = 0):
    """Merge the per-rank tables of ONE sample.  Every rank passes its (already -c filtered) rows; returns the
    merged, sorted (kmers, counts) on ``root`` and ``None`` elsewhere.  ``reducer(parts_k, parts_c, k)`` sums equal
    k-mers of the received parts (``engine_reducer(engine)`` on GPUs)."""
    if dist is None:
        import torch.distributed as dist
    world, rank = dist.get_world_size(), dist.get_rank()
    kmers = np.ascontiguousarray(kmers, dtype=np.uint8).reshape(-1, k)
    counts = np.ascontiguousarray(counts, dtype=np.uint64)
    keys = _as_keys(kmers, k)
    order = np.argsort(keys, kind="stable")
    keys, kmers, counts = keys[order], kmers[order], counts[order]
    splitters = choose_splitters(keys, world, dist)
    cuts = [0] + [int(np.searchsorted(keys, np.array(s, dtype=keys.dtype), side="left")) for s in splitters]
    cuts += [len(keys)] * (world + 1 - len(cuts))
    send_k = [kmers[cuts[d]:cuts[d + 1]] for d in range(world)]
    send_c = [counts[cuts[d]:cuts[d + 1]] for d in range(world)]
    recv_k, recv_c = exchange_rows(send_k, send_c, k, dist)
    mk, mc = reducer(recv_k, recv_c, k)                       # this rank's key range, sorted
    gathered = [None] * world if rank == root else None
    dist.gather_object((np.asarray(mk, np.uint8).reshape(-1, k), np.asarray(mc, np.uint64)), gathered, dst=root)
    if rank != root:
        return None
    return (np.concatenate([g[0] for g in gathered]), np.concatenate([g[1] for g in gathered]))


def tsv_bytes(basename: str, kmers: np.ndarray, counts: np.ndarray) -> bytes:
    """The per-sample TSV (bin/mercat2.py:130-133) from merged arrays."""
    k = kmers.shape[1] if kmers.ndim == 2 and len(kmers) else 0
    text = kmers.tobytes().decode("ascii")
    rows = ["k-mer\t%s_Count\n" % basename]
    rows += ["%s\t%d\n" % (text[i * k:(i + 1) * k], int(c)) for i, c in enumerate(counts.tolist())]
    return "".join(rows).encode()
