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


# =====================================================================================================
# Device-resident exchange (NCCL on engine memory)
# =====================================================================================================
# The functions above bounce tables through numpy (they also serve the gloo CPU tests).  Below, rows never leave
# the GPUs: packed (key, count) rows are cut at splitter keys on the device, exchanged with one NCCL all-to-all per
# array and re-reduced by the engine on the receiving GPU, so each rank ends up owning a disjoint, sorted key range
# (BASELINE north_star item 5: "hash-partitioned by k-mer code with an NCCL all-to-all"); dense small-k tables are
# summed in place with an NCCL reduce.  Rank order = key order, so the per-sample TSV is the concatenation of the
# ranks' parts and every rank writes its own byte range of the file.

ENC_NT2, ENC_AA5, ENC_BYTE = 0, 1, 2
KEY_CODE, KEY_DENSE_AA = 0, 1


class _DeviceMemory:
    """n elements of engine-owned device memory, exposed through the CUDA array interface (no copy)."""

    def __init__(self, ptr: int, n: int, typestr: str):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


def _to_tensor(engine, ptr: int, n: int, device, typestr: str = "<i8"):
    """n words of engine memory as a torch tensor VIEW (no copy): NCCL reads and writes the engine's own buffers.  The
    engine's calls are synchronous at return, so the memory is ready; the tensor must not outlive the engine object
    that owns the memory."""
    import torch
    if not n or not ptr:
        return torch.empty(0, dtype=torch.int64 if typestr == "<i8" else torch.int32, device=device)
    return torch.as_tensor(_DeviceMemory(ptr, n, typestr), device=device)


def decode_key(key: int, k: int, encoding: int, key_kind: int) -> bytes:
    """Text of a packed key (same layouts as csrc/tsv.cuh::tsv_decode)."""
    out = bytearray(k)
    for j in range(k - 1, -1, -1):
        if key_kind == KEY_DENSE_AA:
            out[j] = 65 + key % 26
            key //= 26
        elif encoding == ENC_NT2:
            out[j] = b"ACGT"[key & 3]
            key >>= 2
        elif encoding == ENC_AA5:
            out[j] = 65 + (key & 31)
            key >>= 5
        else:
            out[j] = key & 255
            key >>= 8
    return bytes(out)


def _agree(dist, mine):
    world = dist.get_world_size()
    got = [None] * world
    dist.all_gather_object(got, mine)
    return got


def merge_table_device(engine, table, dist, device):
    """Every rank passes its (already -c filtered) table of ONE sample.  Returns, on every rank, a new table holding
    the merged rows of the key range that rank owns (ranges are disjoint and ascending with the rank).  Packed rows
    travel GPU-to-GPU over NCCL; literal-byte rows (rare) go through the object collectives."""
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    k = table.k
    info = table.info()
    metas = _agree(dist, (k, info["encoding"], info["key_kind"], info["packed_rows"], info["wide_rows"]))
    with_packed = [m for m in metas if m[3] > 0]
    if any(m[0] != k for m in metas):
        raise ValueError("ranks disagree on k")
    if len({(m[1], m[2]) for m in with_packed}) > 1:
        raise ValueError("ranks counted the same sample under different encodings; merge with merge_tables() instead")
    enc, kind = (with_packed[0][1], with_packed[0][2]) if with_packed else (info["encoding"], info["key_kind"])
    keys_ptr, counts_ptr, n = table.device_rows()
    keys = _to_tensor(engine, keys_ptr, n, device)
    counts = _to_tensor(engine, counts_ptr, n, device)
    # splitters from an evenly spaced sample of every rank's sorted keys
    if n:
        m = min(n, 256)
        idx = (torch.arange(m, device=device, dtype=torch.int64) * (n - 1)) // max(m - 1, 1)
        mine = keys[idx].cpu().numpy().view(np.uint64)
    else:
        mine = np.zeros(0, dtype=np.uint64)
    pool = np.sort(np.concatenate(_agree(dist, mine)))
    splitters = np.array([pool[(i * len(pool)) // world] for i in range(1, world)], dtype=np.uint64) if len(pool) else np.zeros(0, np.uint64)
    cuts = [0] + (table.lower_bound(splitters) if len(splitters) else []) + [n]
    cuts += [n] * (world + 1 - len(cuts))
    send = [cuts[d + 1] - cuts[d] for d in range(world)]
    matrix = _agree(dist, send)
    recv = [matrix[src][rank] for src in range(world)]
    rk = torch.empty(sum(recv), dtype=torch.int64, device=device)
    rc = torch.empty(sum(recv), dtype=torch.int64, device=device)
    if world > 1:
        dist.all_to_all_single(rk, keys, recv, send)
        dist.all_to_all_single(rc, counts, recv, send)
    else:
        rk.copy_(keys)
        rc.copy_(counts)
    # literal-byte rows: destination = number of splitter texts <= row text (consistent with the packed cuts)
    wk = wc = None
    if any(m[4] > 0 for m in metas):
        sp_text = np.array([decode_key(int(x), k, enc, kind) for x in splitters], dtype=f"S{k}")
        my_wk, my_wc = table.wide_arrays() if info["wide_rows"] else (np.zeros((0, k), np.uint8), np.zeros(0, np.uint64))
        dest = np.searchsorted(sp_text, _as_keys(my_wk, k), side="right") if len(sp_text) else np.zeros(len(my_wc), np.int64)
        outgoing = [(my_wk[dest == d], my_wc[dest == d]) for d in range(world)]
        incoming = [x[rank] for x in _agree(dist, outgoing)]
        wk = np.concatenate([x[0].reshape(-1, k) for x in incoming])
        wc = np.concatenate([x[1] for x in incoming])
    torch.cuda.synchronize(device)
    return engine.table_from_rows(k, enc, kind, rk.data_ptr(), rc.data_ptr(), int(rk.numel()), True, wk, wc)


def reduce_dense_sample(engine, sample, dist, device, root: int = 0) -> bool:
    """If the sample is on the dense path on the ranks that saw text, sum the per-sample tables onto `root` with one
    NCCL reduce (in place) and return True; return False (nothing done) when the sample is not dense everywhere."""
    import torch
    ptr, bins, enc = sample.dense()
    metas = _agree(dist, (bins, enc))
    dense = [m for m in metas if m[0] > 0]
    if not dense or len({(m[0], m[1]) for m in dense}) > 1:
        return False
    if any(m[0] == 0 and m[1] >= 0 for m in metas):          # some rank counted its pieces on a non-dense path
        return False
    mismatch = False
    if bins == 0:                                   # this rank saw no text: contribute zeros
        sample.dense_plan(dense[0][1])
        ptr, bins, enc = sample.dense()
        mismatch = bins != dense[0][0]
    if any(_agree(dist, mismatch)):                 # (decided together: no rank may leave before the collective)
        raise RuntimeError("dense plan mismatch between ranks")
    if dist.get_world_size() > 1:
        t = _to_tensor(engine, ptr, bins, device)                 # a view: NCCL reduces the engine's table in place
        dist.reduce(t, dst=root, op=dist.ReduceOp.SUM)
        torch.cuda.synchronize(device)
    return True


def write_tsv_sharded(part, path, basename: str, dist) -> bool:
    """Write the per-sample TSV (bin/mercat2.py:130-133) from the per-rank key ranges of merge_table_device: every
    rank formats its rows on its GPU and writes them at its own offset.  No file when no rank has a row."""
    import os
    rank = dist.get_rank()
    body = part.tsv_body() if part.rows else b""
    sizes = _agree(dist, len(body))
    if sum(sizes) == 0:
        return False
    head = ("k-mer\t%s_Count\n" % basename).encode()
    if rank == 0:
        with open(path, "wb") as f:
            f.write(head)
            f.truncate(len(head) + sum(sizes))
    dist.barrier()
    if body:
        fd = os.open(path, os.O_WRONLY)
        try:
            os.pwrite(fd, body, len(head) + sum(sizes[:rank]))
        finally:
            os.close(fd)
    dist.barrier()
    return True


def count_sample_sharded(engine, pieces, k: int, min_count: int, dist, device, out_path=None, basename: str = "sample",
                         chunk_bytes: int = 0):
    """One sample whose pieces (FASTA texts, each filtered on its own like a chunk file) are spread over the ranks:
    count the local pieces, merge across GPUs on the device, optionally write the TSV.  Returns the rank's table part
    (sparse: its key range; dense: the full table on rank 0, None elsewhere)."""
    # every rank picks the sample's encoding (and with it dense / sparse / literal counting) from its OWN first piece;
    # soft-masked or N-rich stretches could make ranks disagree, so the decision is taken once for all ranks from the
    # summed alphabet statistics of every rank's first piece and forced on the engine while the sample is counted
    enc = agree_encoding(pieces[0] if len(pieces) else b"", dist)
    engine.set_option("force_encoding", enc)
    try:
        sample = engine.sample(k, min_count)
        for text in pieces:
            sample.add_text(text, chunk_bytes)
    finally:
        engine.set_option("force_encoding", -1)
    if reduce_dense_sample(engine, sample, dist, device):
        # rank 0 now holds the summed dense table; the literal-byte rows (windows outside the alphabet) of the other
        # ranks are few and follow through the object collective
        rank = dist.get_rank()
        table = sample.finish()
        info = table.info()
        wide = table.wide_arrays() if info["wide_rows"] else None
        gathered = _agree(dist, wide if rank != 0 else None)
        if rank == 0:
            extra = [g for g in gathered if g is not None and len(g[1])]
            if extra:
                own = [wide] if wide is not None else []
                wk = np.concatenate([x[0].reshape(-1, k) for x in own + extra])
                wc = np.concatenate([x[1] for x in own + extra])
                kp, cp, n = table.device_rows()
                merged = engine.table_from_rows(k, info["encoding"], info["key_kind"], kp, cp, n, True, wk, wc)
                table.close()
                table = merged
            if out_path:
                table.write_tsv(out_path, basename)
        else:
            table.close()
            table = None
        dist.barrier()
        return table
    table = sample.finish()
    part = merge_table_device(engine, table, dist, device)
    table.close()
    if out_path:
        write_tsv_sharded(part, out_path, basename, dist)
    return part


# =====================================================================================================
# One piece split across GPUs before the filter (SURVEY.md 8e grain 3)
# =====================================================================================================
def alphabet_stats(text, limit: int = 1 << 20) -> tuple:
    """(sequence bytes, of them ACGT, of them 'A'..'Z') over the first `limit` bytes of a FASTA text (header lines
    skipped) -- the statistics the engine's own plan uses (csrc/host_reduce.inl make_plan)."""
    if hasattr(text, "is_cuda"):
        text = bytes(text[:limit].cpu().numpy().tobytes())
    head = bytes(memoryview(text)[:limit]) if not isinstance(text, (bytes, bytearray)) else bytes(text[:limit])
    body = b"".join(line for line in head.splitlines() if not line.lstrip().startswith(b">"))
    a = np.frombuffer(body.replace(b"*", b""), dtype=np.uint8)
    counts = np.bincount(a, minlength=256)
    return int(len(a)), int(counts[[65, 67, 71, 84]].sum()), int(counts[65:91].sum())


def agree_encoding(first_piece, dist) -> int:
    """One encoding for all ranks (0 = 2-bit ACGT, 1 = 5-bit A-Z, 2 = byte), by the engine's rule on the summed statistics."""
    stats = _agree(dist, alphabet_stats(first_piece))
    n, acgt, upper = (sum(s[i] for s in stats) for i in range(3))
    if n == 0 or acgt * 10 >= n * 9:
        return ENC_NT2
    return ENC_AA5 if upper * 2 >= n else ENC_BYTE


def split_at_headers(text: bytes, world: int) -> list:
    """Cut one FASTA text into `world` byte ranges of similar size, each starting at a header line (so that no
    window crosses a cut).  Host helper for callers that hold the whole text; ranks that read their own byte range
    of a file do the same with two seeks."""
    n = len(text)
    cuts = [0]
    for r in range(1, world):
        at = max(cuts[-1], (n * r) // world)
        if at == 0:                                  # (tiny text: nothing to cut before the first byte)
            cuts.append(0)
            continue
        pos = text.find(b"\n>", at - 1)
        cuts.append(n if pos < 0 else pos + 1)
    cuts.append(n)
    return [text[cuts[i]:cuts[i + 1]] for i in range(world)]


def _sum_rows(parts_k, parts_c, k):
    """numpy reduce-by-key of literal-byte rows (few)."""
    kk = np.concatenate([np.asarray(x, np.uint8).reshape(-1, k) for x in parts_k]) if parts_k else np.zeros((0, k), np.uint8)
    cc = np.concatenate([np.asarray(x, np.uint64) for x in parts_c]) if parts_c else np.zeros(0, np.uint64)
    if not len(cc):
        return kk, cc
    keys = _as_keys(kk, k)
    uniq, inv = np.unique(keys, return_inverse=True)
    sums = np.zeros(len(uniq), dtype=np.uint64)
    np.add.at(sums, inv, cc)
    return np.frombuffer(uniq.tobytes(), dtype=np.uint8).reshape(-1, k), sums


def _prefix_text(p: int, n: int) -> bytes:
    """The first n (<= 16) symbols of the 32-bit key prefix p as text."""
    return bytes(b"ACGT"[(p >> (30 - 2 * j)) & 3] for j in range(n))


def init_nccl(device, dist=None):
    """NCCL process group for one process per GPU, its kernels on a high-priority stream: the exchange of a key range runs
    while the previous range is counted by kernels that fill the GPU, and the exchange's CTAs should not queue behind
    them.  (Measured at 8 ranks: neutral, 439 vs 442 ms per step -- the all-to-all itself, ~230 GB/s per direction while
    the counting kernels run, is what bounds the counting phase there)"""
    if dist is None:
        import torch.distributed as dist
    options = None
    try:
        options = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
    except Exception:                                    # (a torch build without the option: default priority)
        options = None
    dist.init_process_group("nccl", device_id=device, pg_options=options)
    return dist


def count_piece_position_sharded(engine, my_text, k: int, min_count: int, dist, device, out_path=None,
                                 basename: str = "sample", groups_per_rank: int | None = None, timings: dict | None = None):
    """ONE piece (chunk) whose text is split by position over the ranks (each part starts at a header line).  The
    `-c` filter must see whole-piece counts (lib/mercat2_kmers.py:73-78 on the unsplit piece), so keys are exchanged
    BEFORE counting:

    1. every rank packs its text and samples its keys' prefixes; the samples are summed with one NCCL all-reduce (on the
       engine's buffer) so that all ranks cut the key space at the same world * groups_per_rank boundaries;
    2. every rank groups the order-preserving 64-bit keys of all its windows by key range on its GPU;
    3. in groups_per_rank rounds, round j moves range (r, j) to rank r with one NCCL all-to-all straight out of the
       engine's key array; the receiver counts round j (all occurrences of that range: the filter is exact) while round
       j + 1 is already in flight;
    4. rank r ends up with the sorted rows of the r-th slice of the key space: rank order == key order, the table is the
       concatenation of the ranks' parts and nothing is sorted or merged afterwards.

    The (few) windows outside ACGT are counted unfiltered, summed across ranks, filtered, and each row joins the rank
    whose key range it sorts into.  Returns this rank's part of the final table.  `timings` (optional dict) receives
    per-phase milliseconds and the bytes this rank sent / received."""
    import time
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    if groups_per_rank is None:
        # rounds of about 1.7e8 keys: one pass of the engine's two-level partition (a round beyond that takes the level-0
        # route: one more pass over its keys); ~0.75 windows per text byte for short reads
        nbytes = max(_agree(dist, len(my_text) if not hasattr(my_text, "numel") else int(my_text.numel()))) if world > 1 else \
            (len(my_text) if not hasattr(my_text, "numel") else int(my_text.numel()))
        groups_per_rank = min(max(4, -(-int(nbytes * 0.75) // 170_000_000)), max(1, 384 // world))
    m = groups_per_rank
    groups = world * m
    t_last = [time.perf_counter()]

    def mark(what):
        if timings is not None:
            torch.cuda.synchronize(device)
            now = time.perf_counter()
            timings[what] = timings.get(what, 0.0) + (now - t_last[0]) * 1e3
            t_last[0] = now

    keys = engine.open_keys(my_text, k)
    mark("parse_ms")
    if world > 1:
        hist = _to_tensor(engine, keys.sample_ptr, keys.sample_entries, device, "<i4")
        dist.all_reduce(hist)
        torch.cuda.synchronize(device)
    keys.partition(groups)
    mark("partition_ms")
    sizes = torch.tensor(keys.sizes, dtype=torch.int64, device=device)
    if world > 1:
        all_sizes = torch.empty((world, groups), dtype=torch.int64, device=device)
        dist.all_gather_into_tensor(all_sizes, sizes)
        exc = torch.tensor([keys.exception_symbols], dtype=torch.int64, device=device)
        dist.all_reduce(exc)
        any_exceptions = int(exc.item()) > 0
    else:
        all_sizes = sizes.view(1, groups)
        any_exceptions = keys.exception_symbols > 0
    all_sizes = all_sizes.cpu().tolist()
    offsets = [0]
    for n in keys.sizes:
        offsets.append(offsets[-1] + n)
    send_all = _to_tensor(engine, keys.ptr, keys.total, device)
    sample = engine.sample(k, min_count)
    recv_max = max(sum(all_sizes[src][rank * m + j] for src in range(world)) for j in range(m))
    bufs = [torch.empty(max(recv_max, 1), dtype=torch.int64, device=device) for _ in range(2 if world > 1 else 0)]
    sent = received = 0

    def start_round(j):
        """all-to-all of range (r, j) -> rank r; returns (work, tensor holding this rank's range, its key count)"""
        n_in = [all_sizes[src][rank * m + j] for src in range(world)]
        if world == 1:
            g = rank * m + j
            return None, send_all[offsets[g]:offsets[g + 1]], n_in[0]
        buf = bufs[j & 1]
        outs, at = [], 0
        for n in n_in:
            outs.append(buf[at:at + n])
            at += n
        ins = [send_all[offsets[r * m + j]:offsets[r * m + j + 1]] for r in range(world)]
        work = dist.all_to_all(outs, ins, async_op=True)
        return work, buf[:at], at

    # the exchange's NCCL kernels hold SMs while a round is counted: the engine's persistent kernels launch several waves
    # of CTAs so that a late CTA costs a fraction of the kernel, not a second pass (see "grid_waves" in host_count.inl)
    if world > 1:
        engine.set_option("grid_waves", 4)
    try:
        pending = start_round(0)
        for j in range(m):
            work, part, n = pending
            if work is not None:
                work.wait()
                torch.cuda.current_stream(device).synchronize()
            mark("exchange_wait_ms")
            if j + 1 < m:
                pending = start_round(j + 1)          # in flight while round j is counted
            g = rank * m + j
            sent += sum(keys.sizes[r * m + j] for r in range(world) if r != rank) * 8
            received += sum(all_sizes[src][g] for src in range(world) if src != rank) * 8
            if n:
                sample.add_keys(part.data_ptr(), n, True, keys.bounds[g], max(keys.bounds[g + 1], keys.bounds[g] + 1))
            mark("count_ms")
    finally:
        if world > 1:
            engine.set_option("grid_waves", 0)        # (0 = the engine's default)
    bounds = keys.bounds
    del send_all
    keys.close()
    if any_exceptions:                               # literal-byte windows: unfiltered local tables -> sum -> filter -> owner
        t = engine.count_exceptions(my_text, k)
        wk, wc = t.wide_arrays()
        t.close()
        gathered = _agree(dist, (wk, wc)) if world > 1 else [(wk, wc)]
        sk, sc = _sum_rows([g[0] for g in gathered], [g[1] for g in gathered], k)
        keep = sc >= np.uint64(max(1, min_count))
        sk, sc = sk[keep], sc[keep]
        if len(sc):
            kk = min(k, 16)
            cut_text = np.array([_prefix_text(bounds[r * m], kk) for r in range(1, world)], dtype=f"S{kk}")
            heads = _as_keys(np.ascontiguousarray(sk[:, :kk]), kk)
            dest = np.searchsorted(cut_text, heads, side="right") if world > 1 else np.zeros(len(sc), np.int64)
            mine = dest == rank
            if mine.any():
                sample.add_rows(sk[mine], sc[mine])
    table = sample.finish()
    mark("finish_ms")
    if timings is not None:
        timings["nvlink_bytes_sent"] = timings.get("nvlink_bytes_sent", 0) + sent
        timings["nvlink_bytes_received"] = timings.get("nvlink_bytes_received", 0) + received
    if out_path:
        write_tsv_sharded(table, out_path, basename, dist)
    return table
