#!/usr/bin/env python
"""bench.py -- input bases/s of the k-mer counting hot path on B200.

Headline workload = BASELINE.json config 4 AS WRITTEN (SURVEY 8d run (i)): synthetic 150-bp metagenome reads (200
genomes x 5 Mbp, 0.1 % substitutions, FASTA framing '>r%010d\\n' + 150 bases + '\\n' = 164 B per read), nucleotide
k=31, ONE global table: `-s 0 -c 2` (no chunking: the min-count filter sees whole-sample counts).  Every rank owns
reads_per_gpu reads (default 12.5 Gbp, so N=8 is exactly the named 100 Gbp; weak scaling).  A step is one pass of
the hot path over the job's reads:

  N = 1   text -> packed symbols -> order-preserving 64-bit keys, range-partitioned in HBM -> shared-memory tables ->
          -c filter -> the sorted table (rows born sorted: nothing is sorted afterwards)
  N > 1   the same with the key space cut into N slices: every rank partitions the keys of ITS reads by key range,
          NCCL all-to-all (before the filter) moves each slice to its owner, the owner counts it; rank order = key
          order, the job's table is the concatenation of the ranks' parts (bin/mercat2.py:119-127's reducer)

  value      whole-job bases/s with the FASTA text already resident in HBM
  e2e        the same through the public API with the text in pinned HOST memory (H2D inside the timed region) and
             the table (64-bit key + count per row) read back into pinned host memory (D2H)
  roofline   dominant kernel of the timed region against the measured HBM peak; pipeline_* = the whole path against
             SURVEY 8(d)'s algorithmic bytes
  phases     per-phase milliseconds of a step (parse / partition / exchange / count / emit)
  check      table digest on a CPU-sized prefix, engine vs oracle, same pieces and flags
  secondary  the per-chunk configuration (`-s 100 -c 10`, the reference's defaults) on the same reads

`--impl reference` times the CPU reference arm (the oracle port of the reference's Python counter on all host cores;
/root/reference does not exist on the GPU box) on a bounded prefix of the same reads.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import multiprocessing
import os
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

READ_LEN = 150
REC_BYTES = 13 + READ_LEN + 1          # '>r%010d\n' + bases + '\n'
N_GENOMES, GENOME_LEN = 200, 5_000_000
SEED = 20240531
METRIC = "input bases/sec (k-mers counted/sec) per GPU and 8xB200; % of HBM roofline"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--reads-per-gpu", type=int, default=int(os.environ.get("MC2_BENCH_READS", 83_333_334)))
    ap.add_argument("-k", type=int, default=31)
    ap.add_argument("-c", type=int, default=2)
    ap.add_argument("-s", type=int, default=0, help="chunk size in MB (0 = one global table, the headline)")
    ap.add_argument("--cpu-reads", type=int, default=0, help="reads in the CPU sample (0 = auto)")
    ap.add_argument("--genome-scale", type=float, default=1.0, help="fraction of the 1 Gbp metagenome to synthesise (profiling runs)")
    ap.add_argument("--groups-per-rank", type=int, default=0, help="N > 1: exchange rounds (key ranges per rank); 0 = sized from the text")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-secondary", action="store_true")
    ap.add_argument("--no-check", action="store_true")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------
# synthetic reads (torch is plumbing here: device memory + RNG)
# ---------------------------------------------------------------------------------------------------
def make_genomes(device, scale=1.0):
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(SEED)
    n = int(N_GENOMES * GENOME_LEN * scale)
    return torch.randint(0, 4, (n,), dtype=torch.uint8, device=device, generator=g)


def make_reads_text(device, genomes, n_reads, first_read_id, out=None):
    """FASTA text of n_reads reads as a uint8 tensor on `device` (deterministic in first_read_id)."""
    import torch
    text = out if out is not None else torch.empty(n_reads * REC_BYTES, dtype=torch.uint8, device=device)
    view = text.view(n_reads, REC_BYTES)
    lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=device)
    ar = torch.arange(READ_LEN, device=device)
    pow10 = torch.tensor([10 ** (9 - j) for j in range(10)], device=device, dtype=torch.int64)
    B = 1 << 18
    limit = genomes.numel() - READ_LEN
    for b0 in range(0, n_reads, B):
        nb = min(B, n_reads - b0)
        g = torch.Generator(device=device)
        g.manual_seed(SEED + 1 + first_read_id + b0)
        start = torch.randint(0, limit, (nb,), device=device, generator=g)
        codes = genomes[start[:, None] + ar[None, :]]
        err = torch.rand((nb, READ_LEN), device=device, generator=g) < 0.001
        sub = torch.randint(1, 4, (nb, READ_LEN), device=device, generator=g, dtype=torch.uint8)
        codes = torch.where(err, (codes + sub) & 3, codes)
        rows = view[b0:b0 + nb]
        rows[:, 13:13 + READ_LEN] = lut[codes.long()]
        ids = torch.arange(first_read_id + b0, first_read_id + b0 + nb, device=device, dtype=torch.int64)
        rows[:, 2:12] = ((ids[:, None] // pow10[None, :]) % 10 + 48).to(torch.uint8)
        rows[:, 0] = ord(">")
        rows[:, 1] = ord("r")
        rows[:, 12] = 10
        rows[:, REC_BYTES - 1] = 10
    return text


# ---------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi samples every 200 ms.  The process is started BEFORE the warm-up steps (its start-up enumerates every GPU
    of the box and was seen to slow the first two timed steps by ~60 ms each when started at the timed region's edge); only
    the samples that arrive inside the timed region [t0, t1] are used."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def stop(self, t0=None, t1=None):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
        sm, mx, reasons = [], 0, set()
        inside = [ln for ts, ln in self.lines if (t0 is None or ts >= t0) and (t1 is None or ts <= t1 + 0.2)]
        for line in inside or [ln for _, ln in self.lines[-3:]]:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx = max(mx, float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------
# CPU reference arm (oracle port of the reference's counter, all host cores)
# ---------------------------------------------------------------------------------------------------
def _cpu_task(args):
    from oracle import mercat2_oracle as orc
    path, k, c = args
    return orc.find_kmers(Path(path), k, c)


def cpu_sample_files(text_bytes: bytes, n_tasks: int, workdir: str):
    """Split the sample at record boundaries into n_tasks FASTA files (one task per file, like one
    Ray task per chunk file in bin/mercat2.py:120)."""
    n_reads = len(text_bytes) // REC_BYTES
    per = max(1, n_reads // n_tasks)
    files = []
    for i in range(n_tasks):
        a = i * per * REC_BYTES
        b = len(text_bytes) if i == n_tasks - 1 else (i + 1) * per * REC_BYTES
        if a >= b:
            break
        path = os.path.join(workdir, f"sample.{i:05d}.fna")
        with open(path, "wb") as out:
            out.write(text_bytes[a:b])
        files.append(path)
    return files


def cpu_run(files, k, c, pool, out_tsv):
    """countKmers tasks + serial merge + sorted TSV (bin/mercat2.py:115-137); returns seconds."""
    from oracle import mercat2_oracle as orc
    t0 = time.perf_counter()
    tables = pool.map(_cpu_task, [(f, k, c) for f in files])
    total = orc.merge_counts(tables)
    if total:
        with open(out_tsv, "wb") as out:
            out.write(orc.tsv_bytes("sample", total))
    elif os.path.exists(out_tsv):
        os.remove(out_tsv)
    return time.perf_counter() - t0


def host_cores():
    try:
        import psutil
        n = psutil.cpu_count(logical=False) or os.cpu_count()
    except Exception:
        n = os.cpu_count()
    try:
        n = min(n, len(os.sched_getaffinity(0)))
    except Exception:
        pass
    return max(1, int(n))


def cpu_reads_auto(cores, k):
    # ~0.6 Mbases/s/core at k=31: aim at ~12 s of work per core
    per_core_bases = 7_000_000 if k > 16 else 25_000_000
    return cores * per_core_bases // READ_LEN


def workload_name(k, c, s):
    kind = "one global table (SURVEY 8d run (i))" if s == 0 else "per-chunk filter (run (ii))"
    return f"cfg4: synthetic 150-bp metagenome reads, nucleotide k={k} -c {c} -s {s}: {kind}"


# ---------------------------------------------------------------------------------------------------
# roofline accounting
# ---------------------------------------------------------------------------------------------------
def algorithmic_bytes_per_step(kernel, windows, text_bytes, rows):
    """Algorithmic bytes ALL launches of the named kernel move in one step over this rank's reads: what the kernel
    must read and write once, not what it happens to move.  windows = k-mer windows, rows = surviving rows."""
    symbols = text_bytes * (READ_LEN + 1) / REC_BYTES            # bases + one separator per read
    packed = symbols * 0.375                                     # 2-bit codes + 1 validity bit per symbol
    table = {
        "fn_parse_kernel<0>": text_bytes, "fn_parse_kernel<2>": text_bytes, "fn_parse_kernel<1>": text_bytes + packed,
        "fn_hist_kernel": packed, "fn_hist_kernel<level0>": packed,
        "fn_scatter1_kernel": packed + 8.0 * windows,            # every key written once
        "hk_hist_kernel": 8.0 * windows, "hk_hist_kernel<level0>": 8.0 * windows,
        "hk_scatter1_kernel": 16.0 * windows,                    # every key read once and written once
        "hc_scatter2_kernel": 16.0 * windows,
        "rc_count_kernel<0>": 8.0 * windows + 16.0 * rows,       # every key read once, every surviving row written once
        "rc_count_kernel<1>": 8.0 * windows + 16.0 * rows,
        "rc_gather_kernel": 32.0 * rows,
        "chunk_has_cr_kernel": text_bytes,
    }
    return table.get(kernel)


PHASES = (("parse", ("fn_parse", "chunk_", "scan_")), ("partition", ("rp_", "fn_hist", "fn_scatter1", "hk_", "hc_scan", "hc_scatter2")),
          ("count", ("rc_count",)), ("emit", ("rc_offsets", "rc_gather", "part_ends", "rs_", "seg_", "rle_", "gather_")))


def phases_from_profile(profile, steps):
    out = {name: 0.0 for name, _ in PHASES}
    out["other"] = 0.0
    for kern, v in profile.items():
        for name, prefixes in PHASES:
            if kern.startswith(prefixes):
                out[name] += v["us"]
                break
        else:
            out["other"] += v["us"]
    return {f"{k}_ms": round(v / 1e3 / max(1, steps), 3) for k, v in out.items()}


# ---------------------------------------------------------------------------------------------------
def run_cfg5(engine, device, rank, world, barrier, reps=3):
    """BASELINE config 5 (reference sample loop bin/mercat2.py:411-448): 64 proteomes, one table + one metrics block per
    sample; samples are dealt to the ranks by size (LPT), a rank counts its samples in batched passes."""
    import torch
    from mercat2_b200 import distributed as mcd
    from tools import synth_s5
    texts = synth_s5.sample_set(64, 5000)
    residues = sum(len(t) - t.count(b"\n") - sum(len(h) for h in t.split(b"\n") if h.startswith(b">")) for t in texts)
    mine = mcd.shard_lpt([len(t) for t in texts], world)[rank]
    dev = [torch.frombuffer(bytearray(texts[j]), dtype=torch.uint8).to(device) for j in mine]
    best, rows, proteins, err = None, 0, 0, None
    for rep in range(reps + 1):                                       # (pass 0 warms up)
        torch.cuda.synchronize()
        barrier()
        t0 = time.perf_counter()
        try:
            tables = engine.count_batch(dev, 5, 10) if dev else []
            proteins = sum(len(engine.protein_metrics(t)["length"]) for t in dev)
            rows = sum(t.rows for t in tables)
            for t in tables:
                t.close()
        except Exception as exc:                                      # (keep the barriers paired on every rank)
            err = exc
        torch.cuda.synchronize()
        barrier()
        dt = time.perf_counter() - t0
        if rep:
            best = dt if best is None else min(best, dt)
    if err is not None:
        raise err
    return {"seconds": best, "rows": rows, "proteins": proteins, "residues": residues}


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank != 0:
            return 0
        return reference_arm(args, world)

    import numpy as np
    import torch
    import torch.distributed as dist
    import mercat2_b200
    from mercat2_b200 import distributed as mcd

    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    numa_bound = mercat2_b200.bind_to_gpu_numa(local) if world > 1 else False     # pinned shards next to their GPU
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        mcd.init_nccl(device, dist)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    n_reads = args.reads_per_gpu
    bases_per_step = n_reads * READ_LEN
    genomes = make_genomes(device, args.genome_scale)
    text = make_reads_text(device, genomes, n_reads, rank * n_reads)
    torch.cuda.synchronize()
    nbytes = text.numel()
    windows = n_reads * (READ_LEN - args.k + 1)

    engine = mercat2_b200.Engine(local)
    for item in filter(None, os.environ.get("MC2_BENCH_OPTIONS", "").split(",")):       # experiments: "name=value,..." engine tunables
        name, _, value = item.partition("=")
        engine.set_option(name.strip(), int(value))

    def count(buf, k, c, s_mb, timings=None):
        """one pass of the hot path over this rank's reads -> this rank's part of the job's table"""
        chunk_bytes = s_mb * 1024 * 1024 if (s_mb > 0 and nbytes >= s_mb * 1024 * 1024) else 0
        if s_mb == 0 and world > 1:
            # ONE piece whose text is spread over the ranks: keys exchanged before the filter
            return mcd.count_piece_position_sharded(engine, buf, k, c, dist, device, groups_per_rank=args.groups_per_rank or None, timings=timings), 1
        table, offsets = engine.count_sample(buf, k, c, chunk_bytes)
        if world > 1:                       # whole pieces per rank, filtered per piece: sum the filtered tables by key range
            part = mcd.merge_table_device(engine, table, dist, device)
            table.close()
            table = part
        return table, len(offsets)

    def step_resident(timings=None):
        table, n_chunks = count(text, args.k, args.c, args.s, timings)
        rows = table.rows
        table.close()
        return rows, n_chunks

    # ---- correctness gate before anything is timed ---------------------------------------------------
    check = None
    if not args.no_check:
        check = run_check(engine, genomes, args, rank, world, dist if world > 1 else None, device)

    sampler = ClockSampler(local)
    sampler.start()
    warm_ms = []
    for _ in range(args.warmup):
        tw = time.perf_counter()
        rows, n_chunks = step_resident()
        warm_ms.append(round((time.perf_counter() - tw) * 1e3, 2))
    launches0 = engine.stat("launches")
    ovf0, merges0 = engine.stat("overflow_buckets"), engine.stat("row_merges")
    # CUDA events on the engine's stream around every kernel of the path (MC2_BENCH_PROFILE=3: only the long ones)
    engine.set_option("profile", int(os.environ.get("MC2_BENCH_PROFILE", "2")))
    timings = {}
    barrier()
    t0 = time.perf_counter()
    step_ms = []
    for _ in range(args.steps):
        rows, n_chunks = step_resident(timings if world > 1 and args.s == 0 else None)
        step_ms.append(time.perf_counter())                           # (a step returns with its table finished: no extra synchronisation)
    barrier()
    elapsed = time.perf_counter() - t0
    step_ms = [round((b - a) * 1e3, 2) for a, b in zip([t0] + step_ms[:-1], step_ms)]
    clocks = sampler.stop(t0, t0 + elapsed)
    launches = engine.stat("launches") - launches0
    overflow_buckets = engine.stat("overflow_buckets") - ovf0
    profile = engine.profile()
    engine.set_option("profile", 0)
    kernel_us = sum(v["us"] for v in profile.values())

    # ---- secondary: the per-chunk configuration on the same reads --------------------------------------
    secondary = None
    if not args.no_secondary and args.s == 0:
        s2, c2 = 100, 10
        for _ in range(2):
            t2, nch2 = count(text, args.k, c2, s2)
            rows2 = t2.rows
            t2.close()
        engine.set_option("profile", 2)
        barrier()
        t1 = time.perf_counter()
        sec_steps = max(2, min(args.steps, 3))
        for _ in range(sec_steps):
            t2, nch2 = count(text, args.k, c2, s2)
            rows2 = t2.rows
            t2.close()
        barrier()
        sec_elapsed = time.perf_counter() - t1
        prof2 = engine.profile()
        engine.set_option("profile", 0)
        secondary = {"elapsed": sec_elapsed, "steps": sec_steps, "rows": rows2, "chunks": nch2, "profile": prof2, "s": s2, "c": c2}

    # ---- cfg5: 64 synthetic proteomes (BASELINE config 5), whole samples per rank (LPT), batched passes ----
    cfg5 = None
    if not args.no_secondary:
        try:
            cfg5 = run_cfg5(engine, device, rank, world, barrier)
        except Exception as exc:                                      # (reported in the line, never fatal for the headline)
            cfg5 = {"error": f"{type(exc).__name__}: {exc}"[:200]}

    # ---- e2e: host buffer in, table out, through the same public call ----------------------------------
    e2e = None
    if not args.no_e2e:
        import psutil
        out_rows_cap = int(rows * 1.1) + 1024
        need = nbytes + 16 * out_rows_cap
        fits = [psutil.virtual_memory().available > int(1.25 * need * world) + (16 << 30)]
        if world > 1:
            dist.broadcast_object_list(fits, src=0)
        if fits[0]:
            host = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
            host.copy_(text)
            out_rows = torch.empty((out_rows_cap, 2), dtype=torch.int64, pin_memory=True)      # (key, count) per row
            split_rows = world == 1 and args.s == 0 and os.environ.get("MC2_BENCH_ROWS16", "0") != "1"
            torch.cuda.synchronize()
            e_steps = max(1, min(args.steps, 5))

            def step_e2e():
                if world == 1 and args.s == 0:
                    # the public call that delivers the table to host memory: rows of finished key ranges stream out
                    # while later ranges are still being counted
                    # (as uint64 keys + uint32 counts: 12 bytes per row over PCIe; MC2_BENCH_ROWS16=1 asks for 16-byte rows)
                    if split_rows:
                        flat = out_rows.view(-1)
                        return engine.count_text_rows_split(host, args.k, args.c, flat.data_ptr(), flat[out_rows_cap:].data_ptr(), out_rows_cap) * 12
                    return engine.count_text_rows(host, args.k, args.c, out_rows.data_ptr(), out_rows_cap) * 16
                table, _ = count(host, args.k, args.c, args.s)
                n = table.rows
                if n > out_rows_cap:
                    raise RuntimeError("e2e: more rows than the pinned result buffers hold")
                flat = out_rows.view(-1)
                got = table.packed_to_host(flat.data_ptr(), flat[out_rows_cap:].data_ptr(), out_rows_cap)
                table.close()
                return got * 16

            d2h = step_e2e()                   # warm-up
            times = []
            for _ in range(e_steps):
                barrier()
                t1 = time.perf_counter()
                d2h = step_e2e()
                barrier()
                times.append(time.perf_counter() - t1)
            # a plain pinned-H2D ceiling measured in the same run (all ranks at once)
            dst = torch.empty(min(nbytes, 4 << 30), dtype=torch.uint8, device=device)
            barrier()
            t1 = time.perf_counter()
            for _ in range(3):
                dst.copy_(host[:dst.numel()], non_blocking=True)
            barrier()
            h2d_gbs = 3 * dst.numel() / (time.perf_counter() - t1) / 1e9
            del dst
            e2e = {"steps": e_steps, "times": times, "h2d": nbytes, "d2h": d2h, "h2d_ceiling_gbs": h2d_gbs,
                   "rows_format": "uint64 key + uint32 count, 12 B per row" if split_rows else "uint64 key + uint64 count, 16 B per row"}
            del host, out_rows

    # ---- max over ranks -------------------------------------------------------------------------------
    rows_total = rows
    if world > 1:
        vals = [elapsed, kernel_us, float(rows)] + (e2e["times"] + [e2e["h2d_ceiling_gbs"], float(e2e["d2h"])] if e2e else [])
        t = torch.tensor(vals, device=device, dtype=torch.float64)
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        elapsed = tmax[0].item()
        rows_total = int(tsum[2].item())
        if e2e:
            ns = len(e2e["times"])
            e2e["times"] = tmax[3:3 + ns].tolist()
            e2e["h2d_ceiling_gbs"] = tsum[3 + ns].item()                    # aggregate over the ranks
            e2e["d2h"] = int(tsum[4 + ns].item())
            e2e["h2d"] = nbytes * world
        if secondary:
            t2 = torch.tensor([secondary["elapsed"], float(secondary["rows"])], device=device, dtype=torch.float64)
            t2m = t2.clone()
            dist.all_reduce(t2m, op=dist.ReduceOp.MAX)
            dist.all_reduce(t2, op=dist.ReduceOp.SUM)
            secondary["elapsed"], secondary["rows"] = t2m[0].item(), int(t2[1].item())
        if cfg5:                                                      # (every rank takes part, also one whose pass failed)
            bad = "error" in cfg5
            t5 = torch.tensor([0.0 if bad else cfg5["seconds"], 0.0 if bad else float(cfg5["rows"]), 0.0 if bad else float(cfg5["proteins"]),
                               1.0 if bad else 0.0], device=device, dtype=torch.float64)
            t5m = t5.clone()
            dist.all_reduce(t5m, op=dist.ReduceOp.MAX)
            dist.all_reduce(t5, op=dist.ReduceOp.SUM)
            if t5m[3].item() > 0:
                cfg5 = cfg5 if bad else {"error": "another rank failed"}
            else:
                cfg5["seconds"], cfg5["rows"], cfg5["proteins"] = t5m[0].item(), int(t5[1].item()), int(t5[2].item())
        if timings:
            keys = sorted(timings)
            tt = torch.tensor([float(timings[k2]) for k2 in keys], device=device, dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            timings = dict(zip(keys, tt.tolist()))

    out = None
    if rank == 0:
        ms_per_step = elapsed / args.steps * 1e3
        value = world * bases_per_step * args.steps / elapsed
        peaks = {}
        try:
            peaks = json.load(open(ROOT / "MEASURED_PEAKS.json"))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"

        def roofline_of(prof, steps, rows_rank, step_s):
            """dominant kernel (largest share of the per-kernel event time of rank 0) + the whole path"""
            tot_us = sum(v["us"] for v in prof.values()) or 1.0
            top = max(prof.items(), key=lambda kv: kv[1]["us"]) if prof else ("none", {"launches": 1, "us": 1.0})
            alg_step = algorithmic_bytes_per_step(top[0], windows, nbytes, rows_rank)
            alg = alg_step * steps / max(1, top[1]["launches"]) if alg_step else None
            avg_us = top[1]["us"] / max(1, top[1]["launches"])
            achieved = alg / (avg_us * 1e-6) / 1e9 if alg else None
            # the whole path against SURVEY 8(d)'s sparse formula: text read once + every key written once and read
            # once (8 + 8 B per window) + 12 B per row of the emitted table
            pipeline_alg = nbytes + 16.0 * windows + 12.0 * rows_rank
            pipeline_gbs = pipeline_alg / step_s / 1e9
            return top, {"bound": "hbm", "kernel": top[0], "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": (achieved / peak) if achieved else None, "traffic": None, "traffic_unit": "bytes per launch",
                         "traffic_source": None, "peak_source": peak_src,
                         "kernel_share_of_step": top[1]["us"] / tot_us,
                         "algorithmic_bytes_per_launch": alg, "avg_launch_us": avg_us, "launches_per_step": top[1]["launches"] / steps,
                         "pipeline_bytes_per_base": pipeline_alg / bases_per_step,
                         "pipeline_achieved": pipeline_gbs, "pipeline_frac": pipeline_gbs / peak,
                         "kernel_time_share_of_wall": tot_us * 1e-6 / (step_s * steps)}

        top, roof = roofline_of(profile, args.steps, rows, elapsed / args.steps)
        try:                                        # DRAM bytes per launch of the dominant kernel, from the committed ncu capture
            tj = json.loads((ROOT / "profiles" / "traffic.json").read_text())
            if top[0] in tj["kernels"]:
                # measured DRAM bytes per key of the captured launch x the keys one launch of THIS run handles
                per_key = tj["kernels"][top[0]]["dram_bytes_per_key"]
                roof["traffic"] = per_key * windows * args.steps / max(1, top[1]["launches"])
                roof["traffic_source"] = tj["source"]
        except (OSError, ValueError, KeyError):
            pass
        phases = phases_from_profile(profile, args.steps)
        if timings:
            for key, val in timings.items():
                phases["host_" + key] = round(val / args.steps, 3) if key.endswith("_ms") else val / args.steps
            sent = timings.get("nvlink_bytes_sent", 0) / args.steps
            wait_s = timings.get("exchange_wait_ms", 0) / args.steps / 1e3
            phases["nvlink_bytes_per_rank_per_step"] = sent
            phases["nvlink_gbs_per_direction_vs_step"] = sent / (elapsed / args.steps) / 1e9
            phases["nvlink_exposed_wait_s"] = wait_s
            phases["nvlink_nominal_gbs"] = 900.0
        limiting = max(((k2, v) for k2, v in phases.items() if k2.endswith("_ms") and not k2.startswith("host_")), key=lambda kv: kv[1])[0]
        if timings and phases.get("host_exchange_wait_ms", 0) > phases[limiting]:
            limiting = "host_exchange_wait_ms"
        parallelism = (f"key-range partition x{world}: NCCL all-to-all of the keys before the filter, {args.groups_per_rank or 'auto'} rounds per rank"
                       if (world > 1 and args.s == 0) else f"chunk-sharded x{world}" if world > 1 else "single GPU")
        out = {
            "metric": METRIC,
            "value": value, "unit": "bases/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u64", "data": "synthetic",
            "config": {"workload": workload_name(args.k, args.c, args.s),
                       "reads_per_gpu": n_reads, "bases_per_gpu_per_step": bases_per_step, "text_bytes_per_gpu": nbytes,
                       "chunks_per_gpu": n_chunks, "surviving_rows": rows_total, "surviving_rows_rank0": rows, "parallelism": parallelism,
                       "numa_bound": numa_bound, "l2_policy": "input per step (>= 1 GB) is larger than L2; no flush needed"},
            "clocks": clocks,
            "gpu_launches": launches,
            "launches_per_step": launches / args.steps,
            "roofline": roof,
            "phases": phases, "limiting_phase": limiting,
            "kernels": {k2: {"launches": v["launches"], "ms": round(v["us"] / 1e3, 3)} for k2, v in
                        sorted(profile.items(), key=lambda kv: -kv[1]["us"])[:14]},
            "check": check,
            "step_ms_rank0": step_ms, "warmup_ms_rank0": warm_ms, "overflow_buckets_rank0": overflow_buckets,
        }
        if secondary:
            _, roof2 = roofline_of(secondary["profile"], secondary["steps"], secondary["rows"] / world, secondary["elapsed"] / secondary["steps"])
            out["secondary"] = {"workload": workload_name(args.k, secondary["c"], secondary["s"]),
                                "value": world * bases_per_step * secondary["steps"] / secondary["elapsed"], "unit": "bases/s",
                                "ms_per_step": secondary["elapsed"] / secondary["steps"] * 1e3, "steps": secondary["steps"],
                                "chunks_per_gpu": secondary["chunks"], "surviving_rows": secondary["rows"],
                                "roofline": {k2: roof2[k2] for k2 in ("kernel", "achieved", "frac", "kernel_share_of_step", "pipeline_frac", "pipeline_achieved",
                                                                       "kernel_time_share_of_wall")},
                                "phases": phases_from_profile(secondary["profile"], secondary["steps"]),
                                "kernels": {k2: {"launches": v["launches"], "ms": round(v["us"] / 1e3, 3)} for k2, v in
                                            sorted(secondary["profile"].items(), key=lambda kv: -kv[1]["us"])[:12]}}
        if cfg5:
            if "error" not in cfg5:
                cfg5 = {"workload": "cfg5: 64 synthetic proteomes x 5000 proteins (S5), k=5 -c 10 + pI/MW metrics, whole samples per rank (LPT), batched",
                        "samples": 64, "proteins": cfg5["proteins"], "residues": cfg5["residues"], "rows": cfg5["rows"],
                        "ms": cfg5["seconds"] * 1e3, "samples_per_s": 64 / cfg5["seconds"], "residues_per_s": cfg5["residues"] / cfg5["seconds"]}
            out.setdefault("secondary", {})["cfg5"] = cfg5
        if e2e:
            ts = sorted(e2e["times"])
            med = ts[len(ts) // 2]
            total_bases = world * bases_per_step
            out["e2e"] = {"value": total_bases * len(ts) / sum(ts), "unit": "bases/s",
                          "h2d_bytes_per_step": e2e["h2d"], "d2h_bytes_per_step": e2e["d2h"], "steps": len(ts),
                          "min_s": ts[0], "median_s": med, "max_s": ts[-1], "value_at_median": total_bases / med,
                          "h2d_ceiling_gbs": e2e["h2d_ceiling_gbs"], "rows_format": e2e["rows_format"],
                          "pcie_floor_s": (e2e["h2d"] + e2e["d2h"]) / world / (e2e["h2d_ceiling_gbs"] / world * 1e9),
                          "note": "pcie_floor_s = (H2D + D2H bytes of one rank) / measured pinned-H2D rate of one rank when all ranks copy at once"}

    # ---- CPU baseline beside it (rank 0, N = 1 only) ---------------------------------------------------
    if rank == 0 and world == 1 and not args.no_cpu:
        cores = host_cores()
        n_cpu = args.cpu_reads or min(n_reads, cpu_reads_auto(cores, args.k))
        sample = text[: n_cpu * REC_BYTES].cpu().numpy().tobytes()
        with tempfile.TemporaryDirectory() as tmp:
            files = cpu_sample_files(sample, 2 * cores, tmp)
            with multiprocessing.get_context("fork").Pool(cores) as pool:
                secs = cpu_run(files, args.k, args.c, pool, os.path.join(tmp, "out.tsv"))
        out["cpu_baseline"] = {"value": n_cpu * READ_LEN / secs, "unit": "bases/s", "cores": cores, "kind": "port",
                               "sample": f"first {n_cpu} reads of the same shard ({n_cpu * READ_LEN / 1e6:.1f} Mbp) as "
                                         f"{len(files)} chunk files, one oracle find_kmers task per file on a "
                                         f"{cores}-process pool (-c {args.c}), serial merge + sorted TSV; {secs:.1f} s"}
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()
    return 0


def run_check(engine, genomes, args, rank, world, dist, device):
    """Known-answer gate on a CPU-sized prefix of the SAME reads with the SAME flags: the engine's table (through the
    path the timed steps take) against the oracle's, as TSV digests.  N > 1 additionally runs the key-exchange path on
    the prefix split over the ranks."""
    import numpy as np
    from oracle import mercat2_oracle as orc
    n_check = min(args.reads_per_gpu, int(os.environ.get("MC2_BENCH_CHECK_READS", 40_000)))
    pre = make_reads_text(device, genomes, n_check, 0)          # the job's first reads: the same text on every rank
    res = {"reads": n_check, "flags": f"-k {args.k} -c {args.c} -s 0"}
    if rank == 0:
        want = orc.find_kmers_text(pre.cpu().numpy().tobytes().decode(), args.k, args.c)
        want_tsv = orc.tsv_bytes("sample", want) if want else b""
        t = engine.count_text(pre, args.k, args.c)
        got_tsv = t.tsv_bytes("sample") if t.rows else b""
        t.close()
        res.update({"oracle_md5": hashlib.md5(want_tsv).hexdigest(), "engine_md5": hashlib.md5(got_tsv).hexdigest(),
                    "rows": len(want), "ok": got_tsv == want_tsv})
        # the same prefix forced through the level-0 (very large chunk) machinery the headline uses
        engine.set_option("hash_bucket_keys", 16)
        try:
            t = engine.count_text(pre, args.k, args.c)
            big_tsv = t.tsv_bytes("sample") if t.rows else b""
            t.close()
        finally:
            engine.set_option("hash_bucket_keys", 3500)
        res["level0_ok"] = big_tsv == want_tsv
        res["ok"] = res["ok"] and res["level0_ok"]
    if world > 1:
        from mercat2_b200 import distributed as mcd
        # every rank takes its slice of the prefix (cut at read boundaries); rank parts concatenated must equal the oracle
        per = n_check // world
        a = rank * per * REC_BYTES
        b = n_check * REC_BYTES if rank == world - 1 else (rank + 1) * per * REC_BYTES
        part = mcd.count_piece_position_sharded(engine, pre[a:b], args.k, args.c, dist, device, groups_per_rank=3)
        body = part.tsv_body() if part.rows else b""
        part.close()
        bodies = [None] * world
        dist.all_gather_object(bodies, body)
        if rank == 0:
            res["sharded_ok"] = (b"k-mer\tsample_Count\n" + b"".join(bodies) if any(bodies) else b"") == want_tsv
            res["ok"] = res["ok"] and res["sharded_ok"]
    if rank == 0 and not res["ok"]:
        print(json.dumps({"error": "bench check failed: engine table differs from the oracle", "check": res}))
        raise SystemExit(3)
    return res if rank == 0 else None


def reference_arm(args, world):
    """CPU reference arm: the oracle port (kind 'port': /root/reference is absent on the GPU box and the
    reference is pure Python, so the port IS its algorithm and data structures) on all host cores."""
    import numpy as np
    cores = host_cores()
    n_cpu = args.cpu_reads or cpu_reads_auto(cores, args.k)
    # same read definition as the GPU arm when a GPU is present; otherwise an equivalent numpy generator
    try:
        import torch
        if torch.cuda.is_available():
            dev = torch.device("cuda", 0)
            genomes = make_genomes(dev)
            sample = make_reads_text(dev, genomes, n_cpu, 0).cpu().numpy().tobytes()
            del genomes
        else:
            raise RuntimeError
    except Exception:
        rng = np.random.default_rng(SEED)
        genome = rng.integers(0, 4, 20_000_000, dtype=np.uint8)
        lut = np.frombuffer(b"ACGT", dtype=np.uint8)
        starts = rng.integers(0, genome.size - READ_LEN, n_cpu)
        recs = [b">r%010d\n" % i + lut[genome[s:s + READ_LEN]].tobytes() + b"\n" for i, s in enumerate(starts)]
        sample = b"".join(recs)
    times = []
    with tempfile.TemporaryDirectory() as tmp:
        files = cpu_sample_files(sample, 2 * cores, tmp)
        with multiprocessing.get_context("fork").Pool(cores) as pool:
            for i in range(args.warmup + args.steps):
                secs = cpu_run(files, args.k, args.c, pool, os.path.join(tmp, "out.tsv"))
                if i >= args.warmup:
                    times.append(secs)
    elapsed = sum(times)
    value = n_cpu * READ_LEN * len(times) / elapsed
    out = {
        "impl": "reference",
        "metric": METRIC,
        "value": value, "unit": "bases/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": elapsed / len(times) * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64", "data": "synthetic",
        "config": {"workload": workload_name(args.k, args.c, args.s), "sample_reads": n_cpu,
                   "note": "the reference runs one find_kmers task per chunk FILE (bin/mercat2.py:120); with -s 0 a single file "
                           "would be ONE task on ONE core, so the prefix is given to it as 2 x cores chunk files to let it use every core"},
        "cpu_baseline": {"value": value, "unit": "bases/s", "cores": cores, "kind": "port",
                         "sample": f"{n_cpu} reads ({n_cpu * READ_LEN / 1e6:.1f} Mbp) per step as {2 * cores} chunk files, "
                                   f"{cores}-process pool, -c {args.c}, serial merge + sorted TSV"},
        "e2e": {"value": value, "unit": "bases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out))
    return 0


if __name__ == "__main__":
    sys.exit(main())
