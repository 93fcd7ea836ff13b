#!/usr/bin/env python
"""bench.py -- input bases/s of the k-mer counting hot path on B200.

Workload (BASELINE.json config 4, the one the metric is quoted on): synthetic 150-bp metagenome reads
(200 genomes x 5 Mbp, 0.1 % substitutions, FASTA framing '>r%010d\\n' + 150 bases + '\\n' = 164 B per
read), nucleotide k=31, the reference's default flags -c 10 -s 100.  The 100 Gbp job is sharded over 8
GPUs: each rank owns reads_per_gpu reads (default 12.5 Gbp, so N=8 is exactly the named 100 Gbp);
scaling is weak (fixed work per GPU).  A step is one pass of the hot path (chunk -> parse -> extract ->
count -> per-chunk filter -> sample table) over the rank's shard.

  value  : whole-job bases/s with the FASTA text already resident in HBM
  e2e    : the same through the C ABI with the text in pinned HOST memory (H2D inside the timed
           region) and the result table read back (D2H)
  roofline / cpu_baseline: see DESIGN.md

`--impl reference` times the CPU reference arm (the oracle port of the reference's Python counter on
all host cores; /root/reference does not exist on the GPU box) on a bounded sample of the same reads.
"""
from __future__ import annotations

import argparse
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


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--reads-per-gpu", type=int, default=int(os.environ.get("MC2_BENCH_READS", 83_333_334)))
    ap.add_argument("-k", type=int, default=31)
    ap.add_argument("-c", type=int, default=10)
    ap.add_argument("-s", type=int, default=100, help="chunk size in MB (reference default 100)")
    ap.add_argument("--cpu-reads", type=int, default=0, help="reads in the CPU sample (0 = auto)")
    ap.add_argument("--genome-scale", type=float, default=1.0, help="fraction of the 1 Gbp metagenome to synthesise (profiling runs)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
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
    """nvidia-smi samples every 200 ms during the timed region."""
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
            self.lines.append(line.strip())

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
        sm, mx, reasons = [], 0, set()
        for line in self.lines:
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


# ---------------------------------------------------------------------------------------------------
def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank != 0:
            return 0
        return reference_arm(args, world)

    import torch
    import torch.distributed as dist
    import mercat2_b200

    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    numa_bound = mercat2_b200.bind_to_gpu_numa(local) if world > 1 else False     # pinned shards next to their GPU
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)

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
    chunk_bytes = args.s * 1024 * 1024 if (args.s > 0 and nbytes >= args.s * 1024 * 1024) else 0

    engine = mercat2_b200.Engine(local)

    def merged(table):
        """N > 1: the ranks hold pieces of ONE sample -- sum their filtered tables on the devices (key-range
        all-to-all over NCCL); every rank keeps the rows of its own key range."""
        if world == 1:
            return table
        from mercat2_b200 import distributed as mcd
        part = mcd.merge_table_device(engine, table, dist, device)
        table.close()
        return part

    def step_resident():
        table, offsets = engine.count_sample(text, args.k, args.c, chunk_bytes)
        table = merged(table)
        rows = table.rows
        table.close()
        return rows, len(offsets)

    for _ in range(args.warmup):
        rows, n_chunks = step_resident()
    launches0 = engine.stat("launches")
    # CUDA events on the engine's stream around the three long kernels of the path (scatter 1, scatter 2, count) -- the
    # dominant kernel is one of them; MC2_BENCH_PROFILE=2 times every launch (costs ~6 % of the step)
    engine.set_option("profile", int(os.environ.get("MC2_BENCH_PROFILE", "3")))
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    t0 = time.perf_counter()
    dev_us = 0.0
    for _ in range(args.steps):
        rows, n_chunks = step_resident()
        dev_us += engine.stat("device_us")
    barrier()
    elapsed = time.perf_counter() - t0
    clocks = sampler.stop()
    launches = engine.stat("launches") - launches0
    profile = engine.profile()
    engine.set_option("profile", 0)

    # ---- e2e: host buffer in, table out, through the same C-ABI call -------------------------------
    e2e = None
    if not args.no_e2e:
        import psutil
        # every rank pins its own shard; the decision is taken once (rank 0) so that all ranks agree
        fits = [psutil.virtual_memory().available > int(1.25 * nbytes * world) + (16 << 30)]
        if world > 1:
            dist.broadcast_object_list(fits, src=0)
        if fits[0]:
            host = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
            host.copy_(text)
            torch.cuda.synchronize()
            e_steps = max(1, min(args.steps, 3))

            def step_e2e():
                table, _ = engine.count_sample(host, args.k, args.c, chunk_bytes)
                table = merged(table)
                kmers, counts = table.arrays()
                table.close()
                return kmers.nbytes + counts.nbytes

            d2h = step_e2e()                   # warm-up
            barrier()
            t1 = time.perf_counter()
            for _ in range(e_steps):
                d2h = step_e2e()
            barrier()
            e_elapsed = time.perf_counter() - t1
            e2e = {"steps": e_steps, "elapsed": e_elapsed, "h2d": nbytes, "d2h": d2h}
            del host

    # ---- max over ranks -------------------------------------------------------------------------------
    if world > 1:
        t = torch.tensor([elapsed, e2e["elapsed"] if e2e else 0.0, dev_us], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed, e_el, dev_us = t.tolist()
        if e2e:
            e2e["elapsed"] = e_el

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
        # dominant kernel: largest share of the per-kernel event time
        tot_us = dev_us if dev_us > 0 else (sum(v["us"] for v in profile.values()) or 1.0)   # device time of the timed steps
        top = max(profile.items(), key=lambda kv: kv[1]["us"]) if profile else ("none", {"launches": 1, "us": 1.0})
        windows_per_chunk = (bases_per_step * (READ_LEN - args.k + 1) / READ_LEN) / max(1, n_chunks)
        alg = algorithmic_bytes(top[0], windows_per_chunk, nbytes / max(1, n_chunks), args.k)
        avg_us = top[1]["us"] / max(1, top[1]["launches"])
        achieved = alg / (avg_us * 1e-6) / 1e9 if alg else None
        # the whole path against SURVEY 8(d)'s sparse formula: text read once + every key written once and read
        # once (8 + 8 B per window) + 12 B per row of the emitted table (rows = survivors of the -c filter)
        w_total = bases_per_step * (READ_LEN - args.k + 1) / READ_LEN
        pipeline_alg = nbytes + 16.0 * w_total + 12.0 * rows
        pipeline_gbs = pipeline_alg * args.steps / (dev_us * 1e-6) / 1e9 if dev_us else None
        traffic, traffic_src = None, None
        try:                                        # DRAM bytes per launch of the dominant kernel, from the committed ncu capture
            tj = json.loads((Path(__file__).parent / "profiles" / "traffic.json").read_text())
            if top[0] in tj["kernels"] and n_chunks >= 2:
                traffic, traffic_src = tj["kernels"][top[0]]["dram_bytes_per_launch"], tj["source"]
        except (OSError, ValueError, KeyError):
            pass
        out = {
            "metric": "input bases/sec (k-mers counted/sec) per GPU and 8xB200; % of HBM roofline",
            "value": value, "unit": "bases/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u64", "data": "synthetic",
            "config": {"workload": f"cfg4 shard: synthetic 150-bp metagenome reads, nucleotide k={args.k} -c {args.c} -s {args.s}",
                       "reads_per_gpu": n_reads, "bases_per_gpu_per_step": bases_per_step, "text_bytes_per_gpu": nbytes,
                       "chunks_per_gpu": n_chunks, "surviving_rows": rows, "parallelism": f"chunk-sharded x{world}" + (" + NCCL key-range all-to-all of the filtered tables" if world > 1 else ""),
                       "numa_bound": numa_bound, "l2_policy": "input per step (>= 1 GB) is larger than L2; no flush needed"},
            "clocks": clocks,
            "gpu_launches": launches,
            "device_ms_per_step": dev_us / args.steps / 1e3,
            "roofline": {"bound": "hbm", "kernel": top[0], "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": (achieved / peak) if achieved else None, "traffic": traffic, "traffic_unit": "bytes per launch",
                         "traffic_source": traffic_src, "peak_source": peak_src,
                         "kernel_share_of_step": top[1]["us"] / tot_us,
                         "algorithmic_bytes_per_launch": alg, "avg_launch_us": avg_us,
                         "pipeline_bytes_per_base": pipeline_alg / bases_per_step,
                         "pipeline_achieved": pipeline_gbs, "pipeline_frac": (pipeline_gbs / peak) if pipeline_gbs else None},
            "kernels": {k2: {"launches": v["launches"], "ms": round(v["us"] / 1e3, 3)} for k2, v in
                        sorted(profile.items(), key=lambda kv: -kv[1]["us"])[:12]},
        }
        if e2e:
            out["e2e"] = {"value": world * bases_per_step * e2e["steps"] / e2e["elapsed"], "unit": "bases/s",
                          "h2d_bytes_per_step": e2e["h2d"], "d2h_bytes_per_step": e2e["d2h"]}

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
                                         f"{cores}-process pool, serial merge + sorted TSV; {secs:.1f} s"}
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()
    return 0


def algorithmic_bytes(kernel, windows, text_bytes, k):
    """Algorithmic bytes of ONE launch of the named kernel for one chunk: what the
    kernel must read and write once, not what it happens to move."""
    name = kernel.split("<")[0]
    symbols = text_bytes * (READ_LEN + 1) / REC_BYTES            # bases + one separator per read
    packed = symbols * 0.375                                     # 2-bit codes + 1 validity bit per symbol
    table = {
        "fn_parse_kernel": text_bytes + packed if "true" in kernel else text_bytes,
        "fn_hist_kernel": packed,
        "fn_scatter1_kernel": packed + 8.0 * windows,            # every key written once
        "hc_scatter1_kernel": symbols + 8.0 * windows,
        "hc_scatter2_kernel": 16.0 * windows,                    # every key read once and written once
        "hc_count2_kernel": 8.0 * windows,                       # every key read once
        "hc_count_kernel": 8.0 * windows,
        "rs_scatter_kernel": 16.0 * windows,
        "rs_hist_kernel": 8.0 * windows,
        "extract_keys_kernel": symbols + 8.0 * windows,
        "chunk_candidates_kernel": text_bytes,
    }
    if name in table:
        return table[name]
    if name.startswith("parse_"):
        return text_bytes
    if name.startswith("dense_"):
        return symbols
    return None


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
        "metric": "input bases/sec (k-mers counted/sec) per GPU and 8xB200; % of HBM roofline",
        "value": value, "unit": "bases/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": elapsed / len(times) * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64", "data": "synthetic",
        "config": {"workload": f"cfg4 shard: synthetic 150-bp metagenome reads, nucleotide k={args.k} -c {args.c} -s {args.s}",
                   "sample_reads": n_cpu},
        "cpu_baseline": {"value": value, "unit": "bases/s", "cores": cores, "kind": "port",
                         "sample": f"{n_cpu} reads ({n_cpu * READ_LEN / 1e6:.1f} Mbp) per step as {2 * cores} chunk files, "
                                   f"{cores}-process pool, serial merge + sorted TSV"},
        "e2e": {"value": value, "unit": "bases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out))
    return 0


if __name__ == "__main__":
    sys.exit(main())
