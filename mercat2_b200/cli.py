"""mercat2.py-compatible command line for the k-mer counting path (thin re-host of bin/mercat2.py:37-81, :186-503
without Ray, figures, diversity or ORF calling).

Same flags for the path: -i/-f, -k, -c [10], -s [100], -n, -o, -replace, -skipclean, -toupper.  Same output tree
for it: <out>/clean/*, <out>/tsv_<type>/<base>_counts.tsv, <out>/report/metrics-protein.tsv.  Flags of subsystems
that are out of scope (-prod, -fgs, -pca, -lowmem) are accepted and reported as skipped.  Under torchrun
(WORLD_SIZE > 1, one process per GPU, NCCL) whole samples are dealt to the ranks by size (longest processing time first:
SURVEY 8e grain 1, no collective on the data path); each rank writes the TSVs of its samples and rank 0 writes the ONE
report/metrics-protein.tsv from the rows the ranks hand in, in the reference's sample order.
"""
from __future__ import annotations

import argparse
import os
import shutil
import sys
import timeit
from pathlib import Path

from . import __version__, mercat2_fasta, mercat2_metrics, pipeline
from . import distributed as mcd

FILE_EXT_FASTQ = [".fq", ".fastq", ".fq.gz", ".fastq.gz"]
FILE_EXT_NUCLEOTIDE = [".fasta", ".fa", ".fna", ".ffn", ".fasta.gz", ".fa.gz", ".fna.gz", ".ffn.gz"]
FILE_EXT_PROTEIN = [".faa", ".faa.gz"]


def parseargs(argv=None):
    parser = argparse.ArgumentParser(prog="mercat2.py", formatter_class=argparse.RawDescriptionHelpFormatter)
    parser.add_argument("-i", required=False, default=list(), help="path to input file", nargs="+")
    parser.add_argument("-f", type=str, required=False, help="path to folder containing input files")
    parser.add_argument("-k", type=int, required=True, help="kmer length")
    parser.add_argument("-n", type=int, default=os.cpu_count(), help="no of cores [auto detect] (host pre-processing only)")
    parser.add_argument("-c", type=int, default=10, help="minimum kmer count [10]")
    parser.add_argument("-prod", action="store_true", help="(out of scope here) run Prodigal on fasta files")
    parser.add_argument("-fgs", action="store_true", help="(out of scope here) run FragGeneScanRS on fasta files")
    parser.add_argument("-s", type=int, default=100, required=False, help="Split into x MB files. [100]")
    parser.add_argument("-o", type=str, default="mercat_results", required=False, help="Output folder")
    parser.add_argument("-replace", action="store_true", help="Replace existing output directory [False]")
    parser.add_argument("-lowmem", type=str, default=None, help=argparse.SUPPRESS)
    parser.add_argument("-skipclean", action="store_true", help="skip trimming of fastq files")
    parser.add_argument("-toupper", action="store_true", help="convert all input sequences to uppercase")
    parser.add_argument("-pca", action="store_true", help="(out of scope here) PCA plot")
    parser.add_argument("--version", "-v", action="version", version=f"MerCat2 (mercat2_b200 {__version__})")
    args = parser.parse_args(argv)
    if not args.i and not args.f:
        parser.error("Please provide either an input file (-i) or an input folder (-f)")
    for filename in args.i:
        if not os.path.isfile(filename):
            parser.error(f"file '{filename}' is not valid.\n")
    if args.f and not os.path.isdir(args.f):
        parser.error(f"folder {args.f} is not valid.\n")
    return args, parser


def classify(path: Path):
    """(extension class, basename) like bin/mercat2.py:264-283."""
    suffixes = Path(path).suffixes
    f_ext = ""
    for i in reversed(range(len(suffixes))):
        if "".join(suffixes[i:]) in FILE_EXT_FASTQ + FILE_EXT_NUCLEOTIDE + FILE_EXT_PROTEIN:
            f_ext = "".join(suffixes[i:])
    base = Path(path).name.removesuffix(f_ext)
    kind = "fastq" if f_ext in FILE_EXT_FASTQ else "nucleotide" if f_ext in FILE_EXT_NUCLEOTIDE else \
        "protein" if f_ext in FILE_EXT_PROTEIN else None
    return kind, base


def mercat_main(argv=None):
    args, parser = parseargs(argv)
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    out = Path(args.o)
    if rank == 0:
        if out.exists():
            if args.replace:
                shutil.rmtree(out)
            else:
                parser.error(f"Output folder exists, please specify another folder or use the flag '-replace' "
                             f"to override the files. '{out}'")
        out.mkdir(0o777, True, True)
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if torch.cuda.is_available():
            local = int(os.environ.get("LOCAL_RANK", "0"))
            torch.cuda.set_device(local)
            mcd.init_nccl(torch.device("cuda", local), dist)
        else:
            dist.init_process_group("gloo")                      # (CPU-only hosts: the CLI's own tests)
        dist.barrier()
    if rank == 0:
        print(f"\nStarting MerCat2 (mercat2_b200 {__version__}) with k-mer {args.k} on {world} GPU(s)\n")
        for flag in ("prod", "fgs", "pca"):
            if getattr(args, flag):
                print(f"NOTE: -{flag} belongs to a subsystem outside this engine's scope; skipped")
    inputs = list(args.i)
    if args.f:
        folder = os.path.abspath(os.path.expanduser(args.f))
        inputs += [str(Path(folder, name)) for name in sorted(os.listdir(folder))
                   if classify(Path(folder, name))[0] is not None]
    # whole samples per rank, balanced by on-disk size (longest processing time first)
    inputs = sorted(set(inputs))
    mine = set(mcd.shard_lpt([os.stat(f).st_size for f in inputs], world)[rank])
    todo = [f for i, f in enumerate(inputs) if i in mine]
    cleanpath = os.path.join(out, "clean")
    samples = {"nucleotide": {}, "protein": {}}
    start = timeit.default_timer()
    for filename in todo:
        kind, base = classify(Path(filename))
        path = Path(filename).expanduser().absolute()
        if kind == "fastq":
            samples["nucleotide"][base] = mercat2_fasta.fq2fa(str(path), cleanpath, base)
        elif kind == "nucleotide":
            if args.skipclean:
                samples["nucleotide"][base] = str(path)
            else:
                samples["nucleotide"][base] = str(mercat2_fasta.removeN(path, cleanpath, args.toupper)[0])
        elif kind == "protein":
            samples["protein"][base] = str(path)
    print(f"Time to load {len(samples['nucleotide']) + len(samples['protein'])} files: "
          f"{round(timeit.default_timer() - start, 2)} seconds")
    os.makedirs(os.path.join(out, "report"), exist_ok=True)
    for sample_type, label in (("nucleotide", "Nucleotides"), ("protein", "Proteins")):
        if not samples[sample_type]:
            continue
        print(f"Processing {label}")
        out_tsv = os.path.join(out, f"tsv_{sample_type}")
        os.makedirs(out_tsv, exist_ok=True)
        start = timeit.default_timer()
        pipeline.run_samples(samples[sample_type], out_tsv, args.k, args.c, chunk_size_mb=args.s)
        print(f"Time to count {args.k}-mers: {round(timeit.default_timer() - start, 2)} seconds")
    # report/metrics-protein.tsv: every rank computes the rows of its samples, rank 0 writes the one file
    rows = {}
    for base, file in samples["protein"].items():
        rows[base] = list(pipeline.sample_metric_rows(file, args.s))
    gathered = [rows]
    if world > 1:
        gathered = [None] * world if rank == 0 else None
        dist.gather_object(rows, gathered, dst=0)
    if rank == 0:
        merged = {}
        for part in gathered:
            merged.update(part)
        order = [classify(Path(f))[1] for f in inputs if classify(Path(f))[0] == "protein"]
        if merged:
            with open(os.path.join(out, "report", "metrics-protein.tsv"), "w") as writer:
                print("Sample", "seq_name", "length", "PI", "MW", "Hydro", sep="\t", file=writer)
                for base in order:
                    for header, name, length, pi, mw, hydro in merged.get(base, ()):
                        print(header, name, length, "" if pi is None else pi, mw, hydro, sep="\t", file=writer)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(mercat_main())
