#!/usr/bin/env python
"""Golden vectors for merge_tsv / merge_tsv_T: runs the reference's OWN functions
(lib/mercat2_report.py:98-194) on the per-sample TSVs already committed under tests/golden/expected/ and stores the
merged tables.  Run in the build container (needs /root/reference); dominate (HTML report library, absent here) is
stubbed -- merge_tsv does not touch it."""
import gzip
import importlib
import shutil
import sys
import tempfile
import types
from pathlib import Path

REF = Path("/root/reference")
ROOT = Path(__file__).resolve().parents[1]
GOLD = ROOT / "tests" / "golden"


def import_report():
    pkg = types.ModuleType("mercat2_lib")
    pkg.__path__ = [str(REF / "lib")]
    sys.modules["mercat2_lib"] = pkg
    dom = types.ModuleType("dominate")
    tags = types.ModuleType("dominate.tags")
    util = types.ModuleType("dominate.util")
    util.raw = lambda x: x
    tags.__all__ = []
    sys.modules.update({"dominate": dom, "dominate.tags": tags, "dominate.util": util})
    import pkg_resources                       # the module reads its stylesheet / logo at import time
    pkg_resources.resource_stream = lambda _pkg, res: open(REF / "lib" / res, "rb")
    return importlib.import_module("mercat2_lib.mercat2_report")


def main():
    rep = import_report()
    exp = GOLD / "expected"
    with tempfile.TemporaryDirectory() as tmp:
        tmp = Path(tmp)
        for label, pattern in (("nucleotide_k3", "*_k3_c10.tsv"), ("protein_k3", "*_pro_k3_c10.tsv.gz")):
            tsv_list = {}
            for f in sorted(exp.glob(pattern)):
                if label == "nucleotide_k3" and "_pro_" in f.name:
                    continue
                name = f.name.split("_k3_")[0]
                dst = tmp / f"{name}_counts.tsv"
                if f.suffix == ".gz":
                    dst.write_bytes(gzip.open(f, "rb").read())
                else:
                    shutil.copy(f, dst)
                tsv_list[name] = str(dst)
            out = tmp / f"combined_{label}.tsv"
            rep.merge_tsv(tsv_list, out)
            with gzip.open(GOLD / f"merged_{label}.tsv.gz", "wb", 9) as w:
                w.write(out.read_bytes())
            out_t = tmp / f"combined_{label}_T.tsv"
            rep.merge_tsv_T(tsv_list, out_t)
            with gzip.open(GOLD / f"merged_{label}_T.tsv.gz", "wb", 9) as w:
                w.write(out_t.read_bytes())
            print(label, sorted(tsv_list), out.stat().st_size, out_t.stat().st_size)


if __name__ == "__main__":
    main()
