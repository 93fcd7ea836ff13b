"""Debug matrix for the wide path: forced path=3 on synthetic reads for many (k, c)."""
import sys
sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import numpy as np
import mercat2_b200
from oracle import mercat2_oracle as orc
from test_gpu_parity import synth_reads, diff_msg

eng = mercat2_b200.Engine(0)
for nreads in (300, 6000):
    text = synth_reads(nreads, 150, seed=238, n_rate=0.002, lower_rate=0.01)
    for k in (7, 8, 9, 12, 16, 17, 24, 25, 31, 32, 33, 40, 41):
        for c in (1, 2, 3):
            want = orc.find_kmers_text(text.decode(), k, c)
            eng.set_option("force_path", 3)
            got_t = eng.count_text(text, k, c)
            kmers, counts = got_t.arrays()
            got = got_t.to_dict()
            dup = len(counts) - len(got)
            srt = bool((np.lexsort(kmers.T[::-1]) == np.arange(len(counts))).all()) if len(counts) else True
            ok = got == want
            print(f"nreads={nreads} k={k} c={c} ok={ok} rows={len(counts)} want={len(want)} duprows={dup} sorted={srt}",
                  "" if ok else diff_msg(got, want)[:300], flush=True)
