import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
import mercat2_b200
from oracle import mercat2_oracle as orc
from test_gpu_parity import synth_reads, diff_msg
eng = mercat2_b200.Engine(0)
k, c = 13, 2
text = synth_reads(6000, 150, seed=k * 7 + c, n_rate=0.002, lower_rate=0.01)
want = orc.find_kmers_text(text.decode(), k, c)
for fast in (0, 1):
    for bk in (64, 1000, 3500, 7000, 7100, 20000, 1000000):
        for rep in range(2):
            eng.set_option("force_path", 2); eng.set_option("sparse_algo", 2); eng.set_option("hash_bucket_keys", bk); eng.set_option("fast_nt", fast)
            got = eng.count_text(text, k, c).to_dict()
            print(f"fast={fast} bucket_keys={bk} rep={rep} ok={got == want}", "" if got == want else diff_msg(got, want)[:200], flush=True)
