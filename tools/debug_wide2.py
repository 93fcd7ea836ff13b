import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
import mercat2_b200
from test_gpu_parity import synth_reads
eng = mercat2_b200.Engine(0)
eng.set_option("force_path", 3)
text = synth_reads(6000, 150, seed=238, n_rate=0.002, lower_rate=0.01)
for k in (9, 12, 33):
    t1 = eng.count_text(text, k, 1)
    km1, c1 = t1.arrays()
    heads = np.concatenate([[0], np.cumsum(c1)[:-1]]).astype(np.int64)
    keys1 = [bytes(r) for r in km1]
    for c in (2, 3):
        for rep in range(2):
            t2 = eng.count_text(text, k, c)
            km2, c2 = t2.arrays()
            got = {bytes(r): int(n) for r, n in zip(km2, c2)}
            exp_idx = [i for i in range(len(c1)) if c1[i] >= c]
            missing = [i for i in exp_idx if keys1[i] not in got]
            wrong = [i for i in exp_idx if keys1[i] in got and got[keys1[i]] != c1[i]]
            print(f"k={k} c={c} rep={rep} rows1={len(c1)} exp={len(exp_idx)} got={len(got)} missing={len(missing)} wrong={len(wrong)}")
            for i in missing[:12]:
                print("   miss idx", i, keys1[i].decode(), "count", int(c1[i]), "head", int(heads[i]), "head%256", int(heads[i]) % 256,
                      "end%256", int(heads[i] + c1[i] - 1) % 256, "prev", keys1[i-1].decode(), int(c1[i-1]), "next", keys1[i+1].decode(), int(c1[i+1]))
