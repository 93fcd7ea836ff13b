import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import mercat2_b200
from test_gpu_parity import synth_reads
eng = mercat2_b200.Engine(0)
eng.set_option("force_path", 3)
text = synth_reads(6000, 150, seed=238, n_rate=0.002, lower_rate=0.01)
for k, c in ((12, 1), (12, 2), (12, 3), (9, 2)):
    print("== k", k, "c", c, flush=True)
    t = eng.count_text(text, k, c)
    print("rows", t.rows, flush=True)
