import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np, torch
import mercat2_b200, bench
eng = mercat2_b200.Engine(0)
dev = torch.device("cuda", 0)
g = bench.make_genomes(dev, 0.01)
text = bench.make_reads_text(dev, g, 1400000, 0)
print("bytes", text.numel(), flush=True)
t, offs = eng.count_sample(text, 31, 10, 100 << 20)
print("chunks", len(offs), offs, "rows", t.rows)
host = text[:164*20].cpu().numpy().tobytes()
print(host[:400])
t = eng.count_text(host, 31, 2)
print(t.rows)
