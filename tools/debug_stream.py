import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch, numpy as np
import mercat2_b200, bench
eng = mercat2_b200.Engine(0)
dev = torch.device("cuda", 0)
genomes = bench.make_genomes(dev, 0.002)
text = bench.make_reads_text(dev, genomes, 500_000, 0)
print("numel", text.numel())
for chunk in (16 << 20, 8 << 20, 20_000_000):
    print("device chunk_offsets", chunk, eng.chunk_offsets(text, chunk))
    t, offs = eng.count_sample(text, 25, 3, chunk); print("  count_sample dev", offs, t.rows)
    host = text.cpu().numpy()
    t, offs = eng.count_sample(host, 25, 3, chunk); print("  count_sample host", offs, t.rows)
