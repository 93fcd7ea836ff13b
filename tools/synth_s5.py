"""S5 -- the synthetic protein samples of BASELINE config 5: sample j = `proteins` proteins, lengths
30 + geometric(mean 300) capped at 5 000, residues drawn from the empirical 20-letter frequencies of DJ_pro
(tests/golden/data/faa_gz/DJ_pro.faa.gz: 1 544 198 residues), one trailing '*', 60-column lines, seed 1000 + j.
Test / bench infrastructure (numpy), not product code."""
from __future__ import annotations

import numpy as np

RESIDUES = b"ACDEFGHIKLMNPQRSTVWY"
# measured on DJ_pro (python: Counter over all sequence lines, '*' removed), rounded to 1e-5
FREQ = np.array([0.116, 0.01039, 0.05318, 0.06269, 0.03535, 0.08507, 0.02322, 0.04396, 0.0282, 0.12242,
                 0.02045, 0.02535, 0.05217, 0.03879, 0.07949, 0.0533, 0.04368, 0.06735, 0.01446, 0.02449])


def sample_text(j: int, proteins: int = 5000) -> bytes:
    """FASTA text of sample j."""
    rng = np.random.default_rng(1000 + j)
    lengths = np.minimum(30 + rng.geometric(1.0 / 300.0, proteins), 5000)
    total = int(lengths.sum())
    letters = np.frombuffer(RESIDUES, dtype=np.uint8)[rng.choice(20, size=total, p=FREQ / FREQ.sum())]
    out = []
    at = 0
    for i, n in enumerate(lengths.tolist()):
        seq = letters[at:at + n].tobytes() + b"*"
        at += n
        out.append(b">s%d_p%d synthetic protein len=%d\n" % (j, i, n))
        out.append(b"\n".join(seq[c:c + 60] for c in range(0, len(seq), 60)))
        out.append(b"\n")
    return b"".join(out)


def sample_set(samples: int = 64, proteins: int = 5000) -> list:
    return [sample_text(j, proteins) for j in range(samples)]
