#!/usr/bin/env python
"""Entry point with the reference's name (bin/mercat2.py): same flags for the k-mer counting path."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from mercat2_b200.cli import mercat_main  # noqa: E402

if __name__ == "__main__":
    sys.exit(mercat_main())
