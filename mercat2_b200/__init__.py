"""mercat2_b200 -- B200-native (sm_100a CUDA) drop-in for MerCat2's k-mer counting hot path.

Import surface mirrors the reference package ``mercat2_lib`` for the path this repo replaces:

    from mercat2_b200 import mercat2_kmers, mercat2_Chunker, mercat2_metrics

``mercat2_kmers.find_kmers`` / ``calculateKmerCount``, ``mercat2_Chunker.Chunker`` and the
``mercat2_metrics`` functions keep the reference's names, arguments and return types; the work is
done by hand-written CUDA kernels behind the C ABI in ``include/mercat2_b200.h``.  There is no CPU
fallback: without the built library and a CUDA device these calls raise.
"""
__version__ = "0.1.0"

from . import _native  # noqa: F401
from ._native import Engine, Mc2Error, NonAsciiError, bind_to_gpu_numa, default_engine  # noqa: F401
from . import mercat2_kmers, mercat2_Chunker, mercat2_metrics, mercat2_fasta, pipeline  # noqa: F401
