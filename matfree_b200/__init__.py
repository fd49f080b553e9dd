"""matfree_b200 -- B200-native stochastic Lanczos quadrature behind matfree's API.

Drop-in for the SLQ hot path of pnkraemer/matfree: `stochtrace`, `funm`,
`decomp` keep the reference's names, signatures and error behaviour; `ops`
adds the three registered operator kinds (dense, CSR, Gram).  All arithmetic
runs in hand-written sm_100a CUDA kernels inside `libmatfree_b200.so`
(`include/matfree_b200.h`); there is no CPU fallback -- if the library or a
CUDA device is missing, calls raise.
"""

from matfree_b200 import config, decomp, eig, funm, ops, stochtrace  # noqa: F401
from matfree_b200.backend import prng  # noqa: F401

__all__ = ["config", "decomp", "eig", "funm", "ops", "stochtrace", "prng"]
