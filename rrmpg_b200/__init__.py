"""rrmpg_b200 -- B200-native ensemble rainfall-runoff engine.

Drop-in for the hot path of kratzert/RRMPG: ``rrmpg_b200.models`` mirrors ``rrmpg.models``
(ABCModel, HBVEdu, GR4J, Cemaneige, CemaneigeGR4J) and ``rrmpg_b200.tools.monte_carlo`` mirrors
``rrmpg.tools.monte_carlo``; the per-member numba loops underneath are replaced by hand-written
sm_100a CUDA kernels behind the C ABI of ``include/rrmpg_b200.h`` (``librrmpg_b200.so``).
``rrmpg_b200.engine`` is the array-level API (numpy = host buffers, torch CUDA tensors = device
buffers).  There is no CPU fallback.
"""
__version__ = "0.1.0"

from . import engine, models, tools, utils  # noqa: F401,E402
