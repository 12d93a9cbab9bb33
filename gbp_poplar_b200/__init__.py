"""gbp_poplar_b200 -- B200-native GBP bundle-adjustment hot path.

Product code: CUDA kernels + C ABI under csrc/ (built into libgbp_cuda.so),
and thin ctypes mirrors of the reference's host-side interface (host.py,
engine.py).  Nothing here imports the test-only CPU oracle under oracle/.
"""
from . import _capi  # noqa: F401
from .engine import GBPEngine, GBPGroup, default_opts  # noqa: F401
from .host import BALProblem, Setup, cli_options, MODE_BA, MODE_SLAM  # noqa: F401

__all__ = ["GBPEngine", "GBPGroup", "default_opts", "BALProblem", "Setup", "cli_options", "MODE_BA", "MODE_SLAM"]
