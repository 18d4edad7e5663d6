"""ccqppy_b200 -- B200-native drop-in for the projected-gradient hot path of CCQPpy.

Mirrors `ccqppy/__init__.py:1-4` of the reference: star-exports of `solution_spaces` and
`solvers`; `problem_suite` is imported explicitly by callers, as in the reference's tests.
All arithmetic runs in hand-written sm_100a CUDA kernels behind the C-ABI of
`include/ccqp_b200.h`; there is no CPU fallback."""
__version__ = "0.1.0"

from .solution_spaces import *   # noqa: F401,F403
from .solvers import *           # noqa: F401,F403
from . import solution_spaces, solvers   # noqa: F401
