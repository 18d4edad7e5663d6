"""`ccqppy` -- the reference's import name, served by the B200-native implementation.

A user of the reference switches by putting this repository on `sys.path`: `from ccqppy import solvers,
solution_spaces, problem_suite` (what /root/reference/tests/test_module.py:5-8 and the README do) then
resolves to the CUDA-backed modules of `ccqppy_b200`.  Mirrors ccqppy/__init__.py:1-4 of the reference
(star-exports of solution_spaces and solvers).  No code lives here."""
import sys

import ccqppy_b200
from ccqppy_b200 import *   # noqa: F401,F403
from ccqppy_b200 import problem_suite, solution_spaces, solvers   # noqa: F401

__version__ = ccqppy_b200.__version__
for _name in ("problem_suite", "solution_spaces", "solvers"):
    sys.modules[__name__ + "." + _name] = getattr(ccqppy_b200, _name)
del _name
