"""TEST INFRASTRUCTURE ONLY — a stand-in for the `mpi4py` package.

The reference (helmholtz-analytics/heat) imports `mpi4py.MPI` at module load
(heat/core/communication.py:12) and this image has neither mpi4py nor an MPI
library.  Putting this directory on PYTHONPATH lets the *unmodified* reference
import and run so that oracle/generate_golden.py can record its outputs.

Nothing under heat_b200/ imports this package.
"""
from . import MPI  # noqa: F401
