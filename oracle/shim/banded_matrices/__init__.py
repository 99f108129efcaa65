"""numpy/SciPy stand-in for `banded_matrices` (secondmind-labs, wheel 0.0.6, branch
awav/fix-banded-hashable-tensor — reference README.md:38,49).  TEST INFRASTRUCTURE ONLY."""
from . import banded  # noqa: F401
