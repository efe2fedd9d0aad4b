"""NumPy stand-in for the `mlx` package (test infrastructure only; see core.py)."""
from . import core  # noqa: F401
from . import nn  # noqa: F401
