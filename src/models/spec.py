"""Import-path shim for ``from src.models.spec import spectro`` (reference src/solver.py:27,
src/evaluate.py:12).  CUDA implementation: ``aero_b200.spec``."""
from aero_b200.spec import spectro, ispectro  # noqa: F401
