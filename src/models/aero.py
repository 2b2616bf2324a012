"""Import-path shim: the reference's callers do ``from src.models.aero import Aero``
(reference src/models/modelFactory.py:1, predict.py, test.py).  The implementation lives in
``aero_b200.model``."""
from aero_b200.model import Aero  # noqa: F401
