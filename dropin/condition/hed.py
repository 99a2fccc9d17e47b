"""Drop-in shim: `from condition.hed import HEDdetector` (reference condition/hed.py) resolves to the GPU implementation."""
from controlar_b200.condition.hed import *  # noqa: F401,F403
from controlar_b200.condition.hed import HEDdetector, ControlNetHED_Apache2, DoubleConvBlock  # noqa: F401
