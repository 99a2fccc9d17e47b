"""Drop-in shim: `from condition.canny import CannyDetector` (reference condition/canny.py) resolves to the GPU implementation."""
from controlar_b200.condition.canny import *  # noqa: F401,F403
from controlar_b200.condition.canny import CannyDetector, canny_cuda  # noqa: F401
