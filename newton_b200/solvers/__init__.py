from .solver import SolverBase
from .xpbd import SolverXPBD

__all__ = ["SolverBase", "SolverXPBD"]
