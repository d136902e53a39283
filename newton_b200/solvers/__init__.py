from .featherstone import SolverFeatherstone
from .solver import SolverBase
from .xpbd import SolverXPBD

__all__ = ["SolverBase", "SolverFeatherstone", "SolverXPBD"]
