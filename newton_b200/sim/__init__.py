from .articulation import eval_fk, eval_ik
from .builder import JointDofConfig, ModelBuilder, ShapeConfig
from .enums import MAXVAL, BodyFlags, GeoType, JointType, ModelFlags, ShapeFlags, StateFlags
from .model import Contacts, Control, Model, State

__all__ = [
    "MAXVAL", "BodyFlags", "Contacts", "Control", "GeoType", "JointDofConfig", "JointType", "Model",
    "ModelBuilder", "ModelFlags", "ShapeConfig", "ShapeFlags", "State", "StateFlags", "eval_fk", "eval_ik",
]
