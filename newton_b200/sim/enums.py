"""Enumerations shared with the reference data model.

Values mirror the reference so that arrays produced by either implementation are
interchangeable at the drop-in boundary:

* ``GeoType``      - reference ``newton/_src/geometry/types.py:70-111``
* ``JointType``    - reference ``newton/_src/sim/enums.py:178-212``
* ``BodyFlags``    - reference ``newton/_src/sim/enums.py:119-145``
* ``ShapeFlags``   - reference ``newton/_src/geometry/flags.py:20-40``
* ``ModelFlags`` / ``StateFlags`` - reference ``newton/_src/sim/enums.py:8-117``
"""

from enum import IntEnum, IntFlag


class GeoType(IntEnum):
    NONE = 0
    PLANE = 1
    HFIELD = 2
    SPHERE = 3
    CAPSULE = 4
    ELLIPSOID = 5
    CYLINDER = 6
    BOX = 7
    MESH = 8
    CONE = 9
    CONVEX_MESH = 10
    GAUSSIAN = 11


class JointType(IntEnum):
    PRISMATIC = 0
    REVOLUTE = 1
    BALL = 2
    FIXED = 3
    FREE = 4
    DISTANCE = 5
    D6 = 6
    CABLE = 7
    ROD = 7

    def dof_count(self, num_axes: int) -> tuple[int, int]:
        """(dof_count, coord_count) of this joint type (reference ``enums.py:214-250``)."""
        if self == JointType.BALL:
            return 3, 4
        if self in (JointType.FREE, JointType.DISTANCE):
            return 6, 7
        if self == JointType.FIXED:
            return 0, 0
        return num_axes, num_axes


class BodyFlags(IntFlag):
    DYNAMIC = 1 << 0
    KINEMATIC = 1 << 1
    PROXY = 1 << 2
    ALL = DYNAMIC | KINEMATIC | PROXY


class ShapeFlags(IntFlag):
    VISIBLE = 1 << 0
    COLLIDE_SHAPES = 1 << 1
    COLLIDE_PARTICLES = 1 << 2
    SITE = 1 << 3
    HYDROELASTIC = 1 << 4


class ModelFlags(IntFlag):
    JOINT_PROPERTIES = 1 << 0
    JOINT_DOF_PROPERTIES = 1 << 1
    BODY_PROPERTIES = 1 << 2
    BODY_INERTIAL_PROPERTIES = 1 << 3
    SHAPE_PROPERTIES = 1 << 4
    MODEL_PROPERTIES = 1 << 5
    CONSTRAINT_PROPERTIES = 1 << 6
    TENDON_PROPERTIES = 1 << 7
    ACTUATOR_PROPERTIES = 1 << 8
    ALL = (1 << 9) - 1


class StateFlags(IntFlag):
    NONE = 0
    JOINT_Q = 1 << 0
    JOINT_QD = 1 << 1
    BODY_Q = 1 << 2
    BODY_QD = 1 << 3
    PARTICLE_Q = 1 << 4
    PARTICLE_QD = 1 << 5
    BODY_F = 1 << 6
    PARTICLE_F = 1 << 7
    JOINT_F = 1 << 8
    BODY = BODY_Q | BODY_QD
    PARTICLE = PARTICLE_Q | PARTICLE_QD
    JOINT = JOINT_Q | JOINT_QD
    FORCE = BODY_F | PARTICLE_F | JOINT_F
    ALL = BODY | PARTICLE | JOINT | FORCE


MAXVAL = 1.0e10  # reference newton/_src/core/types.py:72
