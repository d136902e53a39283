"""ctypes mirror of ``include/newton_b200.h`` and pointer marshalling for Model/State/Control/Contacts.

Arrays may be ``torch.Tensor`` (``data_ptr()``), Warp arrays (``.ptr``) or NumPy arrays
(``.ctypes.data``): the boundary only sees raw addresses (SURVEY.md §8(b) "Python wrapper obtains
pointers from wp.array.ptr, torch.Tensor.data_ptr(), or NumPy ctypes.data").
"""

from __future__ import annotations

import ctypes as C

import numpy as np

c_float_p = C.c_void_p
c_int_p = C.c_void_p


def ptr(a) -> int | None:
    """Raw address of an array-like (None for missing / empty arrays)."""
    if a is None:
        return None
    if hasattr(a, "data_ptr"):
        if a.numel() == 0:
            return None
        if not a.is_contiguous():
            raise ValueError("arrays crossing the C-ABI must be contiguous")
        return a.data_ptr()
    if hasattr(a, "ptr"):  # wp.array
        return a.ptr
    if isinstance(a, np.ndarray):
        if a.size == 0:
            return None
        if not a.flags["C_CONTIGUOUS"]:
            raise ValueError("arrays crossing the C-ABI must be contiguous")
        return a.ctypes.data
    raise TypeError(f"cannot take the address of {type(a)}")


class ModelDesc(C.Structure):
    _fields_ = [
        ("world_count", C.c_int32),
        ("body_count", C.c_int32),
        ("joint_count", C.c_int32),
        ("joint_dof_count", C.c_int32),
        ("joint_coord_count", C.c_int32),
        ("shape_count", C.c_int32),
        ("shape_pair_count", C.c_int32),
        ("articulation_count", C.c_int32),
        ("body_com", C.c_void_p),
        ("body_mass", C.c_void_p),
        ("body_inv_mass", C.c_void_p),
        ("body_inertia", C.c_void_p),
        ("body_inv_inertia", C.c_void_p),
        ("body_flags", C.c_void_p),
        ("body_world", C.c_void_p),
        ("body_world_start", C.c_void_p),
        ("joint_type", C.c_void_p),
        ("joint_enabled", C.c_void_p),
        ("joint_parent", C.c_void_p),
        ("joint_child", C.c_void_p),
        ("joint_ancestor", C.c_void_p),
        ("joint_articulation", C.c_void_p),
        ("joint_X_p", C.c_void_p),
        ("joint_X_c", C.c_void_p),
        ("joint_q_start", C.c_void_p),
        ("joint_qd_start", C.c_void_p),
        ("joint_target_q_start", C.c_void_p),
        ("joint_dof_dim", C.c_void_p),
        ("joint_world_start", C.c_void_p),
        ("joint_axis", C.c_void_p),
        ("joint_limit_lower", C.c_void_p),
        ("joint_limit_upper", C.c_void_p),
        ("joint_limit_ke", C.c_void_p),
        ("joint_limit_kd", C.c_void_p),
        ("joint_target_ke", C.c_void_p),
        ("joint_target_kd", C.c_void_p),
        ("joint_armature", C.c_void_p),
        ("joint_damping", C.c_void_p),
        ("articulation_start", C.c_void_p),
        ("shape_body", C.c_void_p),
        ("shape_type", C.c_void_p),
        ("shape_transform", C.c_void_p),
        ("shape_scale", C.c_void_p),
        ("shape_margin", C.c_void_p),
        ("shape_gap", C.c_void_p),
        ("shape_collision_radius", C.c_void_p),
        ("shape_flags", C.c_void_p),
        ("shape_world", C.c_void_p),
        ("shape_world_start", C.c_void_p),
        ("shape_material_ke", C.c_void_p),
        ("shape_material_kd", C.c_void_p),
        ("shape_material_kf", C.c_void_p),
        ("shape_material_ka", C.c_void_p),
        ("shape_material_mu", C.c_void_p),
        ("shape_material_mu_torsional", C.c_void_p),
        ("shape_material_mu_rolling", C.c_void_p),
        ("shape_material_restitution", C.c_void_p),
        ("shape_contact_pairs", C.c_void_p),
        ("gravity", C.c_void_p),
        ("gravity_count", C.c_int32),
    ]


class StateView(C.Structure):
    _fields_ = [
        ("body_q", C.c_void_p),
        ("body_qd", C.c_void_p),
        ("body_f", C.c_void_p),
        ("joint_q", C.c_void_p),
        ("joint_qd", C.c_void_p),
        ("body_parent_f", C.c_void_p),
    ]


class ControlView(C.Structure):
    _fields_ = [
        ("joint_f", C.c_void_p),
        ("joint_target_q", C.c_void_p),
        ("joint_target_qd", C.c_void_p),
        ("joint_act", C.c_void_p),
    ]


class ContactsView(C.Structure):
    _fields_ = [
        ("rigid_contact_max", C.c_int32),
        ("rigid_contact_count", C.c_void_p),
        ("shape0", C.c_void_p),
        ("shape1", C.c_void_p),
        ("point0", C.c_void_p),
        ("point1", C.c_void_p),
        ("offset0", C.c_void_p),
        ("offset1", C.c_void_p),
        ("normal", C.c_void_p),
        ("margin0", C.c_void_p),
        ("margin1", C.c_void_p),
        ("tids", C.c_void_p),
        ("force", C.c_void_p),
    ]


class XPBDParams(C.Structure):
    _fields_ = [
        ("iterations", C.c_int32),
        ("joint_linear_relaxation", C.c_float),
        ("joint_angular_relaxation", C.c_float),
        ("joint_linear_compliance", C.c_float),
        ("joint_angular_compliance", C.c_float),
        ("rigid_contact_relaxation", C.c_float),
        ("rigid_contact_con_weighting", C.c_int32),
        ("angular_damping", C.c_float),
        ("enable_restitution", C.c_int32),
        ("compute_body_velocity_from_position_delta", C.c_int32),
    ]


class FeatherstoneParams(C.Structure):
    _fields_ = [
        ("angular_damping", C.c_float),
        ("update_mass_matrix_interval", C.c_int32),
        ("friction_smoothing", C.c_float),
    ]


_COUNT_FIELDS = ("world_count", "body_count", "joint_count", "joint_dof_count", "joint_coord_count", "shape_count",
                 "articulation_count")


class ViewLayout(C.Structure):
    """``nb2_view_layout`` (include/newton_b200.h): how an ArticulationView addresses an attribute array."""

    _fields_ = [
        ("world_count", C.c_int32),
        ("count_per_world", C.c_int32),
        ("value_count", C.c_int32),
        ("row_words", C.c_int32),
        ("offset", C.c_int32),
        ("stride_between_worlds", C.c_int32),
        ("stride_within_worlds", C.c_int32),
        ("slice_start", C.c_int32),
        ("indices", C.c_void_p),
    ]


def model_desc(model) -> ModelDesc:
    """Fill a :class:`ModelDesc` with the addresses of ``model``'s arrays (borrowed, not copied)."""
    d = ModelDesc()
    for n in _COUNT_FIELDS:
        setattr(d, n, int(getattr(model, n)))
    d.shape_pair_count = int(getattr(model, "shape_contact_pair_count", 0))
    for name, _ in ModelDesc._fields_:
        if name in _COUNT_FIELDS or name in ("shape_pair_count", "gravity_count"):
            continue
        setattr(d, name, ptr(getattr(model, name, None)))
    g = model.gravity
    d.gravity_count = int(g.shape[0])
    return d


def state_view(state) -> StateView:
    v = StateView()
    for name, _ in StateView._fields_:
        setattr(v, name, ptr(getattr(state, name, None)))
    return v


def control_view(control) -> ControlView:
    v = ControlView()
    for name, _ in ControlView._fields_:
        setattr(v, name, ptr(getattr(control, name, None)))
    return v


def contacts_view(contacts) -> ContactsView:
    v = ContactsView()
    v.rigid_contact_max = int(contacts.rigid_contact_max)
    v.rigid_contact_count = ptr(contacts.rigid_contact_count)
    for short in ("shape0", "shape1", "point0", "point1", "offset0", "offset1", "normal", "margin0", "margin1", "tids"):
        setattr(v, short, ptr(getattr(contacts, "rigid_contact_" + short)))
    force = getattr(contacts, "force", None)
    v.force = ptr(force) if force is not None else None
    return v
