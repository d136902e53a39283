"""Forward kinematics joint_q -> body_q: ``newton.eval_fk``.

Mirrors ``eval_single_articulation_fk`` of the reference (``newton/_src/sim/articulation.py:237-432``):
``X_wc = X_wp * X_pj * X_j(q) * X_cj^-1`` walked in joint order, body twists reported as COM twists.
Called once before the simulation loop by the examples (``example_basic_urdf.py:87``) and on resets.

Models on a CUDA device run the ``eval_fk_kernel`` of the native library through ``nb2_eval_fk`` (one thread per
articulation; SURVEY.md §8(f) first "next" row).  Models still on the host - the builder finalizes on the CPU
before ``Model.to(device)`` - use the NumPy walk below, which is scene-construction code, not a simulation path.
"""

from __future__ import annotations

import numpy as np
import torch

from ..utils import xform as X
from .enums import JointType


def eval_fk(model, joint_q, joint_qd, state, mask=None, indices=None) -> None:
    """Write ``state.body_q`` / ``state.body_qd`` from generalized coordinates (reference ``sim/articulation.py:500-574``).

    ``state`` may be the model itself (as in ``newton.eval_fk(model, model.joint_q, model.joint_qd, model)``).
    ``mask`` (bool ``[articulation_count]``) or ``indices`` (int ``[n]``) restrict the update to some articulations;
    bodies of the others keep their values.
    """
    if mask is not None and indices is not None:
        raise ValueError("Cannot specify both mask and indices parameters")
    if state.body_q.is_cuda:
        import ctypes as C

        from .. import _abi, _lib

        nm = _lib.native_model(model)
        jq = joint_q.contiguous()
        jqd = joint_qd.contiguous()
        if mask is not None:
            if mask.dtype != torch.bool or mask.numel() != model.articulation_count:
                raise ValueError(f"Expected Boolean mask with shape ({model.articulation_count},)")
            mask = mask.contiguous()
        if indices is not None:
            indices = torch.as_tensor(indices, dtype=torch.int32, device=state.body_q.device).contiguous()
        with torch.cuda.device(nm.device_index):
            if mask is None and indices is None:
                st = _lib.lib().nb2_eval_fk(nm.handle, C.c_void_p(_abi.ptr(jq)), C.c_void_p(_abi.ptr(jqd)),
                                            C.c_void_p(_abi.ptr(state.body_q)), C.c_void_p(_abi.ptr(state.body_qd)),
                                            _lib.current_stream_ptr(model))
            else:
                st = _lib.lib().nb2_eval_fk_masked(
                    nm.handle, C.c_void_p(_abi.ptr(jq)), C.c_void_p(_abi.ptr(jqd)), C.c_void_p(_abi.ptr(state.body_q)),
                    C.c_void_p(_abi.ptr(state.body_qd)), C.c_void_p(None if mask is None else mask.data_ptr()),
                    C.c_void_p(None if indices is None else indices.data_ptr()), 0 if indices is None else indices.numel(),
                    _lib.current_stream_ptr(model))
            _lib.check(st, "nb2_eval_fk")
        return
    enabled = None
    if mask is not None:
        enabled = np.asarray(mask.detach().cpu().numpy(), dtype=bool)
    elif indices is not None:
        enabled = np.zeros(model.articulation_count, dtype=bool)
        ids = np.asarray(torch.as_tensor(indices).cpu().numpy(), dtype=np.int64)
        enabled[ids[(ids >= 0) & (ids < model.articulation_count)]] = True
    q = joint_q.detach().cpu().numpy().astype(np.float64)
    qd = joint_qd.detach().cpu().numpy().astype(np.float64)
    jt = model.numpy("joint_type")
    parent = model.numpy("joint_parent")
    child = model.numpy("joint_child")
    Xp = model.numpy("joint_X_p").astype(np.float64)
    Xc = model.numpy("joint_X_c").astype(np.float64)
    axis = model.numpy("joint_axis").astype(np.float64)
    qs = model.numpy("joint_q_start")
    qds = model.numpy("joint_qd_start")
    dof_dim = model.numpy("joint_dof_dim")
    art = model.numpy("joint_articulation")
    com = model.numpy("body_com").astype(np.float64)
    body_q = state.body_q.detach().cpu().numpy().astype(np.float64)
    body_qd = state.body_qd.detach().cpu().numpy().astype(np.float64)

    for i in range(model.joint_count):
        if art[i] == -1 or (enabled is not None and not enabled[art[i]]):
            continue
        t = jt[i]
        Xj = X.transform_identity()
        v_lin = np.zeros(3)
        v_ang = np.zeros(3)
        a0, c0 = qds[i], qs[i]
        if t == JointType.PRISMATIC:
            Xj = X.transform(axis[a0] * q[c0])
            v_lin = axis[a0] * qd[a0]
        elif t == JointType.REVOLUTE:
            Xj = X.transform((0, 0, 0), X.quat_from_axis_angle(axis[a0], q[c0]))
            v_ang = axis[a0] * qd[a0]
        elif t == JointType.BALL:
            Xj = X.transform((0, 0, 0), q[c0 : c0 + 4])
            v_ang = qd[a0 : a0 + 3]
        elif t in (JointType.FREE, JointType.DISTANCE):
            Xj = X.transform(q[c0 : c0 + 3], q[c0 + 3 : c0 + 7])
            v_lin = qd[a0 : a0 + 3]
            v_ang = qd[a0 + 3 : a0 + 6]
        elif t == JointType.D6:
            nl, na = dof_dim[i]
            pos = np.zeros(3)
            for k in range(nl):
                pos += axis[a0 + k] * q[c0 + k]
                v_lin += axis[a0 + k] * qd[a0 + k]
            rot = X.quat_identity()
            for k in range(na):
                # sequential rotations about the joint axes (matches compute_{2,3}d_rotational_dofs order)
                rot = X.quat_mul(rot, X.quat_from_axis_angle(axis[a0 + nl + k], q[c0 + nl + k]))
                v_ang += axis[a0 + nl + k] * qd[a0 + nl + k]
            Xj = X.transform(pos, rot)
        X_wpj = Xp[i]
        p = parent[i]
        if p >= 0:
            X_wp = body_q[p]
            X_wpj = X.transform_mul(X_wp, X_wpj)
        X_wcj = X.transform_mul(X_wpj, Xj)
        X_wc = X.transform_mul(X_wcj, X.transform_inverse(Xc[i]))
        x_child = X_wc[:3]
        v_parent_origin = np.zeros(3)
        w_parent = np.zeros(3)
        if p >= 0:
            w_parent = body_qd[p][3:]
            com_p = X.transform_point(body_q[p], com[p])
            v_parent_origin = body_qd[p][:3] + X.cross(w_parent, x_child - com_p)
        lin_w = X.transform_vector(X_wpj, v_lin)
        ang_w = X.transform_vector(X_wpj, v_ang)
        c = child[i]
        if t in (JointType.FREE, JointType.DISTANCE):
            com_c = X.transform_point(X_wc, com[c])
            lin_origin = lin_w + X.cross(ang_w, x_child - com_c)  # COM twist -> origin twist
        else:
            lin_origin = lin_w + X.cross(ang_w, x_child - X_wcj[:3])
        v_o = v_parent_origin + lin_origin
        w = w_parent + ang_w
        body_q[c] = X_wc
        com_c = X.transform_point(X_wc, com[c])
        body_qd[c] = np.concatenate([v_o + X.cross(w, com_c - x_child), w])

    state.body_q.copy_(torch.from_numpy(body_q.astype(np.float32)))
    state.body_qd.copy_(torch.from_numpy(body_qd.astype(np.float32)))


def eval_ik(model, state, joint_q, joint_qd) -> None:
    """``newton.eval_ik`` (reference ``sim/articulation.py:883-932``): ``state.body_q`` / ``body_qd`` -> generalized
    ``joint_q`` / ``joint_qd`` for every articulated joint, through ``nb2_eval_ik`` (one thread per joint).  CUDA models only;
    like every simulation call of this package there is no CPU path (the oracle under ``oracle/`` is the CPU checker)."""
    import ctypes as C

    from .. import _abi, _lib

    nm = _lib.native_model(model)  # raises for CPU models
    with torch.cuda.device(nm.device_index):
        _lib.check(
            _lib.lib().nb2_eval_ik(nm.handle, C.c_void_p(_abi.ptr(state.body_q)), C.c_void_p(_abi.ptr(state.body_qd)),
                                   C.c_void_p(_abi.ptr(joint_q)), C.c_void_p(_abi.ptr(joint_qd)),
                                   _lib.current_stream_ptr(model)),
            "nb2_eval_ik",
        )
