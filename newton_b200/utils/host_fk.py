"""Initial body poses for freshly built scenes, on the host (NumPy, float64 -> float32).

INPUT CONSTRUCTION, not a simulation path: ``newton_b200.scenes`` builds every model on the CPU, and the reference examples
call ``newton.eval_fk(model, model.joint_q, model.joint_qd, model)`` once after ``finalize()`` so that ``model.body_q`` reflects
the initial joint angles (``example_basic_urdf.py:87``).  This helper plays that role for the synthetic scenes before
``Model.to(device)``, so that the CPU oracle and the CUDA path start from the *same* arrays.  It is not ``newton_b200.eval_fk``
(CUDA only, ``sim/articulation.py``) and nothing compares results against it: the CUDA FK is checked against the oracle's FK.
Same walk as ``eval_single_articulation_fk`` (``newton/_src/sim/articulation.py:237-432``).
"""

from __future__ import annotations

import numpy as np
import torch

from ..sim.enums import JointType
from . import xform as X


def host_fk(model, joint_q, joint_qd, state) -> None:
    """Write ``state.body_q`` / ``state.body_qd`` (host tensors) from ``joint_q`` / ``joint_qd``; ``state`` may be the model."""
    if state.body_q.is_cuda:
        raise ValueError("host_fk is for models still on the host; use newton_b200.eval_fk on CUDA models")
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
        if art[i] == -1:
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
