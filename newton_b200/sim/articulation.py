"""Forward kinematics joint_q -> body_q: ``newton.eval_fk``.

Mirrors ``eval_single_articulation_fk`` of the reference (``newton/_src/sim/articulation.py:237-432``):
``X_wc = X_wp * X_pj * X_j(q) * X_cj^-1`` walked in joint order, body twists reported as COM twists.
Called once before the simulation loop by the examples (``example_basic_urdf.py:87``) and on resets.

Runs the FK kernels of the native library through ``nb2_eval_fk`` / ``nb2_eval_fk_masked`` (one warp per articulation;
SURVEY.md §8(f) first "next" row).  CUDA models only, like every call of this package: a model still on the host raises.
(Scene builders that need initial body poses before ``Model.to(device)`` use ``newton_b200.utils.host_fk`` - input
construction, not this function.)
"""

from __future__ import annotations

import torch


def eval_fk(model, joint_q, joint_qd, state, mask=None, indices=None, body_flag_filter: int = 3) -> None:
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
        dev, nb = model.device, int(model.body_count)
        p_jq = C.c_void_p(_abi.ptr(jq, "f32", dev, int(model.joint_coord_count), "joint_q"))
        p_jqd = C.c_void_p(_abi.ptr(jqd, "f32", dev, int(model.joint_dof_count), "joint_qd"))
        p_bq = C.c_void_p(_abi.ptr(state.body_q, "f32", dev, 7 * nb, "state.body_q"))
        p_bqd = C.c_void_p(_abi.ptr(state.body_qd, "f32", dev, 6 * nb, "state.body_qd"))
        with torch.cuda.device(nm.device_index):
            if mask is None and indices is None and int(body_flag_filter) == 3:  # BodyFlags.ALL
                st = _lib.lib().nb2_eval_fk(nm.handle, p_jq, p_jqd, p_bq, p_bqd, _lib.current_stream_ptr(model))
            else:
                st = _lib.lib().nb2_eval_fk_masked(
                    nm.handle, p_jq, p_jqd, p_bq, p_bqd, C.c_void_p(None if mask is None else mask.data_ptr()),
                    C.c_void_p(None if indices is None else indices.data_ptr()), 0 if indices is None else indices.numel(),
                    int(body_flag_filter), _lib.current_stream_ptr(model))
            _lib.check(st, "nb2_eval_fk")
        return
    from .. import _lib

    raise _lib.Nb2Error(
        "newton_b200.eval_fk runs on CUDA devices only (no CPU path). Scene builders that need initial body poses on the host "
        "use newton_b200.utils.host_fk.host_fk; CPU checks use oracle.eval_fk (test infrastructure)."
    )


def eval_ik(model, state, joint_q, joint_qd) -> None:
    """``newton.eval_ik`` (reference ``sim/articulation.py:883-932``): ``state.body_q`` / ``body_qd`` -> generalized
    ``joint_q`` / ``joint_qd`` for every articulated joint, through ``nb2_eval_ik`` (one thread per joint).  CUDA models only;
    like every simulation call of this package there is no CPU path (the oracle under ``oracle/`` is the CPU checker)."""
    import ctypes as C

    from .. import _abi, _lib

    nm = _lib.native_model(model)  # raises for CPU models
    with torch.cuda.device(nm.device_index):
        _lib.check(
            _lib.lib().nb2_eval_ik(
                nm.handle, C.c_void_p(_abi.ptr(state.body_q, "f32", model.device, 7 * int(model.body_count), "state.body_q")),
                C.c_void_p(_abi.ptr(state.body_qd, "f32", model.device, 6 * int(model.body_count), "state.body_qd")),
                C.c_void_p(_abi.ptr(joint_q, "f32", model.device, int(model.joint_coord_count), "joint_q")),
                C.c_void_p(_abi.ptr(joint_qd, "f32", model.device, int(model.joint_dof_count), "joint_qd")),
                _lib.current_stream_ptr(model)),
            "nb2_eval_ik",
        )
