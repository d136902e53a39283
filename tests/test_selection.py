"""ArticulationView / reset / State.assign host logic - CPU part.

Known answers are the reference's own (``newton/tests/test_selection.py``): the ant topology counts (:253-271), the
``[world, articulation, value]`` shapes for 1 ant, 10 worlds x 1 ant, 10 worlds x 3 ants with floating and fixed base
(:273-371), the Model articulation masks (:533-582), the non-contiguous shape selection (:376-441), the masked ``eval_fk``
(:443-531).  On top of that every strided layout the product computes is compared with the oracle's stride-free walk of the
model (``oracle.selection.explicit_ids``), and the index arithmetic of the CUDA copy kernels is run on the host
(``orc_view_copy_product_host``) against the NumPy restatement of the reference kernels.
"""

import re
import warnings

import numpy as np
import pytest
import torch

import newton_b200
from newton_b200 import JointType, Model, ModelBuilder, scenes
from newton_b200.selection import ArticulationView, match_labels
from newton_b200.utils import xform as X

F = Model.AttributeFrequency
FREQ_KEY = {F.JOINT: "joint", F.JOINT_DOF: "dof", F.JOINT_COORD: "coord", F.BODY: "link", F.SHAPE: "shape"}
ATTRS = ["joint_type", "joint_X_p", "joint_dof_dim", "joint_qd", "joint_limit_ke", "joint_axis", "joint_q", "body_q", "body_qd", "body_mass",
         "body_inertia", "shape_margin", "shape_scale", "shape_transform"]


def _layout_numbers(view, freq):
    lay = view.frequency_layouts[freq]
    sel = list(range(lay.slice.start, lay.slice.stop)) if lay.indices is None else lay.indices.tolist()
    return lay, sel


# ---------------------------------------------------------------------------------------------- reference known answers
@pytest.mark.parametrize("floating", [True, False])
def test_selection_shapes_known_answers(floating):
    L, J, S = 9, 9, 13
    D, C = (14, 15) if floating else (8, 8)
    for W, A in ((1, 1), (10, 1), (10, 3)):
        model = scenes.ants_model(W, A, floating=floating, ground=False)
        view = ArticulationView(model, "ant")
        assert (view.count, view.world_count, view.count_per_world) == (W * A, W, A)
        assert view.get_root_transforms(model).shape == (W, A, 7)
        if floating:
            assert view.get_root_velocities(model).shape == (W, A, 6)
        else:
            assert view.get_root_velocities(model) is None
        assert view.get_link_transforms(model).shape == (W, A, L, 7)
        assert view.get_link_velocities(model).shape == (W, A, L, 6)
        assert view.get_dof_positions(model).shape == (W, A, C)
        assert view.get_dof_velocities(model).shape == (W, A, D)
        assert view.get_attribute("body_mass", model).shape == (W, A, L)
        assert view.get_attribute("joint_type", model).shape == (W, A, J)
        assert view.get_attribute("joint_dof_dim", model).shape == (W, A, J, 2)
        assert view.get_attribute("joint_limit_ke", model).shape == (W, A, D)
        assert view.get_attribute("shape_margin", model).shape == (W, A, S)
        assert view.is_floating_base == floating and view.is_fixed_base == (not floating)


def test_model_articulation_mask_known_answers(oracle_lib):
    from oracle import selection as osel

    model = scenes.ants_model(4, 3, ground=False)
    view = ArticulationView(model, "ant")
    ids = view.articulation_ids.numpy()
    assert np.array_equal(view.get_model_articulation_mask().numpy(), np.ones(12, dtype=bool))
    assert np.array_equal(osel.model_articulation_mask(ids, 12), np.ones(12, dtype=bool))
    expected = np.array([0, 0, 0, 1, 1, 1, 1, 1, 1, 0, 0, 0], dtype=bool)
    assert np.array_equal(osel.model_articulation_mask(ids, 12, [0, 1, 1, 0]), expected)
    m = [[0, 1, 0], [1, 0, 1], [1, 1, 1], [0, 0, 0]]
    expected = np.array([0, 1, 0, 1, 0, 1, 1, 1, 1, 0, 0, 0], dtype=bool)
    assert np.array_equal(osel.model_articulation_mask(ids, 12, m), expected)
    # the product resolves the same mask spellings (list / tensor, 1-D / 2-D) before the kernel runs
    assert view._resolve_mask([0, 1, 1, 0]).tolist() == [False, True, True, False]
    assert view._resolve_mask(m).shape == (4, 3)
    assert view._resolve_mask(torch.tensor(m, dtype=torch.bool)).shape == (4, 3)


def test_invalid_masks_are_rejected():
    b = ModelBuilder()
    body = b.add_link(mass=1.0, inertia=np.eye(3))
    b.add_articulation([b.add_joint_free(child=body)], label="robot")
    view = ArticulationView(b.finalize(), "robot")
    for bad in (torch.empty(0, dtype=torch.bool), torch.ones(2, dtype=torch.bool), torch.ones((1, 2), dtype=torch.bool),
                torch.ones((1, 1, 1), dtype=torch.bool), torch.ones(1, dtype=torch.int32)):
        with pytest.raises(ValueError):
            view._resolve_mask(bad)
        with pytest.raises(ValueError):
            view._resolve_world_mask(bad)
    with pytest.raises(ValueError):
        view._resolve_mask([[1, 0, 1]])


def _three_link_robot():
    """base -> link1 -> link2, one shape each with a distinct margin (reference test_selection.py:376-421)."""
    robot = ModelBuilder()
    margins = [0.001, 0.002, 0.003]
    cfgs = []
    for mg in margins:
        cfg = newton_b200.ShapeConfig()
        cfg.margin = mg
        cfgs.append(cfg)
    inertia = np.eye(3) * 0.1
    base = robot.add_link(xform=X.transform((0, 0, 0)), mass=1.0, inertia=inertia, label="base")
    robot.add_shape_box(base, hx=0.1, hy=0.1, hz=0.1, cfg=cfgs[0], label="shape_base")
    link1 = robot.add_link(xform=X.transform((0, 0, 0.5)), mass=0.5, inertia=inertia, label="link1")
    robot.add_shape_capsule(link1, radius=0.05, half_height=0.2, cfg=cfgs[1], label="shape_link1")
    link2 = robot.add_link(xform=X.transform((0, 0, 1.0)), mass=0.3, inertia=inertia, label="link2")
    robot.add_shape_sphere(link2, radius=0.05, cfg=cfgs[2], label="shape_link2")
    j0 = robot.add_joint_free(child=base)
    j1 = robot.add_joint_revolute(base, link1, axis=(0, 1, 0))
    j2 = robot.add_joint_revolute(link1, link2, axis=(0, 1, 0))
    robot.add_articulation([j0, j1, j2], label="robot")
    return robot, margins


def test_noncontiguous_shape_selection_known_answer(oracle_lib):
    from oracle import selection as osel

    robot, margins = _three_link_robot()
    W = 3
    scene = ModelBuilder()
    scene.add_shape_plane()  # a global shape first, so shape indices are offset
    scene.replicate(robot, W)
    model = scene.finalize()
    view = ArticulationView(model, "robot", exclude_links=["link1"])
    assert not view.shapes_contiguous and view.shape_count == 2 and view.link_names == ["base", "link2"]
    lay, sel = _layout_numbers(view, F.SHAPE)
    rows = osel.view_rows(W, 1, lay.offset, lay.stride_between_worlds, lay.stride_within_worlds, sel)
    vals = osel.gather(model.numpy("shape_margin"), rows)
    assert vals.shape == (W, 1, 2)
    for w in range(W):
        np.testing.assert_allclose(vals[w, 0], [margins[0], margins[2]], rtol=0, atol=1e-6)
    # the product refuses to run the index-gather on the host
    with pytest.raises(newton_b200._lib.Nb2Error):
        view.get_attribute("shape_margin", model)


def test_eval_fk_uses_mask_known_answer(oracle_lib):
    """Reference test_selection.py:443-531: FK of one of two translated chains; the other keeps its sentinel values, and the
    written body_qd agrees with a finite difference of body_q."""
    import oracle

    builder = ModelBuilder(up_axis="y", gravity=0.0)

    def chain(label, x_offset):
        inertia = np.eye(3) * 0.1
        base = builder.add_link(mass=1.0, inertia=inertia, com=(0.2, 0.0, 0.0))
        slider = builder.add_link(mass=1.0, inertia=inertia, com=(0.35, 0.0, -0.1))
        j0 = builder.add_joint_revolute(-1, base, axis=(0, 0, 1), parent_xform=X.transform((x_offset, 0.0, 0.0)), child_xform=X.transform((0.0, 0.0, 0.0)))
        j1 = builder.add_joint_prismatic(base, slider, axis=(1, 0, 0), parent_xform=X.transform((1.0, 0.0, 0.4)),
                                         child_xform=X.transform((0.2, 0.0, -0.15)))
        builder.add_articulation([j0, j1], label=label)
        return base, slider, j0, j1

    t_base, t_slider, t_j0, t_j1 = chain("translated_target", 0.0)
    o_base, o_slider, o_j0, o_j1 = chain("translated_other", 5.0)
    model = builder.finalize()
    view = ArticulationView(model, "translated_target")
    assert view.articulation_mask.tolist() == [True, False]
    qs, qds = model.numpy("joint_q_start"), model.numpy("joint_qd_start")
    q, qd = model.numpy("joint_q").copy(), model.numpy("joint_qd").copy()
    q[qs[t_j0]], q[qs[t_j1]], qd[qds[t_j0]], qd[qds[t_j1]] = 0.55, 0.8, 1.1, -0.35
    q[qs[o_j0]], q[qs[o_j1]], qd[qds[o_j0]], qd[qds[o_j1]] = -0.3, 0.25, -0.7, 0.45
    dt = 1.0e-4
    q_next = q.copy()
    for j in (t_j0, t_j1, o_j0, o_j1):
        q_next[qs[j]] += qd[qds[j]] * dt
    results = []
    for qq in (q, q_next):
        st = model.state()
        st.body_q[:, :3] = -99.0
        st.body_q[:, 3:] = torch.tensor([0.0, 0.0, 0.0, 1.0])
        st.body_qd[:] = -77.0
        st.joint_q.copy_(torch.from_numpy(qq))
        st.joint_qd.copy_(torch.from_numpy(qd))
        oracle.eval_fk(model, st.joint_q, st.joint_qd, st, mask=view.get_model_articulation_mask())
        results.append((st.body_q.numpy().copy(), st.body_qd.numpy().copy()))
    (bq, bqd), (bq_next, _) = results
    com = model.numpy("body_com")[t_slider]
    rot = X.quat_rotate(bq[t_slider, 3:].astype(np.float64), com.astype(np.float64))
    origin_vel = bqd[t_slider, :3] - np.cross(bqd[t_slider, 3:], rot)  # COM twist -> origin velocity
    np.testing.assert_allclose((bq_next[t_slider, :3] - bq[t_slider, :3]) / dt, origin_vel, atol=5.0e-3)
    assert not np.array_equal(bq[t_base, :3], [-99.0] * 3)
    for b in (o_base, o_slider):
        assert np.array_equal(bq[b], [-99.0, -99.0, -99.0, 0.0, 0.0, 0.0, 1.0]) and np.array_equal(bqd[b], [-77.0] * 6)
    # index list instead of a mask; out-of-range entries are ignored (sim/articulation.py:462-463)
    st = model.state()
    st.body_q[:, :3] = -99.0
    oracle.eval_fk(model, torch.from_numpy(q), torch.from_numpy(qd), st, indices=[1, 7, -1])
    assert np.array_equal(st.body_q.numpy()[t_base, :3], [-99.0] * 3) and not np.array_equal(st.body_q.numpy()[o_base, :3], [-99.0] * 3)
    with pytest.raises(ValueError):
        oracle.eval_fk(model, st.joint_q, st.joint_qd, st, mask=view.articulation_mask, indices=[0])
    with pytest.raises(ValueError):
        newton_b200.eval_fk(model, st.joint_q, st.joint_qd, st, mask=view.articulation_mask, indices=[0])
    with pytest.raises(newton_b200._lib.Nb2Error):  # the product's eval_fk has no CPU path
        newton_b200.eval_fk(model, st.joint_q, st.joint_qd, st)


# ---------------------------------------------------------------------------------------------- selectors and errors
def test_label_matching_and_errors():
    labels = ["env_0/Robot_A", "env_0/Robot_B", "env_1/Robot_A"]
    assert match_labels(labels, "*/Robot_A") == [0, 2]
    assert match_labels(labels, ["*_B", "env_1/*"]) == [1, 2]
    assert match_labels(labels, re.compile(r"env_[0-9]+/Robot_(A|B)")) == [0, 1, 2]
    assert match_labels(labels, re.compile(r"Robot_A")) == []  # full match
    assert match_labels(labels, [2, 0]) == [2, 0]
    with pytest.raises(TypeError):
        match_labels(labels, ["a", 1])
    with pytest.raises(TypeError):
        match_labels(labels, 3)
    model = scenes.ants_model(2, 2, ground=False)
    with pytest.raises(KeyError):
        ArticulationView(model, "spider")
    with pytest.raises(ValueError):
        ArticulationView(model, [1, 0])  # indices must ascend
    with pytest.raises(ValueError):
        ArticulationView(model, [0, 1, 2])  # 2 + 1 articulations per world
    view = ArticulationView(model, [0, 2])  # the first ant of each world
    assert (view.count, view.count_per_world) == (2, 1)
    assert view.frequency_layouts[F.JOINT].stride_between_worlds == 18
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        ArticulationView(model, "ant", include_joints=[3, 1])
    assert any(issubclass(x.category, DeprecationWarning) for x in w)
    with pytest.raises(ValueError):
        ArticulationView(model, "ant", include_joints=[99])
    v = ArticulationView(model, "ant", include_joints=["hip_*"], exclude_joints=["hip_3"])
    assert v.joint_names == ["hip_1", "hip_2", "hip_4"] and v.joint_dof_names == v.joint_names and not v.joints_contiguous
    v = ArticulationView(model, "ant", include_joint_types=[JointType.FREE])
    assert v.joint_names == ["root"] and v.joint_dof_names == [f"root:{k}" for k in range(6)] and v.joint_coord_count == 7
    v = ArticulationView(model, "ant", include_links=re.compile(r".*_foot"))
    assert v.link_names == ["front_left_foot", "front_right_foot", "back_left_foot", "back_right_foot"]
    assert v.link_shapes == [[0], [1], [2], [3]] and v.shape_names[0] == "front_left_ankle_geom"
    assert v.body_names is v.link_names and v.body_shapes is v.link_shapes


def test_heterogeneous_worlds_are_rejected():
    a, b = scenes.ant_builder(True), scenes.ant_builder(False)
    b.articulation_label[0] = "ant"
    scene = ModelBuilder()
    scene.add_world(a)
    scene.add_world(b)
    with pytest.raises(ValueError, match="not identical"):
        ArticulationView(scene.finalize(), "ant")
    scene = ModelBuilder()
    scene.add_world(a)
    scene.begin_world()
    scene.add_builder(a)
    scene.add_builder(a)
    scene.end_world()
    with pytest.raises(ValueError, match="Varying articulation counts"):
        ArticulationView(scene.finalize(), "ant")


# ---------------------------------------------------------------------------------------------- layouts vs the stride-free walk
VIEWS = [
    dict(),
    dict(exclude_joint_types=[int(JointType.FREE)]),
    dict(exclude_joints=["hip_2", "ankle_3"]),
    dict(exclude_links=["front_right_leg", "back_left_foot"]),
]


@pytest.mark.parametrize("kwargs", VIEWS)
@pytest.mark.parametrize("W,A,floating", [(1, 1, True), (5, 1, True), (4, 3, True), (3, 2, False)])
def test_layouts_equal_explicit_walk(oracle_lib, W, A, floating, kwargs):
    """view[w, a, k] of every attribute == the model row found by walking articulation -> joints -> dofs / bodies -> shapes;
    the same through (1) NumPy gather on the product's layout numbers, (2) the product's zero-copy view where the selection is
    contiguous, (3) the CUDA kernels' index arithmetic executed on the host."""
    from oracle import selection as osel

    model = scenes.ants_model(W, A, floating=floating)  # with a trailing global ground plane
    rng = np.random.default_rng(7)
    for name in ("joint_q", "joint_qd", "body_q", "body_qd", "joint_limit_ke", "shape_margin", "body_mass"):
        t = getattr(model, name)
        t.copy_(torch.from_numpy(rng.normal(size=tuple(t.shape)).astype(np.float32)))
    view = ArticulationView(model, "ant", **kwargs)
    ids = osel.explicit_ids(model, "ant", **kwargs)
    assert np.array_equal(view.articulation_ids.numpy(), np.asarray(ids["articulation"]))
    for name in ATTRS:
        attrib = model.numpy(name)
        freq = model.get_attribute_frequency(name)
        expected = osel.take(attrib, ids[FREQ_KEY[freq]])
        lay, sel = _layout_numbers(view, freq)
        rows = osel.view_rows(W, A, lay.offset, lay.stride_between_worlds, lay.stride_within_worlds, sel)
        assert np.array_equal(osel.gather(attrib, rows), expected), name
        if lay.is_contiguous:
            got = view.get_attribute(name, model)
            assert got.data_ptr() == getattr(model, name).data_ptr() + 4 * (getattr(model, name).stride(0) * (lay.offset + lay.slice.start)) or got.numel() == 0
            assert np.array_equal(got.numpy(), expected), name
        # CUDA index arithmetic on the host, 32- and 64-bit instantiations
        row_words = int(np.prod(attrib.shape[1:], dtype=np.int64))
        L = dict(world_count=W, count_per_world=A, value_count=len(sel), row_words=row_words, offset=lay.offset,
                 stride_between_worlds=lay.stride_between_worlds, stride_within_worlds=lay.stride_within_worlds,
                 slice_start=sel[0] if (lay.is_contiguous and sel) else 0, indices=None if lay.is_contiguous else sel)
        for wide in (False, True):
            out = np.zeros(expected.shape, dtype=attrib.dtype)
            osel.view_copy_product_host(np.ascontiguousarray(attrib), L, out, gather_=True, wide=wide)
            assert np.array_equal(out, expected), name
        # masked scatter: product arithmetic == NumPy restatement of the reference kernels, 1-D and 2-D masks
        values = rng.normal(size=expected.shape).astype(np.float32).view(np.uint32).astype(np.uint32).view(attrib.dtype) \
            if attrib.dtype == np.float32 else rng.integers(0, 100, size=expected.shape).astype(attrib.dtype)
        for mask in (None, rng.random(W) < 0.5, rng.random((W, A)) < 0.5):
            ref = attrib.copy()
            osel.scatter_masked(ref, rows, values, mask)
            got = np.ascontiguousarray(attrib.copy())
            osel.view_copy_product_host(got, L, np.ascontiguousarray(values), mask=mask, gather_=False)
            assert np.array_equal(got.view(np.uint32), ref.view(np.uint32)), (name, None if mask is None else mask.shape)
            if mask is not None and not mask.any():
                assert np.array_equal(got.view(np.uint32), attrib.view(np.uint32))


def test_root_slices(oracle_lib):
    """get_root_transforms / velocities address joint_q[:7] / joint_qd[:6] of each articulation (floating base) or joint_X_p of
    the root joint (fixed base) (reference selection.py:1480-1558)."""
    from oracle import selection as osel

    for floating in (True, False):
        model = scenes.ants_model(3, 2, floating=floating)
        ids = osel.explicit_ids(model, "ant")
        view = ArticulationView(model, "ant")
        root = view.get_root_transforms(model).numpy()
        if floating:
            exp = osel.take(model.numpy("joint_q"), [[row[:7] for row in world] for world in ids["coord"]])
            assert np.array_equal(root, exp)
            exp = osel.take(model.numpy("joint_qd"), [[row[:6] for row in world] for world in ids["dof"]])
            assert np.array_equal(view.get_root_velocities(model).numpy(), exp)
            view.get_root_transforms(model)[1, 1, 2] = 5.0  # a view: writes land in the model array
            assert model.joint_q[ids["coord"][1][1][2]] == 5.0
        else:
            exp = osel.take(model.numpy("joint_X_p"), [[row[:1] for row in world] for world in ids["joint"]])[:, :, 0]
            assert np.array_equal(root, exp) and root.shape == (3, 2, 7)
        assert view.get_dof_forces(model.control()).shape == (3, 2, 14 if floating else 8)


def test_set_refuses_cpu_and_checks_shapes():
    model = scenes.ants_model(2, 1)
    view = ArticulationView(model, "ant")
    state = model.state()
    with pytest.raises(newton_b200._lib.Nb2Error):
        view.set_dof_positions(state, np.zeros((2, 1, 15), dtype=np.float32))
    with pytest.raises(ValueError):
        view.set_dof_positions(state, np.zeros((2, 1, 14), dtype=np.float32))
    view.set_dof_positions(state, view.get_dof_positions(state))  # in place: nothing to copy, no kernel needed
    with pytest.raises(NotImplementedError):
        view.set_attribute("joint_enabled", model, np.ones((2, 1, 9), dtype=bool))  # 1-byte attribute
    with pytest.raises(KeyError):
        view.get_attribute("gravity", model)


# ---------------------------------------------------------------------------------------------- reset mask, State.assign
def test_reset_world_mask_normalisation():
    from newton_b200.solvers.solver import normalize_reset_world_mask as norm

    assert norm(None, world_count=3, device="cpu") is None
    m = torch.tensor([True, False, True, False])
    assert norm(m, world_count=3, device="cpu") is m
    with pytest.warns(DeprecationWarning):
        out = norm(torch.tensor([True, False, True]), world_count=3, device="cpu", allow_legacy=True)
    assert out.tolist() == [True, False, True, False]
    with pytest.raises(ValueError):
        norm(torch.tensor([True, False, True]), world_count=3, device="cpu")
    with pytest.raises(ValueError):
        norm(torch.ones(6, dtype=torch.bool), world_count=3, device="cpu", allow_legacy=True)
    with pytest.raises(ValueError):
        norm(torch.ones((2, 2), dtype=torch.bool), world_count=3, device="cpu")
    with pytest.raises(TypeError):
        norm(torch.ones(4, dtype=torch.int32), world_count=3, device="cpu")
    with pytest.raises(TypeError):
        norm([True] * 4, world_count=3, device="cpu")


def test_state_assign():
    model = scenes.ants_model(2, 1)
    s0, s1 = model.state(), model.state()
    s1.joint_q += 1.0
    s1.body_qd += 2.0
    s0.assign(s1)
    assert torch.equal(s0.joint_q, s1.joint_q) and torch.equal(s0.body_qd, s1.body_qd) and s0.joint_q.data_ptr() != s1.joint_q.data_ptr()
    s1.body_parent_f = torch.zeros_like(s1.body_qd)
    with pytest.raises(ValueError):
        s0.assign(s1)


def test_view_on_a_shard_addresses_the_same_rows(oracle_lib):
    """Multi-GPU: each rank builds its ArticulationView on ``model.shard(rank, n)``; the view must address that rank's slice of the
    monolithic view (labels, body -> shape lists and filter pairs travel with the shard)."""
    from oracle import broad_phase as bp

    model = scenes.ants_model(6, 2)
    rng = np.random.default_rng(1)
    model.joint_q.copy_(torch.from_numpy(rng.normal(size=tuple(model.joint_q.shape)).astype(np.float32)))
    model.shape_margin.copy_(torch.from_numpy(rng.random(tuple(model.shape_margin.shape)).astype(np.float32)))
    whole = ArticulationView(model, "ant", exclude_links=["front_right_leg"])
    for rank in range(3):
        shard = model.shard(rank, 3)
        view = ArticulationView(shard, "ant", exclude_links=["front_right_leg"])
        assert (view.world_count, view.count_per_world) == (2, 2) and view.link_names == whole.link_names and view.shape_names == whole.shape_names
        assert torch.equal(view.get_dof_positions(shard), whole.get_dof_positions(model)[2 * rank : 2 * rank + 2])
        lay = view.frequency_layouts[F.SHAPE]
        wlay = whole.frequency_layouts[F.SHAPE]
        assert (lay.stride_between_worlds, lay.stride_within_worlds) == (wlay.stride_between_worlds, wlay.stride_within_worlds)
        assert torch.equal(lay.indices, wlay.indices)
        explicit = {tuple(p) for p in shard.numpy("shape_contact_pairs").tolist()}
        assert explicit == bp.model_nxn_pairs(shard, shard.shape_collision_filter_pairs)  # the shard's pair list is self-consistent too
