"""More of the reference's own ArticulationView known answers (``newton/tests/test_selection.py:34-251``), on CPU models:
selector spellings, error messages, empty selections, fixed-joint-only articulations, base classification, label handling,
doubly driven bodies."""

import re
import warnings

import numpy as np
import pytest

from newton_b200 import JointType, ModelBuilder
from newton_b200.selection import ArticulationView

I3 = np.eye(3) * 0.1


def test_compiled_regex_selectors():
    builder = ModelBuilder()
    for label in ["/World/envs/env_0/Robot_A", "/World/envs/env_0/Robot_B", "/World/envs/env_0/Robot_C", "/World/envs/env_0/Prop"]:
        base = builder.add_link(mass=1.0, inertia=I3, label=f"{label}/base")
        left_foot = builder.add_link(mass=1.0, inertia=I3, label=f"{label}/LF_FOOT")
        right_foot = builder.add_link(mass=1.0, inertia=I3, label=f"{label}/RF_FOOT")
        fixed_mount = builder.add_joint_free(child=base, label=f"{label}/fixed_mount")
        left_hip = builder.add_joint_revolute(base, left_foot, label=f"{label}/LF_HIP")
        right_hip = builder.add_joint_revolute(base, right_foot, label=f"{label}/RF_HIP")
        builder.add_articulation([fixed_mount, left_hip, right_hip], label=label)
    model = builder.finalize(device="cpu")
    view = ArticulationView(model, pattern=re.compile(r"/World/envs/env_[0-9]+/Robot_(A|B|C)"), include_links=re.compile(r"(LF|RF)_FOOT"),
                            exclude_joints=re.compile(r"fixed_.*"))
    assert view.articulation_ids.tolist() == [[0, 1, 2]]
    assert view.link_names == ["LF_FOOT", "RF_FOOT"] and view.joint_names == ["LF_HIP", "RF_HIP"]
    assert view.link_count == 2 and view.joint_count == 2
    with pytest.raises(KeyError, match="No articulations matching pattern"):
        ArticulationView(model, pattern=re.compile(r"/World/envs/env_[0-9]+/Robot_Z"))


def test_articulation_selector_lists():
    builder = ModelBuilder()
    for label in ["robot_a", "robot_b", "prop"]:
        body = builder.add_link(mass=1.0, inertia=I3, label=f"{label}/body")
        builder.add_articulation([builder.add_joint_free(child=body, label=f"{label}/joint")], label=label)
    model = builder.finalize()
    assert ArticulationView(model, pattern=["robot_*", "prop"]).articulation_ids.tolist() == [[0, 1, 2]]
    assert ArticulationView(model, pattern=[0, 2]).articulation_ids.tolist() == [[0, 2]]
    with pytest.raises(ValueError, match="must be unique and in ascending order"):
        ArticulationView(model, pattern=[2, 0])
    with pytest.raises(ValueError, match="must be unique and in ascending order"):
        ArticulationView(model, pattern=[0, 0])
    with pytest.raises(ValueError, match=r"must be in range \[0, 3\)"):
        ArticulationView(model, pattern=[3])
    with pytest.raises(ValueError, match=r"must be in range \[0, 1\)"):
        ArticulationView(model, pattern="robot_a", include_joints=[1])
    with pytest.raises(ValueError, match=r"must be in range \[0, 1\)"):
        ArticulationView(model, pattern="robot_a", include_links=[1])


def test_no_match():
    builder = ModelBuilder()
    builder.add_body(mass=1.0, inertia=I3)
    with pytest.raises(KeyError):
        ArticulationView(builder.finalize(), pattern="no_match")


def test_unsorted_include_indices_deprecated():
    builder = ModelBuilder()
    root = builder.add_link(mass=1.0, inertia=I3, label="root")
    middle = builder.add_link(mass=1.0, inertia=I3, label="middle")
    tip = builder.add_link(mass=1.0, inertia=I3, label="tip")
    joints = [builder.add_joint_free(child=root, label="root_joint"), builder.add_joint_revolute(root, middle, label="middle_joint"),
              builder.add_joint_revolute(middle, tip, label="tip_joint")]
    builder.add_articulation(joints, label="robot")
    model = builder.finalize()
    with pytest.warns(DeprecationWarning, match="include_joints"):
        joint_view = ArticulationView(model, "robot", include_joints=[2, 0])
    assert joint_view.joint_names == ["root_joint", "tip_joint"]
    with pytest.warns(DeprecationWarning, match="include_links"):
        link_view = ArticulationView(model, "robot", include_links=[2, 0])
    assert link_view.link_names == ["root", "tip"]
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        ArticulationView(model, "robot", include_joints=[0, 2])  # sorted: no warning


def test_empty_selection():
    builder = ModelBuilder()
    body = builder.add_link(mass=1.0, inertia=I3)
    builder.add_articulation([builder.add_joint_free(child=body)], label="my_articulation")
    model = builder.finalize()
    control = model.control()
    selection = ArticulationView(model, pattern="my_articulation", exclude_joint_types=[JointType.FREE])
    assert selection.count == 1
    assert selection.get_root_transforms(model).shape == (1, 1, 7)
    assert selection.get_dof_positions(model).shape == (1, 1, 0)
    assert selection.get_dof_velocities(model).shape == (1, 1, 0)
    assert selection.get_dof_forces(control).shape == (1, 1, 0)
    selection.set_dof_positions(model, np.zeros((1, 1, 0), dtype=np.float32))  # nothing to write, nothing launched


def test_fixed_joint_only_articulation():
    """Reference regression test for its issue #920."""
    builder = ModelBuilder()
    parent = builder.add_link(mass=1.0, inertia=I3)
    child = builder.add_link(mass=1.0, inertia=I3)
    j0 = builder.add_joint_fixed(-1, parent)
    j1 = builder.add_joint_fixed(parent, child)
    builder.add_articulation([j0, j1], label="fixed_only")
    model = builder.finalize()
    state, control = model.state(), model.control()
    view = ArticulationView(model, pattern="fixed_only")
    assert view.count == 1 and view.joint_dof_count == 0 and view.joint_coord_count == 0
    assert view.get_root_transforms(model).shape == (1, 1, 7)
    assert view.get_dof_positions(state).shape == (1, 1, 0)
    assert view.get_dof_velocities(state).shape == (1, 1, 0)
    assert view.get_dof_forces(control).shape == (1, 1, 0)
    assert view.get_root_velocities(state) is None


@pytest.mark.parametrize("root_kind,expected_fixed,expected_floating", [("fixed", True, False), ("locked_d6", True, False), ("free", False, True)])
def test_root_base_classification_uses_dof_count(root_kind, expected_fixed, expected_floating):
    builder = ModelBuilder()
    root = builder.add_link(mass=1.0, inertia=I3, label="root")
    if root_kind == "fixed":
        root_joint = builder.add_joint_fixed(-1, root)
    elif root_kind == "locked_d6":
        root_joint = builder.add_joint_d6(-1, root)
    else:
        root_joint = builder.add_joint_free(parent=-1, child=root)
    builder.add_articulation([root_joint], label=root_kind)
    view = ArticulationView(builder.finalize(device="cpu"), root_kind)
    assert view.is_fixed_base == expected_fixed and view.is_floating_base == expected_floating


def test_labels_preserve_full_paths():
    builder = ModelBuilder()
    palm = builder.add_link(mass=1.0, inertia=I3, label="palm")
    left = builder.add_link(mass=1.0, inertia=I3, label="palm/left/fingertip")
    right = builder.add_link(mass=1.0, inertia=I3, label="palm/right/fingertip")
    builder.add_shape_box(left, hx=0.01, hy=0.01, hz=0.02, label="palm/left/tip_collision")
    builder.add_shape_box(right, hx=0.01, hy=0.01, hz=0.02, label="palm/right/tip_collision")
    j_root = builder.add_joint_free(parent=-1, child=palm, label="root")
    j_left = builder.add_joint_revolute(palm, left, axis=(0.0, 0.0, 1.0), label="palm/left/fingertip_joint")
    j_right = builder.add_joint_revolute(palm, right, axis=(0.0, 0.0, 1.0), label="palm/right/fingertip_joint")
    builder.add_articulation([j_root, j_left, j_right], label="gripper")
    view = ArticulationView(builder.finalize(), "gripper", include_links="fingertip")
    assert view.link_count == 2 and view.link_names == ["fingertip", "fingertip"] and view.shape_names == ["tip_collision", "tip_collision"]
    assert view.link_labels == ["palm/left/fingertip", "palm/right/fingertip"]
    assert view.shape_labels == ["palm/left/tip_collision", "palm/right/tip_collision"]
    assert "palm/left/fingertip_joint" in view.joint_labels and "palm/right/fingertip_joint" in view.joint_labels
    assert len(view.joint_labels) == view.joint_count and view.body_labels == view.link_labels


def test_duplicate_joint_child_is_one_link():
    """BODY-frequency link axis uses unique physical bodies, not joint slots."""
    builder = ModelBuilder()
    root = builder.add_link(mass=1.0, inertia=I3, label="root")
    tip = builder.add_link(mass=1.0, inertia=I3, label="tip")
    builder.add_shape_box(tip, hx=0.01, hy=0.01, hz=0.01, label="tip_shape")
    j_root = builder.add_joint_free(parent=-1, child=root, label="root_joint")
    j_tip = builder.add_joint_revolute(root, tip, axis=(0.0, 0.0, 1.0), label="tip_joint")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        j_dup = builder.add_joint_fixed(root, tip, label="tip_duplicate_joint")
    builder.add_articulation([j_root, j_tip, j_dup], label="robot")
    model = builder.finalize()
    view = ArticulationView(model, "robot")
    assert list(model.body_label) == ["root", "tip"]
    assert view.link_count == 2 and view.link_names == ["root", "tip"] and view.link_labels == ["root", "tip"]
    assert view.shape_count == 1 and view.shape_labels == ["tip_shape"]
    assert view.frequency_layouts[model.AttributeFrequency.BODY].value_count == len(model.body_label)
    assert view.get_link_transforms(model).shape == (1, 1, 2, 7) and view.get_link_velocities(model).shape == (1, 1, 2, 6)


# ---- run_test_joint_selection / run_test_link_selection (reference test_selection.py:615-1385) ---------------------------------
from newton_b200.utils import xform as X  # noqa: E402

# literal expectations of the reference test, keyed (use_mask, two articulations per view); 3 worlds x 2 articulations x 3 joints
EXPECTED_JOINT = {
    (True, True): ([0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 4.0, 0, 0, 0, 0, 0, 0], {11: -46.5}, {11: 8.0}),
    (True, False): ([0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 2.0, 0, 0, 0, 0, 0, 0], {11: -48.5}, {11: 4.0}),
    (False, True): ([0, 0, 1.0, 0, 0, 2.0, 0, 0, 3.0, 0, 0, 4.0, 0, 0, 5.0, 0, 0, 6.0],
                    {2: -49.5, 5: -48.5, 8: -47.5, 11: -46.5, 14: -45.5, 17: -44.5}, {2: 2.0, 5: 4.0, 8: 6.0, 11: 8.0, 14: 10.0, 17: 12.0}),
    (False, False): ([0, 0, 0, 0, 0, 1.0, 0, 0, 0, 0, 0, 2.0, 0, 0, 0, 0, 0, 3.0], {5: -49.5, 11: -48.5, 17: -47.5}, {5: 2.0, 11: 4.0, 17: 6.0}),
}


def _slider_world(two_per_view: bool):
    """The MJCF of the reference test, built from primitives: a root fixed to the world with three sliders along x."""
    art = ModelBuilder()
    root = art.add_link(mass=1.0, inertia=np.eye(3) * 0.01, label="myart/root")
    joints = [art.add_joint_fixed(-1, root, label="myart/root_fixed")]
    for k, y in enumerate((-0.5, -0.7, -0.9)):
        link = art.add_link(xform=X.transform((0.0, y, 0.0)), mass=1.0, inertia=np.eye(3) * 0.01, label=f"myart/link{k + 1}")
        joints.append(art.add_joint_prismatic(root, link, axis=(1.0, 0.0, 0.0), parent_xform=X.transform((0.0, y, 0.0)), limit_lower=-50.5,
                                              limit_upper=50.5, label=f"myart/joint{k + 1}"))
    art.add_articulation(joints, label="myart")
    world = ModelBuilder()
    world.add_builder(art)
    world.add_builder(art)
    world.articulation_label[1] = "art1"
    world.articulation_label[0] = "art1" if two_per_view else "art0"
    scene = ModelBuilder()
    for _ in range(3):
        scene.add_world(world)
    return scene.finalize()


@pytest.mark.parametrize("use_mask,two_per_view", list(EXPECTED_JOINT))
def test_joint_subset_writes_known_answers(host_copies, use_mask, two_per_view):
    import torch

    model = _slider_world(two_per_view)
    state, control = model.state(), model.control()
    view = ArticulationView(model, "art1", include_joints=["joint3"])
    q = view.get_dof_positions(model).numpy().copy()
    lower = view.get_attribute("joint_limit_lower", model).numpy().copy()
    target = view.get_attribute("joint_target_q", model).numpy().copy()
    assert q.shape == (3, 2 if two_per_view else 1, 1)
    val = 1.0
    for idx in np.ndindex(q.shape):
        q[idx] = val
        lower[idx] += val
        target[idx] += 2.0 * val
        val += 1.0
    mask = None
    if use_mask:
        mask = torch.tensor([[False, False], [False, True], [False, False]] if two_per_view else [[False], [True], [False]])
    view.set_dof_positions(state, q, mask)
    view.set_dof_positions(model, torch.from_numpy(q), mask)
    view.set_attribute("joint_limit_lower", model, lower, mask)
    view.set_attribute("joint_target_q", control, target, mask)
    view.set_attribute("joint_target_q", model, target.tolist(), mask)
    exp_q, exp_lower, exp_target = EXPECTED_JOINT[(use_mask, two_per_view)]
    full = lambda default, changes: [changes.get(i, default) for i in range(18)]  # noqa: E731
    np.testing.assert_allclose(state.joint_q.numpy(), exp_q, atol=1e-4)
    np.testing.assert_allclose(model.joint_q.numpy(), exp_q, atol=1e-4)
    np.testing.assert_allclose(model.joint_limit_lower.numpy(), full(-50.5, exp_lower), atol=1e-4)
    np.testing.assert_allclose(control.joint_target_q.numpy(), full(0.0, exp_target), atol=1e-4)
    np.testing.assert_allclose(model.joint_target_q.numpy(), full(0.0, exp_target), atol=1e-4)


@pytest.mark.parametrize("use_mask,two_per_view", list(EXPECTED_JOINT))
def test_link_subset_writes(host_copies, oracle_lib, use_mask, two_per_view):
    """Same scene, a link subset (reference run_test_link_selection :1011-1361): body_mass / body_q of "link2" only, against the
    stride-free walk."""
    import torch

    from oracle import selection as osel

    model = _slider_world(two_per_view)
    state = model.state()
    view = ArticulationView(model, "art1", include_links=["link2"])
    ids = osel.explicit_ids(model, "art1", exclude_links=["root", "link1", "link3"])
    rows = np.asarray(ids["link"], dtype=np.int64)
    A = 2 if two_per_view else 1
    assert rows.shape == (3, A, 1) and view.link_names == ["link2"]
    mask = None
    if use_mask:
        mask = torch.tensor([[False, False], [False, True], [False, False]] if two_per_view else [[False], [True], [False]])
    rng = np.random.default_rng(0)
    for name, target in (("body_mass", model), ("body_q", state), ("body_qd", state)):
        before = getattr(target, name).numpy().copy()
        values = rng.normal(size=osel.gather(before, rows).shape).astype(np.float32)
        expected = before.copy()
        osel.scatter_masked(expected, rows, values, None if mask is None else mask.numpy())
        view.set_attribute(name, target, values, mask)
        assert np.array_equal(getattr(target, name).numpy(), expected), name
        assert (expected != before).any()


# ---- newton/tests/test_match_labels.py ----------------------------------------------------------------------------------------
def test_match_labels_known_answers():
    from newton_b200.selection import match_labels

    assert match_labels(["alpha", "beta", "gamma"], "beta") == [1]
    assert match_labels(["arm_left", "arm_right", "leg_left"], "arm_*") == [0, 1]
    assert match_labels(["alpha", "beta", "gamma"], "delta") == []
    assert match_labels(["a", "b", "c"], "*") == [0, 1, 2]
    labels = ["/World/envs/env_0/Object_A", "/World/envs/env_12/Object_B", "/World/envs/env_12/Object_D", "/World/envs/env_x/Object_A"]
    assert match_labels(labels, re.compile(r"/World/envs/env_[0-9]+/Object_(A|B)")) == [0, 1]
    assert match_labels(["robot", "robot_arm"], re.compile(r"robot")) == [0]  # full match
    assert match_labels(labels[:2], r"/World/envs/env_[0-9]+/Object_(A|B)") == []  # a regex-looking string stays a glob
    assert match_labels(["alpha", "beta", "gamma", "delta"], ["alpha", "gamma"]) == [0, 2]
    assert match_labels(["a", "b", "c"], [2, 0]) == [2, 0]
    assert match_labels(["arm_left", "arm_right", "leg_left", "leg_right"], ["arm_*", "leg_left"]) == [0, 1, 2]
    assert match_labels(["a", "b", "c"], []) == []
    with pytest.raises(TypeError):
        match_labels(["a", "b"], [1.5])
    with pytest.raises(TypeError):
        match_labels(["a", "b"], [None])
    assert match_labels(["a", "b"], [99]) == [99]  # indices pass through unchecked
    assert match_labels(["arm_left", "arm_right", "leg_left", "leg_right"], ["arm_*", "*_left"]) == [0, 1, 2]  # no duplicates
