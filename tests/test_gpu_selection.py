"""ArticulationView on the GPU: the CUDA gather / masked-scatter / articulation-mask / masked-FK
kernels against the oracle (stride-free model walk + NumPy restatement of the reference kernels + C++ masked FK), bit for bit."""

import numpy as np
import pytest
import torch

import newton_b200
from newton_b200 import JointType, scenes
from newton_b200.selection import ArticulationView

pytestmark = pytest.mark.gpu

FREQ_KEY = {"joint_type": "joint", "joint_X_p": "joint", "joint_dof_dim": "joint", "joint_qd": "dof", "joint_limit_ke": "dof",
            "joint_axis": "dof", "joint_q": "coord", "body_q": "link", "body_qd": "link", "body_mass": "link", "body_inertia": "link",
            "shape_margin": "shape", "shape_scale": "shape", "shape_transform": "shape"}
VIEWS = [dict(), dict(exclude_joint_types=[int(JointType.FREE)]), dict(exclude_joints=["hip_2", "ankle_3"]),
         dict(exclude_links=["front_right_leg", "back_left_foot"])]


def _randomised(W, A, floating=True):
    model = scenes.ants_model(W, A, floating=floating)
    rng = np.random.default_rng(11)
    for name in ("joint_q", "joint_qd", "body_q", "body_qd", "joint_limit_ke", "shape_margin", "body_mass"):
        t = getattr(model, name)
        t.copy_(torch.from_numpy(rng.normal(size=tuple(t.shape)).astype(np.float32)))
    return model, rng


@pytest.mark.parametrize("kwargs", VIEWS)
def test_gather_and_masked_scatter_bit_exact(cuda_lib, oracle_lib, kwargs):
    from oracle import selection as osel

    W, A = 6, 3
    cpu, rng = _randomised(W, A)
    gpu = cpu.to("cuda:0")
    view = ArticulationView(gpu, "ant", **kwargs)
    ids = osel.explicit_ids(cpu, "ant", **kwargs)
    launches0 = newton_b200._lib.kernel_launch_count()
    for name, key in FREQ_KEY.items():
        attrib = cpu.numpy(name)
        expected = osel.take(attrib, ids[key])
        got = view.get_attribute(name, gpu)
        assert tuple(got.shape) == expected.shape, name
        assert np.array_equal(got.cpu().numpy().view(np.uint32), expected.view(np.uint32)), name
        rows = np.asarray([[r for r in world] for world in ids[key]], dtype=np.int64)
        values = (rng.normal(size=expected.shape).astype(np.float32) if attrib.dtype == np.float32
                  else rng.integers(0, 100, size=expected.shape).astype(attrib.dtype))
        for mask in (None, rng.random(W) < 0.5, rng.random((W, A)) < 0.5, np.zeros(W, dtype=bool)):
            ref = attrib.copy()
            osel.scatter_masked(ref, rows, values, mask)
            target = cpu.to("cuda:0")
            tview = ArticulationView(target, "ant", **kwargs)
            tmask = None if mask is None else torch.from_numpy(mask).to("cuda:0")
            tview.set_attribute(name, target, torch.from_numpy(values).to("cuda:0"), mask=tmask)
            assert np.array_equal(target.numpy(name).view(np.uint32), ref.view(np.uint32)), (name, None if mask is None else mask.shape)
    torch.cuda.synchronize()
    assert newton_b200._lib.kernel_launch_count() > launches0  # the copies ran in this library's kernels


def test_values_as_lists_and_numpy_and_views(cuda_lib, oracle_lib):
    cpu, rng = _randomised(3, 2)
    gpu = cpu.to("cuda:0")
    view = ArticulationView(gpu, "ant")
    state = gpu.state()
    q = view.get_dof_positions(state)
    assert q.data_ptr() == state.joint_q.data_ptr()  # contiguous selection: a view, like the reference
    q_np = q.cpu().numpy().copy()
    q_np[..., 7] = 0.25
    view.set_dof_positions(state, q_np)  # NumPy input
    assert torch.all(view.get_dof_positions(state)[..., 7] == 0.25)
    view.set_dof_velocities(state, np.zeros((3, 2, 14), dtype=np.float32).tolist(), mask=[True, False, True])  # nested lists
    qd = view.get_dof_velocities(state).cpu().numpy()
    assert not qd[0].any() and not qd[2].any() and qd[1].any()
    view.set_dof_positions(state, q)  # the view itself: in place, no launch
    with pytest.raises(ValueError):
        view.set_dof_positions(state, q_np[:2])
    with pytest.raises(ValueError):
        view.set_dof_positions(state, q_np, mask=torch.ones(3, dtype=torch.bool))  # mask on the wrong device
    ctrl = gpu.control()
    view.set_dof_forces(ctrl, torch.full((3, 2, 14), 2.0, device="cuda:0"), mask=torch.tensor([[1, 0], [0, 0], [0, 1]], dtype=torch.bool, device="cuda:0"))
    f = view.get_dof_forces(ctrl).cpu().numpy()
    assert (f[0, 0] == 2).all() and (f[2, 1] == 2).all() and not f[0, 1].any() and not f[1].any() and not f[2, 0].any()


def test_articulation_mask_known_answers(cuda_lib):
    """Reference test_selection.py:533-582."""
    model = scenes.ants_model(4, 3, ground=False, device="cuda:0")
    view = ArticulationView(model, "ant")
    assert view.get_model_articulation_mask().cpu().tolist() == [True] * 12
    expected = [bool(x) for x in [0, 0, 0, 1, 1, 1, 1, 1, 1, 0, 0, 0]]
    assert view.get_model_articulation_mask(mask=[0, 1, 1, 0]).cpu().tolist() == expected
    assert view.get_model_articulation_mask(mask=torch.tensor([0, 1, 1, 0], dtype=torch.bool, device="cuda:0")).cpu().tolist() == expected
    m = [[0, 1, 0], [1, 0, 1], [1, 1, 1], [0, 0, 0]]
    expected = [bool(x) for x in [0, 1, 0, 1, 0, 1, 1, 1, 1, 0, 0, 0]]
    assert view.get_model_articulation_mask(mask=m).cpu().tolist() == expected
    assert view.get_model_articulation_mask(mask=torch.tensor(m, dtype=torch.bool, device="cuda:0")).cpu().tolist() == expected
    sub = ArticulationView(model, [1, 4, 7, 10])  # the middle ant of every world
    assert sub.get_model_articulation_mask(mask=[1, 0, 0, 1]).cpu().tolist() == [i in (1, 10) for i in range(12)]


@pytest.mark.parametrize("floating", [True, False])
def test_masked_eval_fk_bit_exact(cuda_lib, oracle_lib, floating):
    import oracle

    cpu, rng = _randomised(5, 2, floating)
    cpu.joint_q.view(10, -1)[:, 3:7] = torch.nn.functional.normalize(cpu.joint_q.view(10, -1)[:, 3:7], dim=1) if floating else cpu.joint_q.view(10, -1)[:, 3:7]
    gpu = cpu.to("cuda:0")
    view = ArticulationView(gpu, "ant")
    for mask in (None, np.array([1, 0, 1, 1, 0], dtype=bool), rng.random((5, 2)) < 0.5):
        ref, out = cpu.state(), gpu.state()
        ref.body_q[:] = -99.0
        ref.body_qd[:] = -77.0
        out.body_q[:] = -99.0
        out.body_qd[:] = -77.0
        model_mask = np.ones(10, dtype=bool) if mask is None else (np.repeat(mask, 2) if mask.ndim == 1 else mask.reshape(-1))
        oracle.eval_fk(cpu, ref.joint_q, ref.joint_qd, ref, mask=model_mask)
        view.eval_fk(out, mask=None if mask is None else torch.from_numpy(mask).to("cuda:0"))
        assert np.array_equal(out.body_q.cpu().numpy().view(np.uint32), ref.body_q.numpy().view(np.uint32))
        assert np.array_equal(out.body_qd.cpu().numpy().view(np.uint32), ref.body_qd.numpy().view(np.uint32))
        assert (ref.body_q.numpy().reshape(10, 9, 7)[~model_mask] == -99.0).all()
    # index list, with out-of-range entries (ignored)
    ref, out = cpu.state(), gpu.state()
    ref.body_q[:] = -99.0
    out.body_q[:] = -99.0
    oracle.eval_fk(cpu, ref.joint_q, ref.joint_qd, ref, indices=[7, 2, 40, -3])
    newton_b200.eval_fk(gpu, out.joint_q, out.joint_qd, out, indices=[7, 2, 40, -3])
    assert np.array_equal(out.body_q.cpu().numpy().view(np.uint32), ref.body_q.numpy().view(np.uint32))
    with pytest.raises(ValueError):
        newton_b200.eval_fk(gpu, out.joint_q, out.joint_qd, out, mask=view.articulation_mask, indices=[0])


def test_rl_reset_loop_matches_oracle(cuda_lib, oracle_lib):
    """The loop the row exists for: simulate, reset the 'done' worlds to their initial state through the view (root pose, root
    velocity, joint angles and rates, then FK under the same mask), keep simulating.  The CUDA run must equal the oracle run that
    applies the same reset with NumPy indexing, bit for bit, and untouched worlds must not notice the reset."""
    import oracle
    from oracle import selection as osel

    W = 6
    cpu = scenes.quadruped_model(W, seed=3)
    gpu = cpu.to("cuda:0")
    view = ArticulationView(gpu, "quadruped")
    ids = osel.explicit_ids(cpu, "quadruped")
    done = np.array([0, 1, 0, 0, 1, 1], dtype=bool)
    kw = {"iterations": 2}
    dt = 0.005

    def run(model, pipeline_cls, solver_cls, fk, reset):
        solver, pipe = solver_cls(model, **kw), pipeline_cls(model)
        s0, s1, ctrl, contacts = model.state(), model.state(), model.control(), pipe.contacts()
        for phase in range(2):
            for _ in range(12):
                s0.clear_forces()
                pipe.collide(s0, contacts)
                solver.step(s0, s1, ctrl, contacts, dt)
                s0, s1 = s1, s0
            if phase == 0:
                mid = s0.body_q.clone()
                fk(model, s0)  # XPBD does not maintain joint_q / joint_qd: recover them first (eval_ik), as an RL loop does
                reset(model, s0)
                solver.reset(s0, world_mask=torch.cat([torch.from_numpy(done), torch.tensor([False])]).to(model.device))
        return s0, mid

    def reset_gpu(model, s):
        m = torch.from_numpy(done).to("cuda:0")
        view.set_root_transforms(s, view.get_root_transforms(model), mask=m)
        view.set_root_velocities(s, view.get_root_velocities(model), mask=m)
        view.set_dof_positions(s, view.get_dof_positions(model), mask=m)
        view.set_dof_velocities(s, view.get_dof_velocities(model), mask=m)
        view.eval_fk(s, mask=m)

    def reset_cpu(model, s):
        for name, key in (("joint_q", "coord"), ("joint_qd", "dof")):
            rows = np.asarray(ids[key], dtype=np.int64)
            arr = getattr(s, name).numpy()
            osel.scatter_masked(arr, rows, osel.gather(model.numpy(name), rows), done)
        oracle.eval_fk(model, s.joint_q, s.joint_qd, s, mask=osel.model_articulation_mask(np.asarray(ids["articulation"]), W, done))

    ref, ref_mid = run(cpu, oracle.CollisionPipeline, oracle.SolverXPBD, lambda m, s: oracle.eval_ik(m, s, s.joint_q, s.joint_qd), reset_cpu)
    out, out_mid = run(gpu, newton_b200.CollisionPipeline, newton_b200.solvers.SolverXPBD,
                       lambda m, s: newton_b200.eval_ik(m, s, s.joint_q, s.joint_qd), reset_gpu)
    assert np.array_equal(out_mid.cpu().numpy(), ref_mid.numpy())
    for name in ("body_q", "body_qd", "joint_q", "joint_qd"):
        assert np.array_equal(getattr(out, name).cpu().numpy().view(np.uint32), getattr(ref, name).numpy().view(np.uint32)), name
    # worlds that were not reset equal an uninterrupted 24-substep run
    solver, pipe = newton_b200.solvers.SolverXPBD(gpu, **kw), newton_b200.CollisionPipeline(gpu)
    s0, s1, ctrl, contacts = gpu.state(), gpu.state(), gpu.control(), pipe.contacts()
    for _ in range(24):
        s0.clear_forces()
        pipe.collide(s0, contacts)
        solver.step(s0, s1, ctrl, contacts, dt)
        s0, s1 = s1, s0
    a, b = out.body_q.view(W, 13, 7).cpu().numpy(), s0.body_q.view(W, 13, 7).cpu().numpy()
    assert np.array_equal(a[~done], b[~done]) and not np.array_equal(a[done], b[done])


def test_reset_mask_validation_on_solver(cuda_lib):
    model = scenes.quadruped_model(3, device="cuda:0")
    solver = newton_b200.solvers.SolverXPBD(model)
    state = model.state()
    solver.reset(state)
    solver.reset(state, world_mask=torch.ones(4, dtype=torch.bool, device="cuda:0"))
    with pytest.warns(DeprecationWarning):
        solver.reset(state, world_mask=torch.ones(3, dtype=torch.bool, device="cuda:0"))
    with pytest.raises(ValueError):
        solver.reset(state, world_mask=torch.ones(7, dtype=torch.bool, device="cuda:0"))
    with pytest.raises(ValueError):
        solver.reset(state, world_mask=torch.ones(4, dtype=torch.bool))  # wrong device


def test_eval_fk_child_before_parent_order_uses_serial_walk(cuda_lib, oracle_lib):
    """Joints stored child-before-parent: the reference's serial walk then reads the parent's OLD pose for the first joint.  The
    level-parallel kernel cannot reproduce that, so model creation must fall back to the serial kernel (nb2_model::fk_levels)."""
    import oracle
    from newton_b200 import ModelBuilder
    from newton_b200.utils import xform as X

    b = ModelBuilder()
    inertia = np.eye(3) * 0.1
    for _ in range(3):  # three chains so that several threads / warps run
        b1 = b.add_link(mass=1.0, inertia=inertia, xform=X.transform((0.3, 0.1, 0.2)))
        b2 = b.add_link(mass=1.0, inertia=inertia)
        j_child = b.add_joint_revolute(b1, b2, axis=(0, 1, 0), parent_xform=X.transform((0.0, 0.0, 0.5)))
        j_root = b.add_joint_revolute(-1, b1, axis=(0, 0, 1), parent_xform=X.transform((1.0, 0.0, 0.0)))
        b.add_articulation([j_child, j_root])
    cpu = b.finalize()
    cpu.joint_q.copy_(torch.tensor([0.3, -0.7, 0.1, 0.9, -0.4, 0.5]))
    cpu.joint_qd.copy_(torch.tensor([1.0, 2.0, -1.0, 0.5, 0.25, -2.0]))
    gpu = cpu.to("cuda:0")
    ref, out = cpu.state(), gpu.state()
    oracle.eval_fk(cpu, ref.joint_q, ref.joint_qd, ref)
    newton_b200.eval_fk(gpu, out.joint_q, out.joint_qd, out)
    assert np.array_equal(out.body_q.cpu().numpy().view(np.uint32), ref.body_q.numpy().view(np.uint32))
    assert np.array_equal(out.body_qd.cpu().numpy().view(np.uint32), ref.body_qd.numpy().view(np.uint32))
    # ... and that really is the stale-parent result: a second pass (parents now up to date) changes the children
    again = cpu.state()
    again.body_q.copy_(ref.body_q)
    again.body_qd.copy_(ref.body_qd)
    oracle.eval_fk(cpu, again.joint_q, again.joint_qd, again)
    assert not np.array_equal(again.body_q.numpy(), ref.body_q.numpy())
