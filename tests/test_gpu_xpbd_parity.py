"""GPU parity: CUDA collide + XPBD step vs the CPU oracle on identical Model/State inputs, through the C-ABI.

Bar (BASELINE.json north_star): contact shape ids / counts bit-exact; body_q / body_qd within 1e-5 relative after
100 substeps.  The product library is built strict-fp (no FMA contraction, correctly rounded trig), so the bar
enforced here is stronger: *bit-identical* state and contact arrays after 100 substeps.  The FMA-contracted twin
library is checked against the stated tolerances in test_fast_fp_library_within_tolerance.
"""

import os
import subprocess
import sys

import numpy as np
import pytest
import torch

import newton_b200
from newton_b200 import scenes
from newton_b200.sim.builder import ModelBuilder
from newton_b200.utils import xform as X
from tests.helpers import canonical_contacts, rel_err, simulate

pytestmark = pytest.mark.gpu

TOL_STATE_REL = 1e-5  # north_star tolerance (used for the fast-fp twin; the product library must be exact)


def _both(model_cpu, substeps, dt, solver_kwargs, oracle, control_fn=None, **sim_kw):
    ctrl_cpu = control_fn(model_cpu) if control_fn else None
    ref_state, ref_contacts, ref_counts = simulate(model_cpu, oracle.CollisionPipeline, oracle.SolverXPBD, substeps=substeps,
                                                   dt=dt, solver_kwargs=solver_kwargs, record_contacts=True, control=ctrl_cpu,
                                                   **sim_kw)
    model_gpu = model_cpu.to("cuda:0")
    ctrl_gpu = control_fn(model_gpu) if control_fn else None
    gpu_state, gpu_contacts, gpu_counts = simulate(model_gpu, newton_b200.CollisionPipeline, newton_b200.solvers.SolverXPBD,
                                                   substeps=substeps, dt=dt, solver_kwargs=solver_kwargs,
                                                   record_contacts=True, control=ctrl_gpu, **sim_kw)
    torch.cuda.synchronize()
    return ref_state, ref_contacts, ref_counts, gpu_state, gpu_contacts, gpu_counts


def _assert_exact(rs, rc, rcounts, gs, gc, gcounts, model, need_contacts=True):
    assert rcounts == gcounts, "per-substep rigid contact counts must be bit-exact"
    n_ref, cr = canonical_contacts(rc, model)
    n_gpu, cg = canonical_contacts(gc, model)
    assert n_ref == n_gpu
    if need_contacts:
        assert n_ref > 0
    for k in cr:
        np.testing.assert_array_equal(cr[k], cg[k], err_msg=f"contact field {k}")
    np.testing.assert_array_equal(gs.body_q.cpu().numpy(), rs.body_q.numpy())
    np.testing.assert_array_equal(gs.body_qd.cpu().numpy(), rs.body_qd.numpy())


def _drop(model, worlds, z):
    model.joint_q.view(worlds, -1)[:, 2] = z
    scenes.host_fk(model, model.joint_q, model.joint_qd, model)
    return model


@pytest.mark.parametrize("world_count,iterations", [(1, 2), (8, 8), (33, 4)])
def test_quadruped_100_substeps_bit_exact(oracle_lib, cuda_lib, world_count, iterations):
    """Config 3 scene (reduced env count): 100 x (collide, XPBD step) with ground contacts and 12 driven joints/env."""
    model = _drop(scenes.quadruped_model(world_count, seed=1), world_count, 0.48)
    out = _both(model, 100, 1.0 / 50 / 4, {"iterations": iterations}, oracle_lib)
    _assert_exact(*out, model)


def test_quadruped_joint_forces_and_targets(oracle_lib, cuda_lib):
    """Non-zero Control.joint_f (apply_joint_forces path) and moving position targets."""

    def ctrl(model):
        c = model.control()
        g = torch.Generator().manual_seed(5)
        c.joint_f.copy_((torch.rand(c.joint_f.shape, generator=g) * 4.0 - 2.0).to(c.joint_f.device))
        c.joint_target_q.add_((torch.rand(c.joint_target_q.shape, generator=g) * 0.2 - 0.1).to(c.joint_f.device))
        return c

    model = _drop(scenes.quadruped_model(6, seed=2), 6, 0.5)
    out = _both(model, 60, 0.005, {"iterations": 4, "angular_damping": 0.05}, oracle_lib, control_fn=ctrl)
    _assert_exact(*out, model)


def test_shapes_on_plane_bit_exact(oracle_lib, cuda_lib):
    """Sphere / capsule / cylinder / box against the plane: every analytic plane collider + free bodies."""
    model = scenes.shapes_on_plane_model(5, seed=3)
    out = _both(model, 120, 1.0 / 240, {"iterations": 3}, oracle_lib)
    _assert_exact(*out, model)


@pytest.mark.parametrize("world_count", [1, 7])
def test_box_stacks_bit_exact(oracle_lib, cuda_lib, world_count):
    """Config 2 scene (reduced stack count): box-box pairs through MPR + 4-point manifolds, plane-box analytic."""
    model = scenes.box_stack_model(world_count, seed=0)
    out = _both(model, 100, 1.0 / 60 / 4, {"iterations": 4}, oracle_lib)
    _assert_exact(*out, model)
    assert out[2][-1] == 20 * world_count  # 4 contacts on each of the 5 interfaces of every stack


def test_convex_pile_bit_exact(oracle_lib, cuda_lib):
    """Every pair class of the generic convex path (box-box, capsule-box, cylinder-cylinder, cylinder-box,
    ellipsoid-box/sphere/ellipsoid) tumbling for 150 substeps: MPR, GJK fallback, manifold clipping, axial roll."""
    model = scenes.convex_pile_model(4, seed=5)
    out = _both(model, 150, 1.0 / 240, {"iterations": 4}, oracle_lib)
    _assert_exact(*out, model)


def _bouncy_scene(worlds):
    """Spheres, a box and a capsule with restitution dropped on each other and on the plane."""
    from newton_b200.sim.builder import ShapeConfig

    rng = np.random.default_rng(11)
    scene = ModelBuilder()
    ground = ShapeConfig(restitution=0.6, mu=0.4)
    for _ in range(worlds):
        scene.begin_world()
        for i, e in enumerate((0.9, 0.5, 0.2)):
            cfg = ShapeConfig(restitution=e, mu=0.3)
            b = scene.add_body(xform=X.transform(np.array([0.6 * i, 0.0, 0.45 + 0.2 * i]) + rng.uniform(-0.02, 0.02, 3)))
            scene.add_shape_sphere(b, radius=0.2, cfg=cfg)
        b = scene.add_body(xform=X.transform(np.array([0.3, 0.05, 1.2]) + rng.uniform(-0.02, 0.02, 3),
                                             X.quat_from_axis_angle((1.0, 0.0, 0.0), 0.4)))
        scene.add_shape_box(b, hx=0.25, hy=0.2, hz=0.1, cfg=ShapeConfig(restitution=0.7, mu=0.5))
        b = scene.add_body(xform=X.transform(np.array([-0.7, 0.0, 0.5]), X.quat_from_axis_angle((0.0, 1.0, 0.0), 1.0)))
        scene.add_shape_capsule(b, radius=0.15, half_height=0.3, cfg=ShapeConfig(restitution=0.8, mu=0.5))
        scene.end_world()
    scene.add_ground_plane(cfg=ground)
    return scene.finalize()


def test_restitution_bit_exact(oracle_lib, cuda_lib):
    """enable_restitution=True (apply_rigid_restitution + apply_body_delta_velocities, kernels.py:2582-2728, 936-942)."""
    model = _bouncy_scene(6)
    out = _both(model, 200, 1.0 / 400, {"iterations": 4, "enable_restitution": True}, oracle_lib)
    _assert_exact(*out, model)
    assert float(out[3].body_qd.abs().max()) > 0.1  # things are still moving: the restitution pass was exercised


def test_velocity_from_position_delta_bit_exact(oracle_lib, cuda_lib):
    """SolverXPBD.compute_body_velocity_from_position_delta (update_body_velocities, kernels.py:2547-2579)."""
    model = _drop(scenes.quadruped_model(5, seed=7), 5, 0.5)
    out = _both(model, 60, 0.005, {"iterations": 4}, oracle_lib, solver_attrs={"compute_body_velocity_from_position_delta": True})
    _assert_exact(*out, model)


def test_contact_force_and_parent_force_bit_exact(oracle_lib, cuda_lib):
    """Contacts.force via update_contacts() and State.body_parent_f (row a17 reporting kernels), quadrupeds on the ground
    with non-zero joint_f, plus box stacks (body-body contacts, two dynamic bodies per contact)."""

    def ctrl(model):
        c = model.control()
        g = torch.Generator().manual_seed(3)
        c.joint_f.copy_((torch.rand(c.joint_f.shape, generator=g) * 2.0 - 1.0).to(c.joint_f.device))
        return c

    for model, control_fn in ((_drop(scenes.quadruped_model(4, seed=2), 4, 0.47), ctrl), (scenes.box_stack_model(3, seed=1), None)):
        model.request_contact_attributes("force")
        model.request_state_attributes("body_parent_f")
        out = _both(model, 50, 0.005, {"iterations": 6}, oracle_lib, control_fn=control_fn, update_contacts=True)
        _assert_exact(*out, model)
        rs, gs = out[0], out[3]
        np.testing.assert_array_equal(gs.body_parent_f.cpu().numpy(), rs.body_parent_f.numpy())
        assert np.abs(canonical_contacts(out[1], model)[1]["force"]).max() > 1.0


def test_heterogeneous_worlds_bit_exact(oracle_lib, cuda_lib):
    """Worlds of different sizes in one model (quadruped / box stack / empty / single link): partial lane groups."""
    model = scenes.mixed_worlds_model(3)
    out = _both(model, 100, 1.0 / 240, {"iterations": 4}, oracle_lib)
    _assert_exact(*out, model)


def test_implicit_single_world_and_no_contacts(oracle_lib, cuda_lib):
    """Model built without begin_world() (all entities in world -1) and step(contacts=None)."""
    b = ModelBuilder()
    for i in range(3):
        body = b.add_body(xform=X.transform((0.3 * i, 0.0, 0.4 + 0.5 * i), X.quat_from_axis_angle((1, 0, 0), 0.3 * i)))
        b.add_shape_capsule(body, radius=0.1, half_height=0.2)
    b.add_ground_plane()
    model = b.finalize()
    out = _both(model, 80, 1.0 / 200, {"iterations": 2}, oracle_lib)
    _assert_exact(*out, model)
    # contacts=None path
    rs, _, _ = simulate(model, None, oracle_lib.SolverXPBD, substeps=20, dt=0.01, collide=False)
    gs, _, _ = simulate(model.to("cuda:0"), None, newton_b200.solvers.SolverXPBD, substeps=20, dt=0.01, collide=False)
    np.testing.assert_array_equal(gs.body_q.cpu().numpy(), rs.body_q.numpy())
    np.testing.assert_array_equal(gs.body_qd.cpu().numpy(), rs.body_qd.numpy())


def test_integrate_bodies_kernel(oracle_lib, cuda_lib):
    model = scenes.quadruped_model(3, seed=4)
    s_in = model.state()
    g = torch.Generator().manual_seed(1)
    s_in.body_qd.copy_(torch.rand(s_in.body_qd.shape, generator=g) - 0.5)
    s_in.body_f.copy_(torch.rand(s_in.body_f.shape, generator=g) * 10 - 5)
    s_out = model.state()
    oracle_lib.SolverXPBD(model).integrate_bodies(model, s_in, s_out, 0.01, angular_damping=0.1)
    mg = model.to("cuda:0")
    g_in, g_out = mg.state(), mg.state()
    g_in.body_qd.copy_(s_in.body_qd)
    g_in.body_f.copy_(s_in.body_f)
    newton_b200.solvers.SolverXPBD(mg).integrate_bodies(mg, g_in, g_out, 0.01, angular_damping=0.1)
    np.testing.assert_array_equal(g_out.body_q.cpu().numpy(), s_out.body_q.numpy())
    np.testing.assert_array_equal(g_out.body_qd.cpu().numpy(), s_out.body_qd.numpy())


def test_single_collide_contacts_match(oracle_lib, cuda_lib):
    """One collide() on a settled pose: every exported contact field must agree exactly (as sets, canonical order)."""
    model = _drop(scenes.quadruped_model(4, seed=1), 4, 0.45)
    pipe = oracle_lib.CollisionPipeline(model)
    c_ref = pipe.contacts()
    pipe.collide(model.state(), c_ref)
    mg = model.to("cuda:0")
    pg = newton_b200.CollisionPipeline(mg)
    c_gpu = pg.contacts()
    pg.collide(mg.state(), c_gpu)
    n_ref, cr = canonical_contacts(c_ref, model)
    n_gpu, cg = canonical_contacts(c_gpu, mg)
    assert n_ref == n_gpu and n_ref > 0
    for k in cr:
        np.testing.assert_array_equal(cr[k], cg[k], err_msg=k)


def test_cuda_graph_capture_matches_eager(cuda_lib):
    """The step must be stream-capturable like the reference (example_basic_urdf.py:112-115): no sync / alloc inside."""
    model = _drop(scenes.quadruped_model(16, seed=1), 16, 0.48).to("cuda:0")
    pipe = newton_b200.CollisionPipeline(model)
    solver = newton_b200.solvers.SolverXPBD(model, iterations=4)

    def run(graph_mode):
        s0, s1, ctrl, contacts = model.state(), model.state(), model.control(), pipe.contacts()

        def frame():
            nonlocal s0, s1
            for _ in range(2):
                s0.clear_forces()
                pipe.collide(s0, contacts)
                solver.step(s0, s1, ctrl, contacts, 0.005)
                s0, s1 = s1, s0

        if graph_mode:
            stream = torch.cuda.Stream()
            with torch.cuda.stream(stream):
                frame()  # warm-up outside capture
            torch.cuda.synchronize()
            s0.assign(model.state()), s1.assign(model.state())
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=stream):
                frame()
            s0.assign(model.state()), s1.assign(model.state())
            for _ in range(10):
                g.replay()
        else:
            for _ in range(10):
                frame()
        torch.cuda.synchronize()
        return s0.body_q.clone(), s0.body_qd.clone()

    q_e, qd_e = run(False)
    q_g, qd_g = run(True)
    assert torch.equal(q_e, q_g) and torch.equal(qd_e, qd_g)


def test_fast_fp_library_within_tolerance(oracle_lib):
    """The FMA-contracted twin: positions within the north-star 1e-5; velocities sit on the algorithm's fp32 noise floor
    (XPBD derives velocity from position deltas / dt, so 1-ulp position differences become ~ulp/dt in velocity), checked
    at 2e-3 relative."""
    code = r"""
import sys, numpy as np, torch
sys.path.insert(0, %r)
import newton_b200, oracle
from newton_b200 import scenes
from tests.helpers import simulate, rel_err
m = scenes.quadruped_model(8, seed=1)
m.joint_q.view(8, -1)[:, 2] = 0.48
scenes.host_fk(m, m.joint_q, m.joint_qd, m)
kw = {"iterations": 8}
rs, _, rc = simulate(m, oracle.CollisionPipeline, oracle.SolverXPBD, substeps=100, dt=0.005, solver_kwargs=kw, record_contacts=True)
gs, _, gc = simulate(m.to("cuda:0"), newton_b200.CollisionPipeline, newton_b200.solvers.SolverXPBD, substeps=100, dt=0.005, solver_kwargs=kw, record_contacts=True)
assert rc == gc, "contact counts"
print("ERR", rel_err(gs.body_q.cpu().numpy(), rs.body_q.numpy()), rel_err(gs.body_qd.cpu().numpy(), rs.body_qd.numpy()))
""" % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ)
    env["NB2_LIB"] = os.path.join(os.path.dirname(newton_b200.__file__), "libnewton_b200_fast.so")
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    eq, eqd = [float(x) for x in out.stdout.strip().split("ERR")[1].split()]
    assert eq < TOL_STATE_REL
    assert eqd < 2e-3


@pytest.mark.parametrize("solver_name", ["xpbd", "featherstone"])
def test_full_size_batch_matches_single_env_oracle(oracle_lib, cuda_lib, solver_name):
    """BASELINE.json's full batch (4096 quadruped envs on one GPU) through a size-independent property: environments never
    interact, so 4096 copies of one environment must each reproduce, bit for bit, what the oracle computes for that single
    environment - which checks grid sizing, slot offsets and indexing at scale at the cost of a 1-env oracle run."""
    envs = 4096
    if solver_name == "xpbd":
        kw, dt, n, pkg_o, pkg_g = {"iterations": 8}, 0.005, 24, oracle_lib.SolverXPBD, newton_b200.solvers.SolverXPBD
    else:
        kw, dt, n, pkg_o, pkg_g = {}, 0.001, 40, oracle_lib.SolverFeatherstone, newton_b200.solvers.SolverFeatherstone
    one = _drop(scenes.quadruped_model(1, seed=None), 1, 0.47)
    ref, _, rc = simulate(one, oracle_lib.CollisionPipeline, pkg_o, substeps=n, dt=dt, solver_kwargs=kw, record_contacts=True)
    big = _drop(scenes.quadruped_model(envs, seed=None), envs, 0.47).to("cuda:0")
    out, _, gc = simulate(big, newton_b200.CollisionPipeline, pkg_g, substeps=n, dt=dt, solver_kwargs=kw, record_contacts=True)
    assert gc == [c * envs for c in rc] and rc[-1] > 0
    q = out.body_q.cpu().numpy().reshape(envs, -1, 7)
    qd = out.body_qd.cpu().numpy().reshape(envs, -1, 6)
    np.testing.assert_array_equal(q, np.broadcast_to(ref.body_q.numpy()[None], q.shape))
    np.testing.assert_array_equal(qd, np.broadcast_to(ref.body_qd.numpy()[None], qd.shape))


def test_full_size_box_stacks_match_single_env_oracle(oracle_lib, cuda_lib):
    """Config 2 at its full size (512 stacks): every stack equals the oracle's single stack, 20 manifold contacts each."""
    envs = 512
    kw = {"iterations": 8}
    ref, _, rc = simulate(scenes.box_stack_model(1, seed=None), oracle_lib.CollisionPipeline, oracle_lib.SolverXPBD, substeps=24,
                          dt=1.0 / 240, solver_kwargs=kw, record_contacts=True)
    out, _, gc = simulate(scenes.box_stack_model(envs, seed=None).to("cuda:0"), newton_b200.CollisionPipeline,
                          newton_b200.solvers.SolverXPBD, substeps=24, dt=1.0 / 240, solver_kwargs=kw, record_contacts=True)
    assert gc == [c * envs for c in rc] and rc[-1] == 20
    q = out.body_q.cpu().numpy().reshape(envs, -1, 7)
    np.testing.assert_array_equal(q, np.broadcast_to(ref.body_q.numpy()[None], q.shape))


def test_ramp_scene_bit_exact(oracle_lib, cuda_lib):
    """The reference's GJK/MPR multi-contact ramp scene (test_rigid_contact.py:236-432; see tests/test_oracle_known_answers.py):
    dynamic cubes / capsule / cylinder against static box walls and a tilted plane, implicit single world."""
    from tests.test_oracle_known_answers import _ramp_scene

    model, _ = _ramp_scene()
    out = _both(model, 200, 1.0 / 600, {"iterations": 2}, oracle_lib)
    _assert_exact(*out, model)


@pytest.mark.parametrize("solver_name", ["xpbd", "featherstone"])
def test_run_to_run_determinism(cuda_lib, solver_name):
    """newton/tests/determinism/test_solver_determinism.py:205-229: two runs from the same state are bit-equal (ordered sums
    instead of float atomics make this hold by construction, on any batch size)."""
    cls = newton_b200.solvers.SolverXPBD if solver_name == "xpbd" else newton_b200.solvers.SolverFeatherstone
    kw = {"iterations": 4} if solver_name == "xpbd" else {}
    model = _drop(scenes.quadruped_model(257, seed=4), 257, 0.47).to("cuda:0")
    runs = [simulate(model, newton_b200.CollisionPipeline, cls, substeps=30, dt=0.002, solver_kwargs=kw, record_contacts=True)
            for _ in range(2)]
    assert runs[0][2] == runs[1][2]
    for name in ("body_q", "body_qd"):
        assert torch.equal(getattr(runs[0][0], name), getattr(runs[1][0], name)), name


@pytest.mark.parametrize("solver_name", ["xpbd", "featherstone"])
def test_foreign_contacts_import(oracle_lib, cuda_lib, solver_name):
    """nb2_contacts_import: a reference-layout Contacts buffer that newton_b200's collide did not produce.  (a) A plain copy
    of the exported arrays must step exactly like the native contact blocks; (b) the same contacts in the order the ORACLE
    pipeline emits them (global sort-key order: all box-box contacts before the plane contacts) must reproduce the oracle's
    step bit for bit - the per-world stable sort keeps each body's summation order."""
    model = scenes.mixed_worlds_model(2)
    mg = model.to("cuda:0")
    if solver_name == "xpbd":
        mk_g = lambda: newton_b200.solvers.SolverXPBD(mg, iterations=4)  # noqa: E731
        mk_o = lambda: oracle_lib.SolverXPBD(model, iterations=4)  # noqa: E731
        dt = 1.0 / 240
    else:
        mk_g = lambda: newton_b200.solvers.SolverFeatherstone(mg)  # noqa: E731
        mk_o = lambda: oracle_lib.SolverFeatherstone(model)  # noqa: E731
        dt = 1.0 / 480
    pipe = newton_b200.CollisionPipeline(mg)
    native = pipe.contacts()
    s_in = mg.state()
    pipe.collide(s_in, native)
    assert int(native.rigid_contact_count.item()) > 20
    # (a) copy of the exported buffer, no native tag
    copy = pipe.contacts()
    for name, value in vars(native).items():
        if isinstance(value, torch.Tensor):
            getattr(copy, name).copy_(value)
    out_native, out_copy = mg.state(), mg.state()
    mk_g().step(mg.state(), out_native, None, native, dt)
    mk_g().step(mg.state(), out_copy, None, copy, dt)
    assert torch.equal(out_native.body_q, out_copy.body_q) and torch.equal(out_native.body_qd, out_copy.body_qd)
    # (b) the oracle pipeline's buffer (different global order), moved to the device as a foreign buffer
    opipe = oracle_lib.CollisionPipeline(model)
    oc = opipe.contacts()
    o_in = model.state()
    opipe.collide(o_in, oc)
    foreign = pipe.contacts()
    for name, value in vars(oc).items():
        if isinstance(value, torch.Tensor) and getattr(foreign, name, None) is not None:
            n = min(value.shape[0], getattr(foreign, name).shape[0])
            getattr(foreign, name)[:n].copy_(value[:n])
    o_out = model.state()
    mk_o().step(model.state(), o_out, None, oc, dt)
    g_out = mg.state()
    mk_g().step(mg.state(), g_out, None, foreign, dt)
    np.testing.assert_array_equal(g_out.body_q.cpu().numpy(), o_out.body_q.numpy())
    np.testing.assert_array_equal(g_out.body_qd.cpu().numpy(), o_out.body_qd.numpy())


def test_deterministic_export_matches_reference_order(oracle_lib, cuda_lib):
    """CollisionPipeline(deterministic=True): the exported arrays are in the reference's global sort-key order
    (ContactSorter.sort_full, collide.py:2054-2073) - compared element by element with the oracle's sorted buffer, no
    canonicalisation - and Contacts.force follows that order."""
    for model in (scenes.mixed_worlds_model(2), scenes.box_stack_model(3, seed=1)):
        model.request_contact_attributes("force")
        mg = model.to("cuda:0")
        opipe, gpipe = oracle_lib.CollisionPipeline(model, deterministic=True), newton_b200.CollisionPipeline(mg, deterministic=True)
        oc, gc = opipe.contacts(), gpipe.contacts()
        osolver, gsolver = oracle_lib.SolverXPBD(model, iterations=4), newton_b200.solvers.SolverXPBD(mg, iterations=4)
        o0, o1, g0, g1 = model.state(), model.state(), mg.state(), mg.state()
        for _ in range(20):
            opipe.collide(o0, oc)
            osolver.step(o0, o1, None, oc, 1.0 / 240)
            o0, o1 = o1, o0
            gpipe.collide(g0, gc)
            gsolver.step(g0, g1, None, gc, 1.0 / 240)
            g0, g1 = g1, g0
        osolver.update_contacts(oc)
        gsolver.update_contacts(gc)
        n = int(oc.rigid_contact_count.item())
        assert n == int(gc.rigid_contact_count.item()) and n > 10
        for name in ("shape0", "shape1", "point0", "point1", "offset0", "offset1", "normal", "margin0", "margin1"):
            np.testing.assert_array_equal(getattr(gc, "rigid_contact_" + name)[:n].cpu().numpy(),
                                          getattr(oc, "rigid_contact_" + name)[:n].numpy(), err_msg=name)
        np.testing.assert_array_equal(gc.force[:n].cpu().numpy(), oc.force[:n].numpy())
        np.testing.assert_array_equal(g0.body_q.cpu().numpy(), o0.body_q.numpy())


def test_finite_plane_platform_bit_exact(oracle_lib, cuda_lib):
    """Finite plane with tight support AABB (a sphere beyond its extent falls past it), plane-cone through the box proxy,
    per-world static shapes."""
    model = scenes.platform_model(3)
    out = _both(model, 240, 1.0 / 240, {"iterations": 4}, oracle_lib)
    _assert_exact(*out, model)
    z = out[3].body_q.cpu().numpy().reshape(3, 4, 7)[:, :, 2]
    assert np.all(np.abs(z[:, 0] - 0.7) < 0.02) and np.all(np.abs(z[:, 1] - 0.2) < 0.02)  # on the platform / on the ground


@pytest.mark.parametrize("solver_name", ["xpbd", "featherstone"])
def test_convex_hull_pile_bit_exact(oracle_lib, cuda_lib, solver_name):
    """CONVEX_MESH shapes: hull-plane through the box proxy, hull-hull, box / sphere / capsule / cylinder
    against hulls, an off-centre wedge - vertex-scan support map, local-AABB broad phase, AABB-centre MPR seed.  Contact counts
    per substep and the final state equal the oracle's bit for bit."""
    model = scenes.hull_pile_model(3, seed=7)
    if solver_name == "xpbd":
        out = _both(model, 150, 1.0 / 240, {"iterations": 4}, oracle_lib)
        _assert_exact(*out, model)
    else:
        ref, _, rc = simulate(model, oracle_lib.CollisionPipeline, oracle_lib.SolverFeatherstone, substeps=80, dt=1.0 / 960, record_contacts=True)
        got, _, gc = simulate(model.to("cuda:0"), newton_b200.CollisionPipeline, newton_b200.solvers.SolverFeatherstone, substeps=80,
                              dt=1.0 / 960, record_contacts=True)
        assert rc == gc and rc[-1] > 20
        for name in ("body_q", "body_qd", "joint_q", "joint_qd"):
            np.testing.assert_array_equal(getattr(got, name).cpu().numpy(), getattr(ref, name).numpy(), err_msg=name)


class _WarpStyleArray:
    """What the boundary sees of a ``wp.array``: ``.ptr`` (device address), ``.size`` (elements), ``.dtype`` with ``_length_`` scalars per
    element, ``.device``.  No torch API - the arrays of a reference ``State`` / ``Control`` handed to the drop-in solvers look like this
    (``_abi.ptr``'s Warp branch)."""

    class _DType:
        def __init__(self, length):
            self._length_ = length

    def __init__(self, tensor):
        self._keep = tensor  # owns the memory
        self.ptr = tensor.data_ptr()
        width = tensor.shape[-1] if tensor.dim() > 1 else 1
        self.size = tensor.numel() // width
        self.shape = (self.size,)
        self.dtype = self._DType(width)
        self.device = str(tensor.device)


class _Bag:
    pass


def _warp_style(obj, names):
    out = _Bag()
    for n in names:
        v = getattr(obj, n, None)
        setattr(out, n, None if v is None else _WarpStyleArray(v))
    return out


def test_warp_style_arrays_cross_the_boundary(oracle_lib, cuda_lib):
    """States and controls whose arrays expose only ``.ptr`` (the reference's ``wp.array``s, rebuilt over torch memory because Warp is
    not installed) run through ``CollisionPipeline.collide`` / ``SolverXPBD.step`` and give the oracle's result bit for bit; an
    undersized foreign array is refused before the call."""
    m = _drop(scenes.quadruped_model(4, seed=5), 4, 0.5)
    ctl_fn = lambda mm: None  # noqa: E731
    ref, _, rcounts = simulate(m, oracle_lib.CollisionPipeline, oracle_lib.SolverXPBD, substeps=40, dt=0.005,
                               solver_kwargs={"iterations": 4}, record_contacts=True)
    mg = m.to("cuda:0")
    pipe, solver = newton_b200.CollisionPipeline(mg), newton_b200.solvers.SolverXPBD(mg, iterations=4)
    s0, s1, contacts, control = mg.state(), mg.state(), pipe.contacts(), mg.control()
    names = ("body_q", "body_qd", "body_f", "joint_q", "joint_qd")
    w0, w1 = _warp_style(s0, names), _warp_style(s1, names)
    wc = _warp_style(control, ("joint_f", "joint_target_q", "joint_target_qd", "joint_act"))
    counts = []
    for _ in range(40):
        s0.body_f.zero_()
        pipe.collide(w0, contacts)
        counts.append(int(contacts.rigid_contact_count.item()))
        solver.step(w0, w1, wc, contacts, 0.005)
        s0, s1, w0, w1 = s1, s0, w1, w0
    assert counts == rcounts
    np.testing.assert_array_equal(s0.body_q.cpu().numpy(), ref.body_q.numpy())
    np.testing.assert_array_equal(s0.body_qd.cpu().numpy(), ref.body_qd.numpy())
    short = _warp_style(s0, names)
    short.body_q = _WarpStyleArray(s0.body_q[:-1])
    with pytest.raises(ValueError, match="too small"):
        solver.step(short, w1, wc, contacts, 0.005)
