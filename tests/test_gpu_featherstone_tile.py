"""SolverFeatherstone(use_tile_gemm=True): H = J^T M J on the tensor cores (mma.sync m16n8k8, 3xTF32 split; reference tile path
featherstone/kernels.py:1568-1652).  Not bit-exact by construction: compared with the CPU oracle at the north-star tolerance
(1e-5 relative after the substep count), with the default FP32 path - which is bit-exact - as the control."""

import numpy as np
import pytest

import newton_b200
from newton_b200 import scenes
from tests.helpers import rel_err, simulate

pytestmark = pytest.mark.gpu


def _run(model, use_tile, substeps, dt):
    return simulate(model, newton_b200.CollisionPipeline, newton_b200.solvers.SolverFeatherstone, substeps=substeps, dt=dt,
                    solver_kwargs={"use_tile_gemm": use_tile}, record_contacts=True)


def test_tile_gemm_matches_oracle_in_flight(oracle_lib, cuda_lib):
    """64 seeded quadrupeds under PD control, falling (no contacts yet): 100 substeps of 1 ms."""
    model = scenes.quadruped_model(64, seed=2)
    ref, _, _ = simulate(model, oracle_lib.CollisionPipeline, oracle_lib.SolverFeatherstone, substeps=100, dt=1e-3)
    mg = model.to("cuda:0")
    exact, _, _ = _run(mg, False, 100, 1e-3)
    tile, _, _ = _run(mg, True, 100, 1e-3)
    np.testing.assert_array_equal(exact.joint_q.cpu().numpy(), ref.joint_q.numpy())  # control: the FP32 path is bit-exact
    for name in ("joint_q", "joint_qd", "body_q", "body_qd"):
        e = rel_err(getattr(tile, name).cpu().numpy(), getattr(ref, name).numpy())
        assert e < 1e-5, (name, e)
    assert not np.array_equal(tile.joint_qd.cpu().numpy(), ref.joint_qd.numpy())  # the tensor-core path really ran


def test_tile_gemm_standing_with_contacts(oracle_lib, cuda_lib):
    """Standing on the ground with penalty contacts, 200 substeps.  The 1e-6 differences of H move a contact's gap test across its
    threshold a substep earlier or later, and a stiff penalty contact amplifies that: the landing is compared statistically - contact
    counts within 2 %, the typical body within 5 cm and none further than 15 cm from the oracle's after 0.2 s (measured: median 1.2 cm, max 2.9 cm) - not to a parity tolerance
    (the in-flight test above holds the 1e-5)."""
    model = scenes.quadruped_model(16, seed=3)
    model.joint_q.view(16, -1)[:, 2] = 0.47
    scenes.host_fk(model, model.joint_q, model.joint_qd, model)
    ref, _, rc = simulate(model, oracle_lib.CollisionPipeline, oracle_lib.SolverFeatherstone, substeps=200, dt=1e-3, record_contacts=True)
    tile, _, tc = _run(model.to("cuda:0"), True, 200, 1e-3)
    assert rc[-1] > 0 and max(abs(a - b) for a, b in zip(tc, rc)) <= max(4, rc[-1] // 50)
    moved = np.linalg.norm(tile.body_q.cpu().numpy()[:, :3] - ref.body_q.numpy()[:, :3], axis=1)
    print(f"tile vs oracle after landing: median {np.median(moved):.3e} m, max {moved.max():.3e} m")
    assert np.median(moved) < 5e-2 and moved.max() < 0.15


def test_tile_gemm_refuses_large_articulations(cuda_lib):
    b = newton_b200.ModelBuilder()
    prev = -1
    for k in range(30):  # a 30-dof chain: more than the 24 columns the tile path is compiled for
        body = b.add_link(xform=(0.0, 0.0, -0.2 * k, 0.0, 0.0, 0.0, 1.0), mass=1.0)
        b.add_shape_sphere(body, radius=0.05)
        b.add_joint_revolute(prev, body, axis=(0.0, 1.0, 0.0), parent_xform=(0.0, 0.0, -0.2 if k else 0.0, 0.0, 0.0, 0.0, 1.0))
        prev = body
    b.add_articulation(list(range(30)))
    model = b.finalize().to("cuda:0")
    solver = newton_b200.solvers.SolverFeatherstone(model, use_tile_gemm=True)
    s0, s1 = model.state(), model.state()
    with pytest.raises(NotImplementedError):
        solver.step(s0, s1, None, None, 1e-3)
    newton_b200.solvers.SolverFeatherstone(model).step(s0, s1, None, None, 1e-3)  # the default path takes it
