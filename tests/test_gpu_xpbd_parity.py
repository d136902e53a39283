"""GPU parity: CUDA collide + XPBD step vs the CPU oracle on identical Model/State inputs.

Bar (BASELINE.json north_star): contact shape ids / counts bit-exact; body_q / body_qd within 1e-5 relative
after 100 substeps.  Tolerances are written out below.
"""

import numpy as np
import pytest
import torch

import newton_b200
from newton_b200 import scenes
from tests.helpers import canonical_contacts, rel_err, simulate

pytestmark = pytest.mark.gpu

TOL_STATE_REL = 1e-5  # north_star tolerance for body_q / body_qd after 100 substeps


def _both(model_cpu, substeps, dt, solver_kwargs, oracle):
    ref_state, ref_contacts, ref_counts = simulate(model_cpu, oracle.CollisionPipeline, oracle.SolverXPBD, substeps=substeps,
                                                   dt=dt, solver_kwargs=solver_kwargs, record_contacts=True)
    model_gpu = model_cpu.to("cuda:0")
    gpu_state, gpu_contacts, gpu_counts = simulate(model_gpu, newton_b200.CollisionPipeline, newton_b200.solvers.SolverXPBD,
                                                   substeps=substeps, dt=dt, solver_kwargs=solver_kwargs,
                                                   record_contacts=True)
    torch.cuda.synchronize()
    return ref_state, ref_contacts, ref_counts, gpu_state, gpu_contacts, gpu_counts, model_gpu


@pytest.mark.parametrize("world_count,iterations", [(1, 2), (8, 8)])
def test_quadruped_100_substeps(oracle_lib, cuda_lib, world_count, iterations):
    model = scenes.quadruped_model(world_count, seed=1)
    # drop the robots closer to the ground so that contacts are active during most of the 100 substeps
    model.joint_q.view(world_count, -1)[:, 2] = 0.48
    newton_b200.eval_fk(model, model.joint_q, model.joint_qd, model)
    rs, rc, rcounts, gs, gc, gcounts, _ = _both(model, 100, 1.0 / 50 / 4, {"iterations": iterations}, oracle_lib)
    assert rcounts == gcounts, "per-substep rigid contact counts must be bit-exact"
    n_ref, cr = canonical_contacts(rc, model)
    n_gpu, cg = canonical_contacts(gc, model)
    assert n_ref == n_gpu and n_ref > 0
    np.testing.assert_array_equal(cr["shape0"], cg["shape0"])
    np.testing.assert_array_equal(cr["shape1"], cg["shape1"])
    eq = rel_err(gs.body_q.cpu().numpy(), rs.body_q.numpy())
    eqd = rel_err(gs.body_qd.cpu().numpy(), rs.body_qd.numpy())
    print(f"worlds={world_count} iters={iterations}: rel err body_q={eq:.3e} body_qd={eqd:.3e} contacts={n_ref}")
    assert eq < TOL_STATE_REL
    assert eqd < TOL_STATE_REL


def test_single_collide_contacts_match(oracle_lib, cuda_lib):
    """One collide() on a settled pose: every contact field must agree to fp32 rounding."""
    model = scenes.quadruped_model(4, seed=1)
    model.joint_q.view(4, -1)[:, 2] = 0.45
    newton_b200.eval_fk(model, model.joint_q, model.joint_qd, model)
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
        if k.startswith("shape"):
            np.testing.assert_array_equal(cr[k], cg[k])
        else:
            np.testing.assert_allclose(cg[k], cr[k], rtol=2e-6, atol=2e-7, err_msg=k)
