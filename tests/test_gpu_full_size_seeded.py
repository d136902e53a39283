"""BASELINE.json's full-size configurations with DISTINCT, seeded environments (configs[1], [2], [3]): the CUDA path against the
oracle run of the whole batch - every body of every environment, bit for bit, after 100 substeps.

The oracle runs the batch as independent world shards on a pool of native threads (``oracle.FramePool``; environments never
interact, and ``tests/test_sharding.py`` checks shard == monolithic), which keeps the 4096-env reference run to seconds."""

import os

import numpy as np
import pytest

import newton_b200
from newton_b200 import scenes
from tests.helpers import simulate

pytestmark = pytest.mark.gpu


def _oracle_batch(oracle, model, make_solver, substeps, dt, shard_envs=32):
    envs = int(model.world_count)
    n = max(1, envs // shard_envs)
    while envs % n:
        n -= 1
    models = [model.shard(r, n) for r in range(n)] if n > 1 else [model]
    pool = oracle.FramePool(models, make_solver, substeps=1, dt=dt, threads=min(os.cpu_count() or 1, n), deterministic=True)
    pool.run_frames(substeps)
    states, contacts = pool.current_states(), pool.contacts()
    out = {name: np.concatenate([getattr(s, name).numpy() for s in states]) for name in ("body_q", "body_qd", "joint_q", "joint_qd")}
    out["contact_count"] = sum(int(c.rigid_contact_count[0]) for c in contacts)
    pool.close()
    return out


@pytest.mark.parametrize("solver_name", ["xpbd", "featherstone"])
def test_4096_seeded_quadrupeds_100_substeps(oracle_lib, cuda_lib, solver_name):
    """configs[2] / configs[3]: 4096 quadrupeds built exactly like bench.py's scene (per-env joint perturbation, default_rng(1))."""
    envs, n = 4096, 100
    model = scenes.quadruped_model(envs, seed=1)
    if solver_name == "xpbd":
        kw, dt = {"iterations": 8}, 0.005
        mk_o, cls_g = (lambda m: oracle_lib.SolverXPBD(m, **kw)), newton_b200.solvers.SolverXPBD
    else:
        kw, dt = {}, 0.001
        mk_o, cls_g = (lambda m: oracle_lib.SolverFeatherstone(m)), newton_b200.solvers.SolverFeatherstone
    ref = _oracle_batch(oracle_lib, model, mk_o, n, dt)
    out, contacts, counts = simulate(model.to("cuda:0"), newton_b200.CollisionPipeline, cls_g, substeps=n, dt=dt, solver_kwargs=kw,
                                     record_contacts=True)
    # the last recorded count belongs to the collide of the final substep, as does the oracle's buffer
    assert counts[-1] == ref["contact_count"]
    if solver_name == "xpbd":
        assert counts[-1] > 8 * envs  # standing on the ground by then
    names = ("body_q", "body_qd") if solver_name == "xpbd" else ("body_q", "body_qd", "joint_q", "joint_qd")
    for name in names:
        got = getattr(out, name).cpu().numpy()
        np.testing.assert_array_equal(got, ref[name], err_msg=name)
    # distinct environments: the batch is not 4096 copies of one trajectory
    q = out.body_q.cpu().numpy().reshape(envs, -1)
    assert np.unique(q.round(6), axis=0).shape[0] > envs // 2


def test_512_seeded_box_stacks_100_substeps(oracle_lib, cuda_lib):
    """configs[1]: 512 five-box stacks, each yawed by its own seeded angle (default_rng(0)); box-box MPR manifolds + XPBD."""
    envs, n, dt, kw = 512, 100, 1.0 / 240, {"iterations": 8}
    model = scenes.box_stack_model(envs, seed=0)
    ref = _oracle_batch(oracle_lib, model, lambda m: oracle_lib.SolverXPBD(m, **kw), n, dt, shard_envs=8)
    out, _, counts = simulate(model.to("cuda:0"), newton_b200.CollisionPipeline, newton_b200.solvers.SolverXPBD, substeps=n, dt=dt,
                              solver_kwargs=kw, record_contacts=True)
    assert counts[-1] == ref["contact_count"] and counts[-1] >= 20 * envs - envs
    for name in ("body_q", "body_qd"):
        np.testing.assert_array_equal(getattr(out, name).cpu().numpy(), ref[name], err_msg=name)
