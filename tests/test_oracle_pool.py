"""The oracle's native thread pool (bench.py's CPU arm, the full-size GPU parity tests): a batch run as independent world shards
must equal the monolithic oracle run bit for bit, for both solvers and for odd / even substep counts (state ping-pong)."""

import numpy as np
import pytest

from newton_b200 import scenes
from tests.helpers import simulate


@pytest.mark.parametrize("solver_name,substeps", [("xpbd", 12), ("xpbd", 13), ("featherstone", 21), ("xpbd_stacks", 30)])
def test_sharded_pool_equals_monolithic(oracle_lib, solver_name, substeps):
    oracle = oracle_lib
    if solver_name == "xpbd_stacks":  # box-box manifolds + plane contacts: the per-body summation order depends on the contact order
        m, solver_name = scenes.box_stack_model(24, seed=0), "xpbd"
    else:
        m = scenes.quadruped_model(24, seed=1)
        m.joint_q.view(24, -1)[:, 2] = 0.5
        scenes.host_fk(m, m.joint_q, m.joint_qd, m)
    if solver_name == "xpbd":
        kw, dt, cls = {"iterations": 4}, 0.005, oracle.SolverXPBD
    else:
        kw, dt, cls = {}, 0.001, oracle.SolverFeatherstone
    mono, _, counts = simulate(m, oracle.CollisionPipeline, cls, substeps=substeps, dt=dt, solver_kwargs=kw, record_contacts=True)
    pool = oracle.FramePool([m.shard(r, 6) for r in range(6)], lambda mm: cls(mm, **kw), substeps=1, dt=dt, threads=3, deterministic=True)
    sec = pool.run_frames(substeps)
    assert sec > 0.0
    states = pool.current_states()
    for name in ("body_q", "body_qd", "joint_q", "joint_qd"):
        if solver_name == "xpbd" and name.startswith("joint"):
            continue
        got = np.concatenate([getattr(s, name).numpy() for s in states])
        np.testing.assert_array_equal(got, getattr(mono, name).numpy(), err_msg=name)
    assert sum(int(c.rigid_contact_count[0]) for c in pool.contacts()) == counts[-1]
    pool.close()
