"""Worker of tests/test_gpu_multi.py (launched by torch.distributed.run, one rank per GPU): a world-range-sharded simulation whose
end-of-frame state is gathered (a) with NVLink peer writes on the copy engines (PeerStateGather) and (b) with NCCL; both must
equal the monolithic CPU oracle run of the WHOLE batch, bit for bit, on every rank."""

import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import newton_b200
    import oracle
    from newton_b200 import scenes
    from newton_b200.sim.sharding import PeerStateGather
    from tests.helpers import simulate

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=dev)
    envs, frames, substeps, dt, kw = 16 * world, 6, 4, 0.005, {"iterations": 4}
    full = scenes.quadruped_model(envs, seed=3)
    full.joint_q.view(envs, -1)[:, 2] = 0.5
    scenes.host_fk(full, full.joint_q, full.joint_qd, full)
    model = full.shard(rank, world).to(dev)
    pipe, solver = newton_b200.CollisionPipeline(model), newton_b200.solvers.SolverXPBD(model, **kw)
    s0, s1, ctrl, contacts = model.state(), model.state(), model.control(), pipe.contacts()
    gather = PeerStateGather([s0.body_q, s0.body_qd])
    for _ in range(frames):
        for _ in range(substeps):
            s0.clear_forces()
            pipe.collide(s0, contacts)
            solver.step(s0, s1, ctrl, contacts, dt)
            s0, s1 = s1, s0
        seq = gather.push([s0.body_q, s0.body_qd])  # every frame, overlapped with the next one
    gather.wait(seq)
    gq, gqd = [t.clone() for t in gather.gathered(seq)]
    nq = torch.empty((world,) + tuple(s0.body_q.shape), device=dev)
    nqd = torch.empty((world,) + tuple(s0.body_qd.shape), device=dev)
    dist.all_gather_into_tensor(nq, s0.body_q)
    dist.all_gather_into_tensor(nqd, s0.body_qd)
    torch.cuda.synchronize()
    assert torch.equal(gq, nq) and torch.equal(gqd, nqd), "peer gather != NCCL all-gather"
    ref, _, _ = simulate(full, oracle.CollisionPipeline, oracle.SolverXPBD, substeps=frames * substeps, dt=dt, solver_kwargs=kw)
    assert np.array_equal(gq.reshape(-1, 7).cpu().numpy(), ref.body_q.numpy()), "gathered body_q != monolithic oracle"
    assert np.array_equal(gqd.reshape(-1, 6).cpu().numpy(), ref.body_qd.numpy()), "gathered body_qd != monolithic oracle"
    gather.close()
    dist.barrier()
    if rank == 0:
        print(f"MULTI_GPU_OK world={world} launches={newton_b200._lib.kernel_launch_count()}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
