"""An RL-style loop on the BASELINE quadruped batch: step, observe through an ArticulationView, reset the worlds that are "done".

    python scripts/rl_loop_example.py [envs=1024] [policy_steps=50]

Everything on the GPU goes through the package's own kernels: collide + XPBD substeps (``CollisionPipeline`` / ``SolverXPBD``),
``eval_ik`` for the joint observations, the view's strided reads (zero-copy) and masked writes, ``eval_fk`` under the reset mask.
The "policy" is a fixed PD target wiggle and the "done" signal is the base dropping below a height - placeholders for a learner.
"""

import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import newton_b200  # noqa: E402
from newton_b200 import JointType, scenes  # noqa: E402
from newton_b200.selection import ArticulationView  # noqa: E402


def main():
    envs = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    policy_steps = int(sys.argv[2]) if len(sys.argv) > 2 else 50
    substeps, dt = 4, 0.005
    model = scenes.quadruped_model(envs, device="cuda:0")
    pipeline = newton_b200.CollisionPipeline(model)
    solver = newton_b200.solvers.SolverXPBD(model, iterations=8)
    state_0, state_1, control, contacts = model.state(), model.state(), model.control(), pipeline.contacts()

    robot = ArticulationView(model, "quadruped")
    legs = ArticulationView(model, "quadruped", exclude_joint_types=[int(JointType.FREE)])
    default_root = robot.get_root_transforms(model).clone()
    default_root_vel = robot.get_root_velocities(model).clone()
    default_q = robot.get_dof_positions(model).clone()
    default_qd = robot.get_dof_velocities(model).clone()
    default_targets = legs.get_attribute("joint_target_q", control).clone()  # [envs, 1, 12]
    resets = 0
    for step in range(policy_steps):
        # action: PD targets for the 12 actuated joints (written through the view)
        action = default_targets + 0.3 * torch.sin(torch.tensor(0.2 * step, device="cuda:0"))
        legs.set_attribute("joint_target_q", control, action)
        for _ in range(substeps):
            state_0.clear_forces()
            pipeline.collide(state_0, contacts)
            solver.step(state_0, state_1, control, contacts, dt)
            state_0, state_1 = state_1, state_0
        # observation: XPBD advances body_q / body_qd; recover generalized coordinates, then read through the views (no copies)
        newton_b200.eval_ik(model, state_0, state_0.joint_q, state_0.joint_qd)
        base_height = robot.get_root_transforms(state_0)[:, 0, 2]
        joint_angles = legs.get_dof_positions(state_0)[:, 0]  # [envs, 12]
        joint_rates = legs.get_dof_velocities(state_0)[:, 0]
        observation = torch.cat([base_height[:, None], joint_angles, joint_rates], dim=1)  # what a policy would consume
        # reset the fallen robots
        done = base_height < 0.25
        if bool(done.any()):
            resets += int(done.sum())
            robot.set_root_transforms(state_0, default_root, mask=done)
            robot.set_root_velocities(state_0, default_root_vel, mask=done)
            robot.set_dof_positions(state_0, default_q, mask=done)
            robot.set_dof_velocities(state_0, default_qd, mask=done)
            robot.eval_fk(state_0, mask=done)
            solver.reset(state_0, world_mask=torch.cat([done, done.new_zeros(1)]))
    torch.cuda.synchronize()
    print(f"{policy_steps} policy steps x {substeps} substeps x {envs} envs; {resets} resets; observation {tuple(observation.shape)}; "
          f"mean base height {float(base_height.mean()):.3f} m; kernel launches {newton_b200._lib.kernel_launch_count()}")


if __name__ == "__main__":
    main()
