import sys, numpy as np, torch
sys.path.insert(0, ".")
import newton_b200, oracle
from newton_b200 import scenes
from tests.helpers import simulate
np.set_printoptions(precision=7, suppress=False, linewidth=220)
W=int(sys.argv[1]) if len(sys.argv)>1 else 1
its=int(sys.argv[2]) if len(sys.argv)>2 else 2
N=int(sys.argv[3]) if len(sys.argv)>3 else 100
model = scenes.quadruped_model(W, seed=1)
model.joint_q.view(W, -1)[:, 2] = 0.48
scenes.host_fk(model, model.joint_q, model.joint_qd, model)
kw={"iterations":its}
mg = model.to("cuda:0")
for n in [1,2,5,10,20,50,100][:]:
    if n>N: break
    rs,_,_ = simulate(model, oracle.CollisionPipeline, oracle.SolverXPBD, substeps=n, dt=0.005, solver_kwargs=kw)
    gs,_,_ = simulate(mg, newton_b200.CollisionPipeline, newton_b200.solvers.SolverXPBD, substeps=n, dt=0.005, solver_kwargs=kw)
    dq = np.abs(gs.body_q.cpu().numpy()-rs.body_q.numpy()); dqd=np.abs(gs.body_qd.cpu().numpy()-rs.body_qd.numpy())
    i = np.unravel_index(np.argmax(dqd), dqd.shape)
    print(n, "max dq", dq.max(), "max dqd", dqd.max(), "at", i, "ref", rs.body_qd.numpy()[i[0]], "gpu", gs.body_qd.cpu().numpy()[i[0]], "bitexact_q", np.array_equal(gs.body_q.cpu().numpy(), rs.body_q.numpy()))
