"""Quick timing of the collide + XPBD step kernels (L2-warm, graph replay) for kernel iteration."""
import sys, time, torch
sys.path.insert(0, ".")
import newton_b200
from newton_b200 import scenes, _lib
E = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
its = int(sys.argv[2]) if len(sys.argv) > 2 else 8
scene = sys.argv[3] if len(sys.argv) > 3 else "quad"
solver_name = sys.argv[4] if len(sys.argv) > 4 else "xpbd"
model = (scenes.quadruped_model(E, seed=1) if scene == "quad" else scenes.box_stack_model(E, seed=0)).to("cuda:0")
DT = 0.005 if solver_name == "xpbd" else 0.0025
pipe = newton_b200.CollisionPipeline(model)
if solver_name == "xpbd":
    solver = newton_b200.solvers.SolverXPBD(model, iterations=its)
else:
    import os
    solver = newton_b200.solvers.SolverFeatherstone(model, use_tile_gemm=os.environ.get("NB2_TILE", "0") == "1")
s0, s1, ctrl, contacts = model.state(), model.state(), model.control(), pipe.contacts()
def frame():
    global s0, s1
    for _ in range(4):
        s0.clear_forces(); pipe.collide(s0, contacts); solver.step(s0, s1, ctrl, contacts, DT); s0, s1 = s1, s0
st = torch.cuda.Stream()
with torch.cuda.stream(st):
    for _ in range(60): frame()
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g, stream=st): frame()
for _ in range(20): g.replay()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
N = 200
a.record()
for _ in range(N): g.replay()
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / N
print(f"scene={scene} solver={solver_name} envs={E} iters={its}: frame {ms*1e3:.1f} us  -> {E*4/ms*1e3/1e6:.2f} M env-steps/s  contacts/env={contacts.rigid_contact_count.item()/E:.1f}")
# kernel-only timings
def t(fn, n=100):
    for _ in range(10): fn()
    torch.cuda.synchronize(); a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b)/n*1e3
print("collide(+export) us:", t(lambda: pipe.collide(s0, contacts)), " xpbd_step us:", t(lambda: solver.step(s0, s1, ctrl, contacts, DT)))
