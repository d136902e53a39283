"""Digest of the exported `Contacts` arrays after a few frames - run once per setting of NB2_COLLIDE_FUSED_EXPORT / NB2_COLLIDE_WARPS
(the library reads them once per process) and compare the lines: the fused export of collide_kernel<.., EXPORT=true> must write exactly
what collide + contact_export_kernel wrote.  usage: export_digest.py ENVS quad|stacks|heap [frames]"""
import hashlib, sys, torch
sys.path.insert(0, ".")
import newton_b200
from newton_b200 import scenes

E, scene = int(sys.argv[1]), sys.argv[2]
frames = int(sys.argv[3]) if len(sys.argv) > 3 else 3
if scene == "quad":
    model, kw = scenes.quadruped_model(E, seed=1), {}
elif scene == "stacks":
    model, kw = scenes.box_stack_model(E, seed=0), {}
else:
    model, kw = scenes.free_bodies_model(E, drop_pairs=True), {"broad_phase": "sap"}
model = model.to("cuda:0")
pipe = newton_b200.CollisionPipeline(model, **kw)
solver = newton_b200.solvers.SolverXPBD(model, iterations=4)
s0, s1, ctrl, contacts = model.state(), model.state(), model.control(), pipe.contacts()
h = hashlib.sha256()
counts = []
for _ in range(frames * 4):
    s0.clear_forces(); pipe.collide(s0, contacts); solver.step(s0, s1, ctrl, contacts, 0.005); s0, s1 = s1, s0
    n = int(contacts.rigid_contact_count.item())
    counts.append(n)
    for f in ("shape0", "shape1", "point0", "point1", "offset0", "offset1", "normal", "margin0", "margin1"):
        h.update(getattr(contacts, "rigid_contact_" + f)[:n].cpu().numpy().tobytes())
solver.update_contacts if False else None
print(f"{scene} envs={E} counts={counts[0]}..{counts[-1]} digest={h.hexdigest()[:24]} state={hashlib.sha256(s0.body_q.cpu().numpy().tobytes()).hexdigest()[:16]}")
