"""Shared helpers for parity tests: run the same loop through the oracle (CPU) and the CUDA path."""

from __future__ import annotations

import numpy as np
import torch


def simulate(model, pipeline_cls, solver_cls, *, substeps, dt, solver_kwargs=None, collide=True, control=None,
             record_contacts=False, solver_attrs=None, update_contacts=False, pipeline_kwargs=None, collide_dt=None):
    """`collide -> step -> swap` loop exactly as the reference examples run it (example_basic_urdf.py:117-135)."""
    solver = solver_cls(model, **(solver_kwargs or {}))
    for k, v in (solver_attrs or {}).items():
        setattr(solver, k, v)
    pipe = pipeline_cls(model, **(pipeline_kwargs or {})) if (collide and pipeline_cls is not None) else None
    s0, s1 = model.state(), model.state()
    ctrl = control if control is not None else model.control()
    contacts = pipe.contacts() if pipe is not None else None
    counts = []
    for _ in range(substeps):
        s0.clear_forces()
        if pipe is not None:
            if collide_dt is None:
                pipe.collide(s0, contacts)
            else:
                pipe.collide(s0, contacts, dt=collide_dt)
            if record_contacts:
                counts.append(int(contacts.rigid_contact_count.item()))
        solver.step(s0, s1, ctrl, contacts, dt)
        s0, s1 = s1, s0
    if update_contacts:
        solver.update_contacts(contacts, s0)
    return s0, contacts, counts


def canonical_contacts(contacts, model):
    """Contacts sorted by (shape0, shape1, insertion order) as numpy arrays - order-independent comparison key."""
    n = int(contacts.rigid_contact_count.item())
    n = min(n, contacts.rigid_contact_max)
    g = lambda t: t[:n].detach().cpu().numpy()  # noqa: E731
    s0, s1 = g(contacts.rigid_contact_shape0), g(contacts.rigid_contact_shape1)
    order = np.lexsort((np.arange(n), s1, s0))
    out = {"shape0": s0[order], "shape1": s1[order]}
    for name in ("point0", "point1", "offset0", "offset1", "normal", "margin0", "margin1"):
        out[name] = g(getattr(contacts, "rigid_contact_" + name))[order]
    if getattr(contacts, "force", None) is not None:
        out["force"] = g(contacts.force)[order]
    return n, out


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(1e-12, np.max(np.abs(b))))
