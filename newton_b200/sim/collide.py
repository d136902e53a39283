"""``CollisionPipeline`` - drop-in for the reference class (``newton/_src/sim/collide.py:1104-2207``).

Same constructor kwargs and methods (``contacts()``, ``collide(state, contacts, *, dt=None)``) for the
in-scope configuration: ``broad_phase="explicit"`` over ``model.shape_contact_pairs``, primitive +
convex shapes.  All work happens in ``nb2_collide`` (one fused kernel,
plus scan + scatter when exporting to the ``Contacts`` arrays).
"""

from __future__ import annotations

import ctypes as C

from .. import _abi, _lib
from .model import Contacts


# reference constructor arguments (sim/collide.py:1104-1133) that only size or tune paths outside the primitive / convex scope;
# accepted with any value so that a call site spelling out the reference defaults keeps working
_IGNORED_OPTIONS = frozenset({"reduce_contacts", "max_triangle_pairs", "soft_contact_max", "shape_pairs_max", "verify_buffers",
                              "contact_reduction_hashtable_size_factor", "contact_matching_pos_threshold",
                              "contact_matching_normal_dot_threshold"})


class CollisionPipeline:
    def __init__(self, model, *, broad_phase: str | None = None, rigid_contact_max: int | None = None,
                 deterministic: bool = False, soft_contact_margin: float = 0.01, requires_grad: bool = False,
                 export_contacts: bool = True, include_static_kinematic_pairs: bool = True, **unsupported):
        if broad_phase not in (None, "explicit", "nxn", "sap"):
            raise ValueError(f"unknown broad_phase {broad_phase!r} (expected 'explicit', 'nxn' or 'sap')")
        # "nxn" / "sap" (broad_phase_nxn.py:132-218, broad_phase_sap.py) enumerate candidates at run time with the same
        # world / collision-group / filter-pair rules the builder used to precompute model.shape_contact_pairs
        # ("exact same filtering logic ... to ensure consistency between EXPLICIT mode and NXN/SAP modes",
        # sim/builder.py:12796-12797), followed by the same AABB test.  The candidate set - and therefore the contact set
        # in deterministic order - is identical, so all three options run the env-local AABB sweep over the pair list.
        if getattr(model, "shape_contact_pairs", None) is None:
            raise ValueError("model.shape_contact_pairs is missing (ModelBuilder.finalize() generates it)")
        self.broad_phase = broad_phase or "explicit"
        if requires_grad:
            raise NotImplementedError("differentiable contacts are out of scope")
        if not include_static_kinematic_pairs:
            # reference default True (collide.py:1112): pairs of two immovable shapes are kept, as in the explicit pair list;
            # False would prune them in the broad phase (broad_phase_common.py:166-201)
            raise NotImplementedError("CollisionPipeline(include_static_kinematic_pairs=False) is not implemented")
        for k, v in unsupported.items():
            if k in _IGNORED_OPTIONS:  # tuning / capacity knobs of machinery this pipeline does not have (mesh reduction, buffers)
                continue
            if k == "contact_matching" and v == "disabled":
                continue
            if v not in (None, False):
                raise NotImplementedError(f"CollisionPipeline option {k!r}={v!r} is outside the hot-path scope")
        self.model = model
        self.device = model.device
        self._native = _lib.native_model(model)
        # Contacts are always produced in a deterministic (world, sort-key) order.  deterministic=True additionally reorders
        # the exported arrays into the reference's global sort-key order (ContactSorter.sort_full, collide.py:2054-2073).
        self.deterministic = deterministic
        self.export_contacts = export_contacts
        native_max = self._native.rigid_contact_max
        if rigid_contact_max is not None and rigid_contact_max < native_max and export_contacts:
            # smaller user capacity is allowed: like the reference, the count keeps growing and excess writes are dropped
            pass
        self.rigid_contact_max = int(rigid_contact_max) if rigid_contact_max is not None else max(native_max, 1)
        self.soft_contact_margin = soft_contact_margin

    def contacts(self) -> Contacts:
        """Allocate a :class:`Contacts` buffer sized for this pipeline (reference ``collide.py:1691-1730``)."""
        c = Contacts(self.rigid_contact_max, 0, device=self.device,
                     requested_attributes=self.model._requested_contact_attributes)
        return c

    def collide(self, state, contacts, *, soft_contact_margin=None, dt=None):
        """Populate ``contacts`` from ``state.body_q`` (reference ``collide.py:1765-2207``)."""
        view = None
        # every collide() overwrites the model's single set of contact blocks: stamp them, so that a Contacts object filled by
        # an EARLIER collide (or cleared / edited since) is recognised as stale by the solvers and re-imported from its arrays
        self._native.contact_stamp += 1
        if contacts is not None:
            contacts._nb2_blocks = self._native
            contacts._nb2_stamp = self._native.contact_stamp
            contacts._nb2_exported = bool(self.export_contacts)
            if self.export_contacts:
                view = C.byref(_abi.contacts_view(contacts, self.model))
        st = _lib.lib().nb2_collide(self._native.handle, C.c_void_p(_abi.ptr(state.body_q, "f32", self.device, 7 * int(self.model.body_count), "state.body_q")), view,
                                    _lib.current_stream_ptr(self.model))
        _lib.check(st, "nb2_collide")
        if self.deterministic and view is not None:
            st = _lib.lib().nb2_contacts_sort(self._native.handle, view, _lib.current_stream_ptr(self.model))
            _lib.check(st, "nb2_contacts_sort")
