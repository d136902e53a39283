"""``CollisionPipeline`` - drop-in for the reference class (``newton/_src/sim/collide.py:1104-2207``).

Same constructor kwargs and methods (``contacts()``, ``collide(state, contacts, *, dt=None)``) for the
in-scope configuration: ``broad_phase="explicit"`` over ``model.shape_contact_pairs`` or the run-time ``"nxn"`` / ``"sap"``
candidates, primitive + convex shapes, optional speculative contacts, contact matching.  All work
happens in ``nb2_collide`` / ``nb2_collide_speculative`` (one fused kernel, plus scan + scatter when exporting to the ``Contacts`` arrays).
"""

from __future__ import annotations

import ctypes as C

from .. import _abi, _lib
from .model import Contacts


# reference constructor arguments (sim/collide.py:1104-1133) that only size or tune paths outside the primitive / convex scope;
# accepted with any value so that a call site spelling out the reference defaults keeps working
_IGNORED_OPTIONS = frozenset({"reduce_contacts", "max_triangle_pairs", "soft_contact_max", "verify_buffers",
                              "contact_reduction_hashtable_size_factor"})
MATCH_NOT_FOUND, MATCH_BROKEN = -1, -2  # reference geometry/contact_match.py:113-117


class SpeculativeContactConfig:
    """Reference ``CollisionPipeline.SpeculativeContactConfig`` (``sim/collide.py:1076-1102``): admit contacts that are predicted to
    close within the collision-update interval ``dt`` passed to ``collide()``; the per-shape search gap grows by at most
    ``max_speculative_extension`` [m]."""

    def __init__(self, max_speculative_extension: float = 0.1):
        import math

        value = float(max_speculative_extension)
        if not math.isfinite(value) or value < 0.0:
            raise ValueError(f"max_speculative_extension must be a non-negative finite number, got {value!r}")
        self.max_speculative_extension = value

    def __repr__(self):
        return f"SpeculativeContactConfig(max_speculative_extension={self.max_speculative_extension})"


class CollisionPipeline:
    SpeculativeContactConfig = SpeculativeContactConfig

    def __init__(self, model, *, broad_phase: str | None = None, rigid_contact_max: int | None = None,
                 deterministic: bool = False, soft_contact_margin: float = 0.01, requires_grad: bool = False,
                 export_contacts: bool = True, include_static_kinematic_pairs: bool = True, contact_matching: str = "disabled",
                 contact_matching_pos_threshold: float = 0.0005, contact_matching_normal_dot_threshold: float = 0.995,
                 contact_report: bool = False, speculative_config: SpeculativeContactConfig | None = None, **unsupported):
        if broad_phase not in (None, "explicit", "nxn", "sap"):
            raise ValueError(f"unknown broad_phase {broad_phase!r} (expected 'explicit', 'nxn' or 'sap')")
        if contact_matching not in ("disabled", "latest", "sticky"):
            raise ValueError(f"contact_matching must be one of 'disabled', 'latest', 'sticky', got {contact_matching!r}")
        if contact_matching_pos_threshold < 0.0:
            raise ValueError(f"contact_matching_pos_threshold must be non-negative, got {contact_matching_pos_threshold}")
        if not -1.0 <= contact_matching_normal_dot_threshold <= 1.0:
            raise ValueError(f"contact_matching_normal_dot_threshold must be in [-1, 1], got {contact_matching_normal_dot_threshold}")
        if contact_report and contact_matching == "disabled":
            raise ValueError('contact_report=True requires contact_matching != "disabled"')
        self.contact_matching = contact_matching
        self.contact_report = bool(contact_report)
        self.contact_matching_pos_threshold = float(contact_matching_pos_threshold)
        self.contact_matching_normal_dot_threshold = float(contact_matching_normal_dot_threshold)
        if contact_matching != "disabled":
            deterministic = True  # any matching mode implies deterministic sorting (collide.py:1269-1271)
            if not export_contacts:
                raise ValueError("contact matching works on the exported Contacts arrays: export_contacts must stay True")
        self._match_reset_all = False
        self._match_reset_mask = None
        # speculative contacts (sim/collide.py:1315-1317): the writer admits predicted contacts, collide() needs dt
        if speculative_config is not None:
            SpeculativeContactConfig(speculative_config.max_speculative_extension)  # same validation for duck-typed configs
        self.speculative_config = speculative_config
        self.broad_phase = broad_phase or "explicit"
        if self.broad_phase == "explicit" and getattr(model, "shape_contact_pairs", None) is None:
            raise ValueError("model.shape_contact_pairs is missing (ModelBuilder.finalize() generates it)")
        if requires_grad:
            raise NotImplementedError("differentiable contacts are out of scope")
        shape_pairs_max = unsupported.pop("shape_pairs_max", None)
        st = getattr(model, "shape_type", None)
        if st is not None and st.numel() and bool((st == 8).any()):  # GeoType.MESH
            # mesh-vs-infinite-plane contacts are produced per vertex, unreduced (narrow_phase.py:1761-1861); the reference's default
            # reduce_contacts=True runs them through its global hash-table reduction, which this pipeline does not have
            if unsupported.get("reduce_contacts", True) is not False:
                raise NotImplementedError("models with MESH shapes need CollisionPipeline(reduce_contacts=False): mesh-plane contacts are "
                                          "written per vertex; the global contact reduction is outside the hot-path scope")
            if self.broad_phase != "explicit" or speculative_config is not None:
                raise NotImplementedError("MESH shapes are supported with broad_phase='explicit' and without speculative contacts")
        for k, v in unsupported.items():
            if k in _IGNORED_OPTIONS:  # tuning / capacity knobs of machinery this pipeline does not have (mesh reduction, buffers)
                continue
            if v not in (None, False):
                raise NotImplementedError(f"CollisionPipeline option {k!r}={v!r} is outside the hot-path scope")
        self.model = model
        self.device = model.device
        self._native = _lib.native_model(model)
        # "nxn" / "sap": candidates are generated on the device every collide() from the current AABBs with the reference's
        # run-time filters (world, collision group, excluded pairs, immovable pairs): broad_phase_nxn.py:132-218 /
        # broad_phase_sap.py:159-515 -> broadphase_kernel (csrc/nb2_collide.cu).  model.shape_contact_pairs is not read.
        # `shape_pairs_max` (reference: capacity of the global candidate buffer) is taken per world here: 0 / None = every pair.
        self.include_static_kinematic_pairs = bool(include_static_kinematic_pairs)
        mode = {"explicit": 0, "nxn": 1, "sap": 2}[self.broad_phase]
        per_world = 0
        if shape_pairs_max is not None and mode != 0:
            per_world = max(1, -(-int(shape_pairs_max) // max(1, int(model.world_count))))
        self._native.configure_broad_phase(mode, per_world, self.include_static_kinematic_pairs)
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
                     requested_attributes=self.model._requested_contact_attributes,
                     contact_matching=self.contact_matching != "disabled", contact_report=self.contact_report)
        c._contact_matching_mode = self.contact_matching
        return c

    def reset(self, world_mask=None) -> None:
        """Forget the contact-matching history (reference ``CollisionPipeline.reset``, ``sim/collide.py:1735-1752``): all of it, or
        only for contacts touching the worlds selected by ``world_mask`` (bool ``[world_count + 1]``, last entry = world -1).
        Takes effect at the next ``collide()``; masks accumulate until then."""
        if self.contact_matching == "disabled":
            return
        if world_mask is None:
            self._match_reset_all = True
            return
        from ..solvers.solver import normalize_reset_world_mask

        mask = normalize_reset_world_mask(world_mask, world_count=int(self.model.world_count), device=self.device)
        self._match_reset_mask = mask.clone() if self._match_reset_mask is None else (self._match_reset_mask | mask)

    def collide(self, state, contacts, *, soft_contact_margin=None, dt=None):
        """Populate ``contacts`` from ``state.body_q`` (reference ``collide.py:1765-2207``)."""
        if self.deterministic and contacts is not None and self.export_contacts and contacts.rigid_contact_max != self.rigid_contact_max:
            # same rule as the reference's fixed-capacity sorter (sim/collide.py:1984-1991)
            raise ValueError(f"Contacts buffer capacity ({contacts.rigid_contact_max}) does not match the deterministic sort buffer size "
                             f"({self.rigid_contact_max}). Use CollisionPipeline.contacts() or pass matching rigid_contact_max.")
        spec_dt = 0.0
        if self.speculative_config is not None:  # sim/collide.py:1823-1833
            import math

            if dt is None:
                raise ValueError("dt must be provided when speculative contacts are enabled")
            spec_dt = float(dt)
            if not math.isfinite(spec_dt) or spec_dt < 0.0:
                raise ValueError(f"dt must be a non-negative finite number, got {spec_dt!r}")
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
        nb = int(self.model.body_count)
        body_q = C.c_void_p(_abi.ptr(state.body_q, "f32", self.device, 7 * nb, "state.body_q"))
        if self.speculative_config is not None:
            st = _lib.lib().nb2_collide_speculative(
                self._native.handle, body_q, C.c_void_p(_abi.ptr(state.body_qd, "f32", self.device, 6 * nb, "state.body_qd")),
                C.c_float(spec_dt), C.c_float(float(self.speculative_config.max_speculative_extension)), view,
                _lib.current_stream_ptr(self.model))
            _lib.check(st, "nb2_collide_speculative")
        else:
            st = _lib.lib().nb2_collide(self._native.handle, body_q, view, _lib.current_stream_ptr(self.model))
            _lib.check(st, "nb2_collide")
        if self.deterministic and view is not None:
            st = _lib.lib().nb2_contacts_sort(self._native.handle, view, _lib.current_stream_ptr(self.model))
            _lib.check(st, "nb2_contacts_sort")
        if self.contact_matching != "disabled" and contacts is not None:
            if contacts.rigid_contact_match_index is None:
                raise ValueError("CollisionPipeline has contact_matching enabled but the Contacts buffer was created without "
                                 "contact_matching. Use pipeline.contacts() to create a compatible buffer.")
            if self.contact_report and contacts.rigid_contact_new_indices is None:
                raise ValueError("CollisionPipeline has contact_report enabled but the Contacts buffer was created without "
                                 "contact_report=True. Use pipeline.contacts() to create a compatible buffer.")
            mask = self._match_reset_mask
            mask_u8 = None if mask is None else mask.to(dtype=__import__("torch").uint8)
            opt = _abi.MatchOptions(self.contact_matching_pos_threshold, self.contact_matching_normal_dot_threshold,
                                    None if mask_u8 is None else mask_u8.data_ptr(), 1 if self._match_reset_all else 0,
                                    1 if self.contact_matching == "sticky" else 0)
            if self.contact_report:
                n = contacts.rigid_contact_max
                opt.new_indices = _abi.ptr(contacts.rigid_contact_new_indices, "i32", self.device, n, "contacts.rigid_contact_new_indices")
                opt.new_count = _abi.ptr(contacts.rigid_contact_new_count, "i32", self.device, 1, "contacts.rigid_contact_new_count")
                opt.broken_indices = _abi.ptr(contacts.rigid_contact_broken_indices, "i32", self.device, n, "contacts.rigid_contact_broken_indices")
                opt.broken_count = _abi.ptr(contacts.rigid_contact_broken_count, "i32", self.device, 1, "contacts.rigid_contact_broken_count")
            st = _lib.lib().nb2_contacts_match(
                self._native.handle, C.c_void_p(_abi.ptr(state.body_q)), view, C.c_void_p(contacts.rigid_contact_match_index.data_ptr()),
                C.byref(opt), _lib.current_stream_ptr(self.model))
            _lib.check(st, "nb2_contacts_match")
            if self.contact_matching == "sticky":
                # matched rows now carry last frame's geometry: the contact blocks nb2_collide left on the device no longer
                # describe this buffer, so the solvers load it through nb2_contacts_import
                contacts.invalidate_native()
            self._match_reset_all, self._match_reset_mask = False, None
        if contacts is not None:
            contacts._contact_matching_mode = self.contact_matching
