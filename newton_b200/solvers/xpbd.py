"""``SolverXPBD`` - drop-in for the reference class (``newton/_src/solvers/xpbd/solver_xpbd.py``).

Same constructor kwargs (``solver_xpbd.py:99-116``) and ``step(state_in, state_out, control, contacts, dt)``
signature (``:329-337``).  One fused CUDA kernel per call (``newton_b200/csrc/nb2_xpbd.cu``).
"""

from __future__ import annotations

import ctypes as C

from .. import _abi, _lib
from .solver import SolverBase


class SolverXPBD(SolverBase):
    def __init__(self, model, *, iterations: int = 2, soft_body_relaxation: float = 0.9,
                 soft_contact_relaxation: float = 0.9, joint_linear_relaxation: float = 0.7,
                 joint_angular_relaxation: float = 0.4, joint_linear_compliance: float = 0.0,
                 joint_angular_compliance: float = 0.0, rigid_contact_relaxation: float = 0.8,
                 rigid_contact_con_weighting: bool = True, angular_damping: float = 0.0,
                 enable_restitution: bool = False, deterministic=None):
        super().__init__(model)
        if getattr(model, "particle_count", 0):
            raise NotImplementedError("particles / soft bodies are outside the hot-path scope (SURVEY.md §2 row 16)")
        self.iterations = iterations
        self.soft_body_relaxation = soft_body_relaxation
        self.soft_contact_relaxation = soft_contact_relaxation
        self.joint_linear_relaxation = joint_linear_relaxation
        self.joint_angular_relaxation = joint_angular_relaxation
        self.joint_linear_compliance = joint_linear_compliance
        self.joint_angular_compliance = joint_angular_compliance
        self.rigid_contact_relaxation = rigid_contact_relaxation
        self.rigid_contact_con_weighting = rigid_contact_con_weighting
        self.angular_damping = angular_damping
        self.enable_restitution = enable_restitution
        self.compute_body_velocity_from_position_delta = False  # reference attribute (solver_xpbd.py:171)
        self._contact_impulse_capacity = None  # rigid_contact_max of the Contacts the last step accumulated impulses for

    def _params(self) -> _abi.XPBDParams:
        return _abi.XPBDParams(
            int(self.iterations), self.joint_linear_relaxation, self.joint_angular_relaxation,
            self.joint_linear_compliance, self.joint_angular_compliance, self.rigid_contact_relaxation,
            1 if self.rigid_contact_con_weighting else 0, self.angular_damping, 1 if self.enable_restitution else 0,
            1 if self.compute_body_velocity_from_position_delta else 0,
        )

    def step(self, state_in, state_out, control, contacts, dt: float) -> None:
        """Advance by ``dt`` (reference ``solver_xpbd.py:329-862``).  ``control``/``contacts`` may be ``None``."""
        model = self.model
        if control is None:
            control = model.control(clone_variables=False)
        use_contacts = 0
        if contacts is not None:
            foreign = not self._contacts_are_native(contacts)
            use_contacts = 1 if self._prepare_contacts(contacts) else 0
            if use_contacts and getattr(contacts, "force", None) is not None:  # the reference keeps impulses when contacts.force exists
                if foreign or not getattr(contacts, "_nb2_exported", False):
                    raise NotImplementedError(
                        "contacts.force needs contacts produced by newton_b200.CollisionPipeline(export_contacts=True)")
                use_contacts |= 2
                self._contact_impulse_capacity = contacts.rigid_contact_max
        p = self._params()
        st = _lib.lib().nb2_xpbd_step(
            self._native.handle, C.byref(p), C.byref(_abi.state_view(state_in, model)), C.byref(_abi.state_view(state_out, model)),
            C.byref(_abi.control_view(control, model)), use_contacts, C.c_float(dt), _lib.current_stream_ptr(model),
        )
        _lib.check(st, "nb2_xpbd_step")

    def update_contacts(self, contacts, state=None) -> None:
        """Populate ``contacts.force`` from the impulses of the last :meth:`step` (reference ``solver_xpbd.py:864-925``)."""
        if getattr(contacts, "force", None) is None:
            raise ValueError(
                "contacts.force is not allocated. Call model.request_contact_attributes('force') before creating the "
                "Contacts object."
            )
        if self._contact_impulse_capacity is None:
            raise ValueError("No contact impulse data available. Call step() before update_contacts().")
        if contacts.rigid_contact_max != self._contact_impulse_capacity:
            raise ValueError(
                f"Contacts capacity mismatch: update_contacts() received rigid_contact_max={contacts.rigid_contact_max}, "
                f"but step() used {self._contact_impulse_capacity}. Pass the same Contacts instance to both."
            )
        st = _lib.lib().nb2_xpbd_update_contacts(self._native.handle, C.byref(_abi.contacts_view(contacts, self.model)),
                                                 _lib.current_stream_ptr(self.model))
        _lib.check(st, "nb2_xpbd_update_contacts")
