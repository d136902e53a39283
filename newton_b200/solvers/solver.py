"""``SolverBase`` - mirror of the reference solver plugin API (``newton/_src/solvers/solver.py:190-450``)."""

from __future__ import annotations

import ctypes as C

from .. import _abi, _lib


class SolverBase:
    def __init__(self, model):
        self.model = model
        self._native = _lib.native_model(model)

    @property
    def device(self):
        return self.model.device

    def step(self, state_in, state_out, control, contacts, dt: float) -> None:
        raise NotImplementedError()

    def _prepare_contacts(self, contacts) -> bool:
        """True when ``contacts`` should be consumed.  Contacts produced by ``newton_b200.CollisionPipeline.collide()`` on this
        model already sit in the native env-major contact blocks; any other ``Contacts`` buffer with the reference's
        attributes (the reference's own ``CollisionPipeline``, hand-written contacts ...) is loaded through
        ``nb2_contacts_import`` first - per world, in array order."""
        if contacts is None or not contacts.rigid_contact_max:
            return False
        if getattr(contacts, "_nb2_blocks", None) is not self._native:
            st = _lib.lib().nb2_contacts_import(self._native.handle, C.byref(_abi.contacts_view(contacts)),
                                                _lib.current_stream_ptr(self.model))
            _lib.check(st, "nb2_contacts_import")
        return True

    def notify_model_changed(self, flags: int) -> None:
        """Reference ``solver.py:394-429``: the kernels read the Model arrays live; refresh borrowed pointers."""
        self._native.notify_model_changed(int(flags))

    def update_contacts(self, contacts, state=None) -> None:
        raise NotImplementedError()

    def reset(self, state, world_mask=None, flags=None) -> None:
        """Reference ``solver.py:344-375``: XPBD / Featherstone keep no per-world solver state to reset."""
        return None

    @classmethod
    def register_custom_attributes(cls, builder) -> None:
        return None

    def integrate_bodies(self, model, state_in, state_out, dt: float, angular_damping: float = 0.0) -> None:
        """Semi-implicit Euler on all bodies (reference ``solver.py:267-307``, kernel ``:112-170``)."""
        st = _lib.lib().nb2_integrate_bodies(
            self._native.handle, C.byref(_abi.state_view(state_in)), C.byref(_abi.state_view(state_out)),
            C.c_float(angular_damping), C.c_float(dt), _lib.current_stream_ptr(self.model),
        )
        _lib.check(st, "nb2_integrate_bodies")
