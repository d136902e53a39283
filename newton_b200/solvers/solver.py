"""``SolverBase`` - mirror of the reference solver plugin API (``newton/_src/solvers/solver.py:190-450``)."""

from __future__ import annotations

import ctypes as C

from .. import _abi, _lib


def normalize_reset_world_mask(world_mask, *, world_count: int, device, allow_legacy: bool = False):
    """Validate a reset mask and return it in the canonical ``(world_count + 1,)`` shape, last entry = world ``-1``
    (reference ``core/reset.py:21-60``).  ``allow_legacy`` also accepts the deprecated ``(world_count,)`` shape."""
    import warnings

    import torch

    if world_mask is None:
        return None
    if not isinstance(world_mask, torch.Tensor):
        raise TypeError("'world_mask' must be a torch tensor or None.")
    if world_mask.dtype != torch.bool:
        raise TypeError("'world_mask' must have dtype bool.")
    if world_mask.dim() != 1:
        raise ValueError("'world_mask' must be one-dimensional.")
    want = torch.device(device)
    if world_mask.device.type != want.type or (world_mask.device.index or 0) != (want.index or 0):
        raise ValueError(f"'world_mask' device {world_mask.device} does not match expected device {device}.")
    mask_size = world_mask.shape[0]
    if mask_size == world_count + 1:
        return world_mask
    if allow_legacy and mask_size == world_count:
        warnings.warn("world_mask with shape (world_count,) is deprecated; use shape (world_count + 1,), "
                      "where the final entry selects global entities in world -1.", DeprecationWarning, stacklevel=4)
        normalized = torch.zeros(world_count + 1, dtype=torch.bool, device=world_mask.device)
        normalized[:world_count] = world_mask
        return normalized
    if allow_legacy:
        raise ValueError(f"world_mask has size {mask_size}, expected {world_count} or {world_count + 1}.")
    raise ValueError(f"'world_mask' length {mask_size} must equal model.world_count + 1 ({world_count + 1}).")


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
        if not self._contacts_are_native(contacts):
            if getattr(contacts, "_nb2_blocks", None) is self._native and not getattr(contacts, "_nb2_exported", True):
                raise ValueError("these Contacts were produced with export_contacts=False and the model's contact blocks have "
                                 "been overwritten by a later collide(): nothing is left to import")
            self._native.contact_stamp += 1  # the import overwrites the blocks: whatever was stamped before is stale now
            st = _lib.lib().nb2_contacts_import(self._native.handle, C.byref(_abi.contacts_view(contacts, self.model)),
                                                _lib.current_stream_ptr(self.model))
            _lib.check(st, "nb2_contacts_import")
        return True

    def _contacts_are_native(self, contacts) -> bool:
        """True when the model's contact blocks still hold exactly what ``contacts`` describes: produced by the LAST
        ``collide()`` / import on this model and not cleared since (``Contacts._nb2_stamp`` vs ``NativeModel.contact_stamp``)."""
        return (getattr(contacts, "_nb2_blocks", None) is self._native
                and getattr(contacts, "_nb2_stamp", -1) == self._native.contact_stamp)

    def notify_model_changed(self, flags: int) -> None:
        """Reference ``solver.py:394-429``: the kernels read the Model arrays live; refresh borrowed pointers."""
        self._native.notify_model_changed(int(flags))

    def update_contacts(self, contacts, state=None) -> None:
        raise NotImplementedError()

    def reset(self, state, world_mask=None, flags=None) -> None:
        """Reference ``solver.py:344-375``: the base implementation validates the mask and resets nothing; neither
        ``SolverXPBD`` nor ``SolverFeatherstone`` overrides it upstream (they keep no per-world solver state)."""
        normalize_reset_world_mask(world_mask, world_count=int(self.model.world_count), device=self.model.device, allow_legacy=True)

    @classmethod
    def register_custom_attributes(cls, builder) -> None:
        return None

    def integrate_bodies(self, model, state_in, state_out, dt: float, angular_damping: float = 0.0) -> None:
        """Semi-implicit Euler on all bodies (reference ``solver.py:267-307``, kernel ``:112-170``)."""
        st = _lib.lib().nb2_integrate_bodies(
            self._native.handle, C.byref(_abi.state_view(state_in, self.model)), C.byref(_abi.state_view(state_out, self.model)),
            C.c_float(angular_damping), C.c_float(dt), _lib.current_stream_ptr(self.model),
        )
        _lib.check(st, "nb2_integrate_bodies")
