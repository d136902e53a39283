"""Host-side guards added after the round-1 review: C-ABI argument validation (dtype / device / element count), the native
cache is not shared by ``Model.to()`` copies, and stale ``Contacts`` objects are recognised through the generation stamp."""

import numpy as np
import pytest
import torch

from newton_b200 import _abi, scenes
from newton_b200.sim.model import Contacts


def test_ptr_rejects_wrong_dtype_device_and_size():
    a = torch.zeros(10, dtype=torch.float32)
    assert _abi.ptr(a, "f32", "cpu", 10) == a.data_ptr()
    with pytest.raises(ValueError, match="dtype"):
        _abi.ptr(a.double(), "f32")
    with pytest.raises(ValueError, match="dtype"):
        _abi.ptr(a.long(), "i32")
    with pytest.raises(ValueError, match="at least"):
        _abi.ptr(a, "f32", "cpu", 11)
    with pytest.raises(ValueError, match="model is on"):
        _abi.ptr(a, "f32", "cuda:0")
    with pytest.raises(ValueError, match="dtype"):
        _abi.ptr(np.zeros(4, dtype=np.float64), "f32")
    with pytest.raises(ValueError, match="host NumPy"):
        _abi.ptr(np.zeros(4, dtype=np.float32), "f32", "cuda:0")
    assert _abi.ptr(torch.zeros(3, dtype=torch.bool), "u8") is not None  # bool / uint8 masks are interchangeable


def test_state_and_control_views_check_sizes_against_the_model():
    m = scenes.quadruped_model(2, seed=0)
    s, c = m.state(), m.control()
    _abi.state_view(s, m)
    _abi.control_view(c, m)
    s.joint_q = s.joint_q.double()
    with pytest.raises(ValueError, match="state.joint_q"):
        _abi.state_view(s, m)
    s = m.state()
    s.body_q = s.body_q[:-1].contiguous()
    with pytest.raises(ValueError, match="state.body_q"):
        _abi.state_view(s, m)
    c.joint_target_q = c.joint_target_q[:5].contiguous()
    with pytest.raises(ValueError, match="control.joint_target_q"):
        _abi.control_view(c, m)


def test_model_to_does_not_share_native_cache():
    m = scenes.quadruped_model(1, seed=0)
    m._nb2_native = object()  # what _lib.native_model() caches on a model
    m2 = m.to("cpu")
    assert getattr(m2, "_nb2_native", None) is None
    assert m2.body_q.data_ptr() != m.body_q.data_ptr()


def test_contacts_clear_and_invalidate_drop_the_native_stamp():
    c = Contacts(8)
    c._nb2_blocks, c._nb2_stamp = object(), 5
    c.clear()
    assert c._nb2_blocks is None and c._nb2_stamp == -1
    c._nb2_blocks, c._nb2_stamp = object(), 6
    c.invalidate_native()
    assert c._nb2_blocks is None


def test_stale_contacts_are_not_treated_as_native():
    from newton_b200.solvers.solver import SolverBase

    class FakeNative:
        contact_stamp = 0

    n = FakeNative()
    solver = SolverBase.__new__(SolverBase)
    solver._native = n
    c1, c2 = Contacts(4), Contacts(4)
    for c in (c1, c2):  # what CollisionPipeline.collide() does
        n.contact_stamp += 1
        c._nb2_blocks, c._nb2_stamp = n, n.contact_stamp
    assert solver._contacts_are_native(c2)
    assert not solver._contacts_are_native(c1)  # overwritten by the second collide
    c2.clear()
    assert not solver._contacts_are_native(c2)
