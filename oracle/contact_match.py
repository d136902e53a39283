"""oracle/contact_match.py - TEST INFRASTRUCTURE ONLY.

NumPy restatement of the reference's frame-to-frame contact matcher (``newton/_src/geometry/contact_match.py``:
``_match_contacts_kernel`` :266-354, ``_resolve_claims_kernel`` :357-390, ``_save_sorted_state_kernel`` :442-477; ``"sticky"``:
``_replay_matched_kernel`` :529-561; ``contact_report``: ``_collect_contact_report_kernel`` :568-595) on the sorted
``Contacts`` arrays, in the call order of ``CollisionPipeline.collide`` (``sim/collide.py:2033-2137``): match, replay, report, save.  Sort keys follow ``make_contact_sort_key`` (``geometry/contact_data.py:59-87``) with the contact's position
inside its pair's run as the sub key."""

from __future__ import annotations

import numpy as np

MATCH_NOT_FOUND, MATCH_BROKEN = -1, -2


def _qrot(q, v):
    x, y, z, w = q
    qv = np.array([x, y, z], dtype=np.float32)
    v = np.asarray(v, dtype=np.float32)
    return v * np.float32(2.0 * w * w - 1.0) + qv * np.float32(2.0 * np.dot(qv, v)) + np.cross(qv, v) * np.float32(2.0 * w)


def _world(point, body, body_q):
    if body == -1:
        return np.asarray(point, dtype=np.float32)
    X = body_q[body]
    return _qrot(X[3:], point) + X[:3]


class ContactMatcher:
    def __init__(self, model, pos_threshold: float = 0.0005, normal_dot_threshold: float = 0.995, sticky: bool = False):
        self.sticky = bool(sticky)
        self.prev_record = None  # sticky: (point0, point1, offset0, offset1) of the saved frame
        self.new_indices = np.zeros(0, dtype=np.int32)     # contact_report of the last match(): rows without a match, ascending
        self.broken_indices = np.zeros(0, dtype=np.int32)  # rows of the previous frame nothing matched (and not reset), ascending
        self.shape_body = model.numpy("shape_body")
        self.shape_world = model.numpy("shape_world")
        self.world_count = int(model.world_count)
        self.pos_threshold_sq = np.float32(pos_threshold) * np.float32(pos_threshold)
        self.normal_dot_threshold = np.float32(normal_dot_threshold)
        self.prev_keys = np.zeros(0, dtype=np.int64)
        self.prev_pos = np.zeros((0, 3), dtype=np.float32)
        self.prev_normal = np.zeros((0, 3), dtype=np.float32)
        self.reset_mask = None

    def reset(self, world_mask=None):
        if world_mask is None:
            self.prev_keys = self.prev_keys[:0]
        else:
            m = np.asarray(world_mask, dtype=bool)
            self.reset_mask = m if self.reset_mask is None else (self.reset_mask | m)

    def _selected(self, world):
        if self.reset_mask is None:
            return False
        if 0 <= world < self.world_count:
            return bool(self.reset_mask[world])
        return world == -1 and bool(self.reset_mask[self.world_count])

    def match(self, contacts, body_q):
        """Returns match_index[:count] for the (sorted) contacts and stores them as the new history."""
        n = min(int(contacts.rigid_contact_count[0]), contacts.rigid_contact_max)
        s0, s1 = contacts.rigid_contact_shape0[:n].numpy().astype(np.int64), contacts.rigid_contact_shape1[:n].numpy().astype(np.int64)
        p0, p1 = contacts.rigid_contact_point0[:n].numpy(), contacts.rigid_contact_point1[:n].numpy()
        nrm = contacts.rigid_contact_normal[:n].numpy()
        body_q = np.asarray(body_q, dtype=np.float32).reshape(-1, 7)
        prefix = ((s0 & 0xFFFFF) << 43) | ((s1 & 0xFFFFF) << 23)
        assert np.all(prefix[1:] >= prefix[:-1]), "contact matching needs key-sorted contacts (deterministic=True)"
        keys = prefix.copy()
        for i in range(n):
            keys[i] |= i - int(np.searchsorted(prefix, prefix[i], side="left"))
        pos = np.stack([np.float32(0.5) * (_world(p0[i], self.shape_body[s0[i]], body_q) + _world(p1[i], self.shape_body[s1[i]], body_q))
                        for i in range(n)]) if n else np.zeros((0, 3), np.float32)
        match = np.full(n, MATCH_NOT_FOUND, dtype=np.int32)
        claim = {}
        for i in range(n):
            if len(self.prev_keys) == 0:
                continue
            if self._selected(int(self.shape_world[s0[i]])) or self._selected(int(self.shape_world[s1[i]])):
                continue
            lo = int(np.searchsorted(self.prev_keys, prefix[i], side="left"))
            hi = int(np.searchsorted(self.prev_keys, prefix[i] + 0x800000, side="left"))
            if lo >= hi:
                continue
            best, best_d = -1, self.pos_threshold_sq
            for k in range(lo, hi):
                diff = pos[i] - self.prev_pos[k]
                dsq = np.float32(np.dot(diff, diff))
                if dsq <= best_d and np.float32(np.dot(nrm[i], self.prev_normal[k])) >= self.normal_dot_threshold:
                    best_d, best = dsq, k
            if best >= 0:
                match[i] = best
                c = (float(best_d), int(keys[i]) & 0xFFFFFFFF)
                if best not in claim or c < claim[best]:
                    claim[best] = c
            else:
                match[i] = MATCH_BROKEN
        was_matched = np.zeros(len(self.prev_keys), dtype=bool)
        for i in range(n):
            if match[i] < 0:
                continue
            if claim[int(match[i])][1] != (int(keys[i]) & 0xFFFFFFFF):
                match[i] = MATCH_BROKEN
            else:
                was_matched[int(match[i])] = True
        if self.sticky and self.prev_record is not None:  # replay: matched rows still in contact keep last frame's record
            o0, o1 = contacts.rigid_contact_offset0[:n].numpy(), contacts.rigid_contact_offset1[:n].numpy()
            m0, m1 = contacts.rigid_contact_margin0[:n].numpy(), contacts.rigid_contact_margin1[:n].numpy()
            for i in range(n):
                k = int(match[i])
                if k < 0:
                    continue
                w0, w1 = _world(p0[i], self.shape_body[s0[i]], body_q), _world(p1[i], self.shape_body[s1[i]], body_q)
                if np.float32(np.dot(w1 - w0, nrm[i])) - (m0[i] + m1[i]) > 0.0:
                    continue
                p0[i], p1[i], o0[i], o1[i] = (self.prev_record[f][k] for f in range(4))  # views: writes land in `contacts`
                nrm[i] = self.prev_normal[k]
            pos = np.stack([np.float32(0.5) * (_world(p0[i], self.shape_body[s0[i]], body_q) + _world(p1[i], self.shape_body[s1[i]], body_q))
                            for i in range(n)]) if n else pos
        # report (before the history is replaced)
        self.new_indices = np.flatnonzero(match < 0).astype(np.int32)
        broken = []
        for k in range(len(self.prev_keys)):
            key = int(self.prev_keys[k])
            a, b = (key >> 43) & 0xFFFFF, (key >> 23) & 0xFFFFF
            if not was_matched[k] and not (self._selected(int(self.shape_world[a])) or self._selected(int(self.shape_world[b]))):
                broken.append(k)
        self.broken_indices = np.asarray(broken, dtype=np.int32)
        self.prev_keys, self.prev_pos, self.prev_normal = keys, pos.astype(np.float32), nrm.copy()
        if self.sticky:
            self.prev_record = tuple(a[:n].numpy().copy() for a in (contacts.rigid_contact_point0, contacts.rigid_contact_point1,
                                                                     contacts.rigid_contact_offset0, contacts.rigid_contact_offset1))
        self.reset_mask = None
        return match
