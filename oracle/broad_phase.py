"""oracle/broad_phase.py - TEST INFRASTRUCTURE ONLY (see oracle/README.md).

Restatement of WHICH shape pairs the reference's dynamic broad phases may emit, independent of AABBs: the filter half of
``_nxn_broadphase_kernel`` (``newton/_src/geometry/broad_phase_nxn.py:124-216``) over the world map of
``precompute_world_map`` (``geometry/broad_phase_common.py:271-380``) - the SAP broad phase applies the same per-pair filter
(``broad_phase_sap.py``) to the pairs its sweep finds.  ``CollisionPipeline(broad_phase="nxn" | "sap")`` of the product runs
its AABB sweep over ``model.shape_contact_pairs`` (the explicit list) on the strength of the reference's statement that the
builder computes that list with "the exact same filtering logic as the broad phase kernels" (``sim/builder.py:12796-12797``);
:func:`nxn_filter_pairs` is the other side of that equation, and ``tests/test_oracle_known_answers_more.py`` checks the two
sets are equal for every scene of the test-suite.  Round 2: the product generates candidates at run time (``broadphase_kernel``), so this module also restates the AABB half:
:func:`nxn_candidate_pairs` (filter + ``check_aabb_overlap``) and :func:`sap_candidate_pairs` (projection on the fixed axis,
per-world sort, sweep while ``lower_j < upper_i``, then the same per-pair filter; ``broad_phase_sap.py:44-80, 159-420``).
Python loops: small cases only.
"""

from __future__ import annotations

import numpy as np

COLLIDE_SHAPES = 2  # ShapeFlags.COLLIDE_SHAPES (geometry/flags.py:35)
BODY_FLAG_KINEMATIC = 2  # BodyFlags.KINEMATIC (sim/enums.py:139)


def test_group_pair(group_a: int, group_b: int) -> bool:
    """broad_phase_common.py:221-238"""
    if group_a == 0 or group_b == 0:
        return False
    if group_a > 0:
        return group_a == group_b or group_b < 0
    return group_a != group_b


def test_world_and_group_pair(world_a: int, world_b: int, group_a: int, group_b: int) -> bool:
    """broad_phase_common.py:242-268"""
    if world_a != -1 and world_b != -1 and world_a != world_b:
        return False
    return test_group_pair(group_a, group_b)


def precompute_world_map(shape_world, shape_flags=None):
    """broad_phase_common.py:271-380: per regular world its colliding shapes followed by the shared (world -1) ones, then one
    dedicated segment holding only the shared shapes.  Returns (index_map, slice_ends)."""
    shape_world = np.asarray(shape_world)
    colliding = np.ones(len(shape_world), dtype=bool) if shape_flags is None else (np.asarray(shape_flags) & COLLIDE_SHAPES) != 0
    if (shape_world < -1).any():
        raise ValueError("Invalid world IDs")
    valid = np.flatnonzero(colliding)
    worlds = shape_world[valid]
    shared = valid[worlds == -1]
    index_map, slice_ends = [], []
    for w in np.unique(worlds[worlds >= 0]):
        index_map.extend(valid[worlds == w].tolist())
        index_map.extend(shared.tolist())
        slice_ends.append(len(index_map))
    index_map.extend(shared.tolist())
    slice_ends.append(len(index_map))
    return np.asarray(index_map, dtype=np.int64), np.asarray(slice_ends, dtype=np.int64)


def is_shape_pair_immovable_filtered(a, b, shape_body, body_flags, include_static_kinematic_pairs: bool) -> bool:
    """broad_phase_common.py:166-201"""
    if include_static_kinematic_pairs or len(shape_body) == 0:
        return False
    body_a, body_b = int(shape_body[a]), int(shape_body[b])
    static_a, static_b = body_a < 0, body_b < 0
    if static_a and static_b:
        return True
    if len(body_flags) == 0:
        return False
    immovable_a = static_a or (int(body_flags[body_a]) & BODY_FLAG_KINEMATIC) != 0
    immovable_b = static_b or (int(body_flags[body_b]) & BODY_FLAG_KINEMATIC) != 0
    return immovable_a and immovable_b


def nxn_filter_pairs(shape_world, shape_flags, collision_group, filter_pairs=(), shape_body=(), body_flags=(),
                     include_static_kinematic_pairs: bool = True) -> set[tuple[int, int]]:
    """Every (shape1 < shape2) the NxN kernel would test for AABB overlap and not reject (broad_phase_nxn.py:141-216)."""
    index_map, slice_ends = precompute_world_map(shape_world, shape_flags)
    excluded = {(min(a, b), max(a, b)) for a, b in filter_pairs}
    num_regular_worlds = len(slice_ends) - 1
    out: set[tuple[int, int]] = set()
    start = 0
    for segment, end in enumerate(slice_ends):
        members = index_map[start:end]
        dedicated = segment >= num_regular_worlds
        for i in range(len(members)):
            for j in range(i + 1, len(members)):
                s1, s2 = int(min(members[i], members[j])), int(max(members[i], members[j]))
                w1, w2 = int(shape_world[s1]), int(shape_world[s2])
                if w1 == -1 and w2 == -1 and not dedicated:
                    continue  # shared-vs-shared pairs are emitted once, by the dedicated segment
                if not test_world_and_group_pair(w1, w2, int(collision_group[s1]), int(collision_group[s2])):
                    continue
                if is_shape_pair_immovable_filtered(s1, s2, shape_body, body_flags, include_static_kinematic_pairs):
                    continue
                if (s1, s2) in excluded:
                    continue
                out.add((s1, s2))
        start = end
    return out


def model_nxn_pairs(model, filter_pairs=(), include_static_kinematic_pairs: bool = True) -> set[tuple[int, int]]:
    return nxn_filter_pairs(model.numpy("shape_world"), model.numpy("shape_flags"), model.numpy("shape_collision_group"), filter_pairs,
                            model.numpy("shape_body"), model.numpy("body_flags") if model.body_count else (), include_static_kinematic_pairs)


def check_aabb_overlap(lo1, hi1, lo2, hi2, cutoff: float = 0.0) -> bool:
    """broad_phase_common.py:20-37 (AABBs pre-expanded by CollisionPipeline: cutoff 0)."""
    return bool(np.all(lo1 <= hi2 + cutoff) and np.all(hi1 >= lo2 - cutoff))


def check_aabb_overlap_moving(lo1, hi1, lo2, hi2, disp1, disp2) -> bool:
    """broad_phase_common.py:41-80: swept overlap of box 1 moving by ``disp1 - disp2`` relative to box 2 (speculative contacts);
    float32 arithmetic, cutoffs 0 (the AABBs are pre-expanded)."""
    f = np.float32
    rel = np.asarray(disp1, dtype=np.float32) - np.asarray(disp2, dtype=np.float32)
    enter, exit_time = f(0.0), f(1.0)
    for axis in range(3):
        lower1, upper1, lower2, upper2, delta = f(lo1[axis]), f(hi1[axis]), f(lo2[axis]), f(hi2[axis]), f(rel[axis])
        if delta == 0.0:
            if lower1 > upper2 or upper1 < lower2:
                return False
        else:
            with np.errstate(over="ignore"):
                axis_enter, axis_exit = f(lower2 - upper1) / delta, f(upper2 - lower1) / delta
            if axis_enter > axis_exit:
                axis_enter, axis_exit = axis_exit, axis_enter
            enter = axis_enter if axis_enter > enter else enter
            exit_time = axis_exit if axis_exit < exit_time else exit_time
            if enter > exit_time:
                return False
    return True


def _overlap(lo, hi, s1, s2, displacement):
    if displacement is None:
        return check_aabb_overlap(lo[s1], hi[s1], lo[s2], hi[s2])
    return check_aabb_overlap_moving(lo[s1], hi[s1], lo[s2], hi[s2], displacement[s1], displacement[s2])


def nxn_candidate_pairs(model, aabb_lower, aabb_upper, filter_pairs=(), include_static_kinematic_pairs: bool = True, displacement=None):
    """Canonical (min, max) pairs ``_nxn_broadphase_kernel`` writes: the filter set intersected with AABB overlap (swept overlap when
    the speculative pipeline passes per-shape displacements)."""
    out = []
    for s1, s2 in sorted(model_nxn_pairs(model, filter_pairs, include_static_kinematic_pairs)):
        if _overlap(aabb_lower, aabb_upper, s1, s2, displacement):
            out.append((s1, s2))
    return out


SAP_DIRECTION = np.array([0.5935, 0.7790, 0.1235], dtype=np.float32)


def sap_candidate_pairs(model, aabb_lower, aabb_upper, filter_pairs=(), include_static_kinematic_pairs: bool = True, displacement=None,
                        sort_axis_displacement_limit=None):
    """``BroadPhaseSAP.launch`` (broad_phase_sap.py:631-849): project, sort per world segment, range by binary search, sweep.
    With ``displacement`` (speculative contacts) the projected interval is extended by the displacement along the sort axis, clamped
    to ``sort_axis_displacement_limit`` (``_sap_project_aabb``, :44-78), and the final test is the swept overlap."""
    shape_world, shape_flags = model.numpy("shape_world"), model.numpy("shape_flags")
    group, shape_body = model.numpy("shape_collision_group"), model.numpy("shape_body")
    body_flags = model.numpy("body_flags") if model.body_count else ()
    index_map, slice_ends = precompute_world_map(shape_world, shape_flags)
    excluded = {(min(a, b), max(a, b)) for a, b in filter_pairs}
    d = SAP_DIRECTION / np.float32(np.sqrt(np.float32(np.dot(SAP_DIRECTION, SAP_DIRECTION))))
    lo, hi = np.asarray(aabb_lower, dtype=np.float32), np.asarray(aabb_upper, dtype=np.float32)
    half = np.float32(0.5) * (hi - lo)
    radius = (np.abs(d)[None, :] * half).sum(axis=1, dtype=np.float32)
    center = (d[None, :] * (np.float32(0.5) * (lo + hi))).sum(axis=1, dtype=np.float32)
    plo, phi = center - radius, center + radius
    if displacement is not None:
        pd = (d[None, :] * np.asarray(displacement, dtype=np.float32)).sum(axis=1, dtype=np.float32)
        if sort_axis_displacement_limit is not None and sort_axis_displacement_limit >= 0.0:
            lim = np.float32(sort_axis_displacement_limit)
            pd = np.minimum(np.maximum(pd, -lim), lim)  # wp.clamp = min(max(x, lo), hi)
        plo = plo + np.minimum(pd, np.float32(0.0))
        phi = phi + np.maximum(pd, np.float32(0.0))
    num_regular = len(slice_ends) - 1
    out = set()
    start = 0
    for segment, end in enumerate(slice_ends):
        members = index_map[start:end]
        order = sorted(range(len(members)), key=lambda k: (plo[members[k]], k))
        lower_sorted = [plo[members[k]] for k in order]
        for i, ki in enumerate(order):
            upper = phi[members[ki]]
            j = i + 1
            while j < len(order) and lower_sorted[j] < upper:  # binary_search_segment: first lower >= upper ends the range
                a, b = int(members[ki]), int(members[order[j]])
                j += 1
                if a == b:
                    continue
                s1, s2 = min(a, b), max(a, b)
                w1, w2 = int(shape_world[s1]), int(shape_world[s2])
                if w1 == -1 and w2 == -1 and segment < num_regular:
                    continue
                if not test_world_and_group_pair(w1, w2, int(group[s1]), int(group[s2])):
                    continue
                if is_shape_pair_immovable_filtered(s1, s2, shape_body, body_flags, include_static_kinematic_pairs):
                    continue
                if (s1, s2) in excluded:
                    continue
                if _overlap(lo, hi, s1, s2, displacement):
                    out.add((s1, s2))
        start = end
    return sorted(out)
