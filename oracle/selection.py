"""oracle/selection.py - TEST INFRASTRUCTURE ONLY (see oracle/README.md).

NumPy restatement of what the reference's ``ArticulationView`` (``newton/_src/utils/selection.py``) reads and writes:

* :func:`explicit_ids` - which model rows a view addresses, obtained by WALKING the model (articulation -> joints ->
  dofs / coords / child bodies -> shapes) per selected articulation, with no stride arithmetic at all.  This is the
  meaning of ``view.get_attribute(name, src)[w, a, k] == src.<name>[ids[w][a][k]]`` that the strided layouts of the reference
  (``FrequencyLayout``, :346-382; ``_get_attribute_array``, :1232-1357) and of the product must reproduce.
* :func:`gather` / :func:`scatter_masked` - ``_gather_indexed_{3,4}d_kernel`` (:185-203) and
  ``set_articulation_attribute_{3,4}d[_per_world]_kernel`` (:85-152) on a layout given as plain numbers.
* :func:`model_articulation_mask` - ``set_model_articulation_mask[_per_world]_kernel`` (:35-61).

Pinned by the reference's own known answers in ``newton/tests/test_selection.py`` (tests/test_selection.py).
"""

from __future__ import annotations

from fnmatch import fnmatch

import numpy as np


def _leaf(label: str) -> str:
    return label.rsplit("/", maxsplit=1)[-1]


def explicit_ids(model, pattern, *, exclude_joint_types=(), exclude_joints=(), exclude_links=()):
    """Per frequency, the nested list ``ids[world][articulation] -> [model row, ...]`` a view of ``pattern`` addresses.

    Filters are lists of glob strings (leaf names) / joint types - the subset of the reference's selectors the tests use.
    Returns a dict with keys ``joint``, ``dof``, ``coord``, ``link``, ``shape`` plus ``articulation`` (``[W][A]`` ids).
    """
    art_start = model.numpy("articulation_start")
    art_end = model.numpy("articulation_end")
    art_world = model.numpy("articulation_world")
    jtype = model.numpy("joint_type")
    jchild = model.numpy("joint_child")
    qs = model.numpy("joint_q_start")
    qds = model.numpy("joint_qd_start")
    W = model.world_count
    out = {k: [[] for _ in range(W)] for k in ("articulation", "joint", "dof", "coord", "link", "shape")}
    for art, label in enumerate(model.articulation_label):
        if not fnmatch(label, pattern):
            continue
        w = int(art_world[art])
        joints_all = list(range(int(art_start[art]), int(art_end[art])))
        joints = [j for j in joints_all if int(jtype[j]) not in exclude_joint_types
                  and not any(fnmatch(_leaf(model.joint_label[j]), p) for p in exclude_joints)]
        links_all = sorted({int(jchild[j]) for j in joints_all})
        links = [b for b in links_all if not any(fnmatch(_leaf(model.body_label[b]), p) for p in exclude_links)]
        out["articulation"][w].append(art)
        out["joint"][w].append(joints)
        out["dof"][w].append([d for j in joints for d in range(int(qds[j]), int(qds[j + 1]))])
        out["coord"][w].append([c for j in joints for c in range(int(qs[j]), int(qs[j + 1]))])
        out["link"][w].append(links)
        out["shape"][w].append(sorted(s for b in links for s in model.body_shapes.get(b, [])))
    return out


def take(attrib: np.ndarray, ids) -> np.ndarray:
    """``attrib[ids[w][a][k]]`` stacked into ``[W, A, K, ...]``."""
    return np.stack([np.stack([attrib[np.asarray(row, dtype=np.int64)] for row in world]) for world in ids])


def view_rows(W, A, offset, stride_between_worlds, stride_within_worlds, selection) -> np.ndarray:
    """``[W, A, K]`` attribute rows of a layout; ``selection`` = the K value indices relative to the articulation start."""
    w = np.arange(W, dtype=np.int64)[:, None, None]
    a = np.arange(A, dtype=np.int64)[None, :, None]
    k = np.asarray(selection, dtype=np.int64)[None, None, :]
    return offset + w * stride_between_worlds + a * stride_within_worlds + k


def gather(attrib: np.ndarray, rows: np.ndarray) -> np.ndarray:
    """``dst[i, j, k] = src[i, j, indices[k]]`` (selection.py:185-203): values ``[W, A, K, ...]``."""
    return attrib[rows]


def scatter_masked(attrib: np.ndarray, rows: np.ndarray, values: np.ndarray, mask=None) -> None:
    """``if view_mask[i(, j)]: attrib[i, j, k] = values[i, j, k]`` (selection.py:85-152), in place."""
    W, A, _K = rows.shape
    if mask is None:
        sel = np.ones((W, A), dtype=bool)
    else:
        mask = np.asarray(mask, dtype=bool)
        sel = np.broadcast_to(mask[:, None], (W, A)) if mask.ndim == 1 else mask
    for w in range(W):
        for a in range(A):
            if sel[w, a]:
                attrib[rows[w, a]] = values[w, a]


def model_articulation_mask(articulation_ids: np.ndarray, articulation_count: int, mask=None) -> np.ndarray:
    """selection.py:35-61 + :1727-1753."""
    out = np.zeros(articulation_count, dtype=bool)
    W, A = articulation_ids.shape
    for w in range(W):
        for a in range(A):
            on = True if mask is None else (bool(np.asarray(mask)[w]) if np.asarray(mask).ndim == 1 else bool(np.asarray(mask)[w][a]))
            if on:
                out[articulation_ids[w, a]] = True
    return out


def view_copy_product_host(attrib: np.ndarray, layout: dict, values: np.ndarray, mask=None, gather_: bool = True, wide: bool = False) -> None:
    """Runs the PRODUCT's index arithmetic (newton_b200/csrc/nb2_selection.cuh) on the host, word by word, through
    ``orc_view_copy_product_host`` - so the CPU suite can compare it with :func:`gather` / :func:`scatter_masked`."""
    import ctypes as C

    from newton_b200 import _abi

    from . import lib

    assert attrib.dtype.itemsize == 4 and values.dtype.itemsize == 4 and attrib.flags.c_contiguous and values.flags.c_contiguous
    idx = layout.get("indices")
    idx_arr = None if idx is None else np.ascontiguousarray(idx, dtype=np.int32)
    L = _abi.ViewLayout(layout["world_count"], layout["count_per_world"], layout["value_count"], layout["row_words"], layout["offset"],
                        layout["stride_between_worlds"], layout["stride_within_worlds"], layout.get("slice_start", 0),
                        None if idx_arr is None else idx_arr.ctypes.data)
    m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
    lib().orc_view_copy_product_host(C.c_void_p(attrib.ctypes.data), C.byref(L), C.c_void_p(values.ctypes.data),
                                     C.c_void_p(None if m is None else m.ctypes.data), C.c_int(0 if m is None else m.ndim),
                                     C.c_int(1 if gather_ else 0), C.c_int(1 if wide else 0))
