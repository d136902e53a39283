"""``newton_b200.selection.ArticulationView`` - batched access to articulation state for RL-style loops.

Mirror of the reference ``newton.selection.ArticulationView`` (``newton/_src/utils/selection.py:500-1925``): selects the
articulations whose label matches a pattern, checks that they are laid out uniformly across worlds, and exposes every
Model / State / Control attribute of the selection as a ``[world, articulation, value, ...]`` array - observations out
(``get_root_transforms``, ``get_dof_positions`` ...) and masked resets in (``set_root_transforms(..., mask=done)``,
``set_dof_positions`` ..., ``eval_fk(state, mask=done)``).  SURVEY.md §8(f) rank 2.

Host side (this file): label matching and the layout bookkeeping (``FrequencyLayout``: offset, stride between worlds,
stride within a world, selected value indices), the same decisions as the reference's constructor.  Device side
(``csrc/nb2_selection.cu`` through ``nb2_view_gather`` / ``nb2_view_scatter`` / ``nb2_view_articulation_mask`` /
``nb2_eval_fk_masked``): every copy that is not a zero-copy view.  As in the reference, a contiguous selection is returned
as a strided *view* of the attribute (writes through it alias the source array); an index-selected (non-contiguous)
selection is gathered into a staging tensor.  Arrays are torch tensors; a ``wp.transform`` / ``wp.spatial_vector`` element
of the reference appears as a trailing dimension of 7 / 6 floats.

There is no CPU path for the copies: models on the host can be inspected (layouts, names, zero-copy views) but
``set_*``, index-gathers and ``eval_fk`` need the CUDA library, like every other call of this package.
"""

from __future__ import annotations

import ctypes as C
import re
import warnings
from fnmatch import fnmatch

import numpy as np
import torch

from . import _abi, _lib
from .sim.enums import JointType
from .sim.model import AttributeFrequency, Model


class Slice:
    """Hashable ``slice`` stand-in (reference ``selection.py:328-343``)."""

    def __init__(self, start=None, stop=None):
        self.start = start
        self.stop = stop

    def __hash__(self):
        return hash((self.start, self.stop))

    def __eq__(self, other):
        return isinstance(other, Slice) and self.start == other.start and self.stop == other.stop

    def __str__(self):
        return f"({self.start}, {self.stop})"

    def get(self):
        return slice(self.start, self.stop)


def is_contiguous_slice(indices) -> bool:
    return all(indices[i] == indices[i - 1] + 1 for i in range(1, len(indices)))


class FrequencyLayout:
    """Where the selected values of one attribute frequency live (reference ``selection.py:346-382``)."""

    def __init__(self, offset: int, stride_between_worlds: int, stride_within_worlds: int, value_count: int, indices: list[int], device):
        self.offset = offset  # values to skip at the beginning of the attribute array
        self.stride_between_worlds = stride_between_worlds
        self.stride_within_worlds = stride_within_worlds
        self.value_count = value_count  # values per articulation before selection
        self.slice = None
        self.indices = None
        if len(indices) == 0:
            self.slice = slice(0, 0)
        elif is_contiguous_slice(indices):
            self.slice = slice(indices[0], indices[-1] + 1)
        else:
            self.indices = torch.tensor(indices, dtype=torch.int32, device=device)

    @property
    def is_contiguous(self) -> bool:
        return self.slice is not None

    @property
    def selected_value_count(self) -> int:
        return self.slice.stop - self.slice.start if self.slice is not None else len(self.indices)

    def __str__(self):
        indices = self.indices if self.indices is not None else self.slice
        return (f"FrequencyLayout(\n    offset: {self.offset}\n    stride_between_worlds: {self.stride_between_worlds}\n"
                f"    stride_within_worlds: {self.stride_within_worlds}\n    indices: {indices}\n)")


def get_name_from_label(label: str) -> str:
    """Leaf component of a slash-delimited label (reference ``selection.py:385-394``)."""
    return label.rsplit("/", maxsplit=1)[-1]


def match_labels(labels: list[str], pattern) -> list[int]:
    """Indices of the labels matching a glob string, list of globs, compiled regex (full match) or list of indices
    (reference ``selection.py:426-473``)."""
    if isinstance(pattern, str):
        return [idx for idx, label in enumerate(labels) if fnmatch(label, pattern)]
    if isinstance(pattern, re.Pattern):
        return [idx for idx, label in enumerate(labels) if pattern.fullmatch(label) is not None]
    if not isinstance(pattern, list):
        raise TypeError("Expected a glob string, list of glob strings, compiled string pattern, "
                        f"or list of int indices, got: {type(pattern)}")
    if len(pattern) == 0:
        return pattern
    if isinstance(pattern[0], int):
        if all(isinstance(item, int) for item in pattern):
            return pattern
    elif all(isinstance(item, str) for item in pattern):
        return [idx for idx, label in enumerate(labels) if any(fnmatch(label, item) for item in pattern)]
    types = {type(item).__name__ for item in pattern}
    raise TypeError(f"Expected a list of str patterns or a list of int indices, got: {', '.join(sorted(types))}")


def find_matching_ids(pattern, labels, world_ids, world_count: int):
    """Matching ids grouped by world + those of the global world -1 (reference ``selection.py:397-423``)."""
    matching_ids = match_labels(labels, pattern)
    if isinstance(pattern, list) and pattern and isinstance(pattern[0], int):
        for idx in range(1, len(matching_ids)):
            if matching_ids[idx] <= matching_ids[idx - 1]:
                raise ValueError("Articulation indices must be unique and in ascending order")
        if matching_ids[0] < 0 or matching_ids[-1] >= len(labels):
            raise ValueError(f"Articulation indices must be in range [0, {len(labels)})")
    grouped_ids = [[] for _ in range(world_count)]
    global_ids = []
    for idx in matching_ids:
        world = int(world_ids[idx])
        if world == -1:
            global_ids.append(idx)
        elif 0 <= world < world_count:
            grouped_ids[world].append(idx)
        else:
            raise ValueError(f"World index out of range: {world}")
    return grouped_ids, global_ids


def _same_device(a, b) -> bool:
    a, b = torch.device(a), torch.device(b)
    return a.type == b.type and (a.index or 0) == (b.index or 0)


def _all_equal(values) -> bool:
    return all(x == values[0] for x in values)


def _uniform_stride(starts, what: str) -> int | None:
    """The common difference of consecutive starts, or ``None`` when there is only one."""
    strides = [starts[i] - starts[i - 1] for i in range(1, len(starts))]
    if not strides:
        return None
    if not _all_equal(strides):
        raise ValueError(f"Non-uniform strides {what} are not supported")
    return strides[0]


class _Plan:
    """Cached addressing of one attribute array through a view (see ``ArticulationView._plan``)."""

    __slots__ = ("shape", "numel", "indices", "abi", "view", "view_ptr")


class ArticulationView:
    """Selection of identical articulations across worlds (reference ``selection.py:500-561`` for the contract).

    ``pattern`` is matched against full articulation labels; ``include_joints`` / ``exclude_joints`` / ``include_links`` /
    ``exclude_links`` against the leaf component of joint / body labels (glob, list of globs, compiled regex, or indices);
    ``include_joint_types`` / ``exclude_joint_types`` filter by :class:`JointType`.  Masks are per world ``(world_count,)``
    or per articulation ``(world_count, count_per_world)``.
    """

    def __init__(self, model: Model, pattern, *, include_joints=None, exclude_joints=None, include_links=None, exclude_links=None,
                 include_joint_types=None, exclude_joint_types=None, include_loop_closing_joints: bool = False, verbose: bool | None = None):
        self.model = model
        self.device = model.device
        self._plans: dict = {}
        for parameter_name, indices in (("include_joints", include_joints), ("include_links", include_links)):
            if (isinstance(indices, list) and all(isinstance(index, int) for index in indices)
                    and any(indices[i] < indices[i - 1] for i in range(1, len(indices)))):
                warnings.warn(f"Passing unsorted integer indices to ArticulationView({parameter_name}=...) is deprecated and "
                              "will raise a ValueError in a future release. Sort the indices in ascending order before passing them.",
                              DeprecationWarning, stacklevel=2)

        art_start = model.numpy("articulation_start")
        art_end = model.numpy("articulation_end")
        art_world = model.numpy("articulation_world")
        joint_type = model.numpy("joint_type")
        joint_child = model.numpy("joint_child")
        q_start = model.numpy("joint_q_start")
        qd_start = model.numpy("joint_qd_start")

        articulation_ids, global_articulation_ids = find_matching_ids(pattern, model.articulation_label, art_world, model.world_count)
        world_count = model.world_count
        counts_per_world = [len(ids) for ids in articulation_ids]
        articulation_count = sum(counts_per_world)
        if articulation_count > 0 and global_articulation_ids:
            raise ValueError(f"Articulation pattern '{pattern}' matches global and per-world articulations, which is currently not supported")
        if articulation_count == 0 and global_articulation_ids:  # scenes with only global articulations
            world_count = 1
            articulation_count = len(global_articulation_ids)
            counts_per_world = [articulation_count]
            articulation_ids = [global_articulation_ids]
        if articulation_count == 0:
            raise KeyError(f"No articulations matching pattern '{pattern}'")
        if not _all_equal(counts_per_world):
            raise ValueError("Varying articulation counts per world are not supported")
        count_per_world = counts_per_world[0]

        def joint_range(arti_id):
            begin = int(art_start[arti_id])
            end = int(art_start[arti_id + 1]) if include_loop_closing_joints else int(art_end[arti_id])
            return begin, end

        # the first articulation is the template for names and selections
        arti_0 = articulation_ids[0][0]
        arti_joint_begin, arti_joint_end = joint_range(arti_0)
        arti_joint_count = arti_joint_end - arti_joint_begin
        arti_joint_dof_count = int(qd_start[arti_joint_end]) - int(qd_start[arti_joint_begin])
        arti_joint_coord_count = int(q_start[arti_joint_end]) - int(q_start[arti_joint_begin])
        arti_joint_ids = list(range(arti_joint_begin, arti_joint_end))
        arti_joint_labels = [model.joint_label[j] for j in arti_joint_ids]
        arti_joint_names = [get_name_from_label(label) for label in arti_joint_labels]
        arti_joint_types = [int(joint_type[j]) for j in arti_joint_ids]
        arti_link_ids = sorted({int(joint_child[j]) for j in arti_joint_ids})  # unique bodies, in model order
        arti_link_count = len(arti_link_ids)
        arti_link_labels = [model.body_label[b] for b in arti_link_ids]
        arti_link_names = [get_name_from_label(label) for label in arti_link_labels]
        arti_shape_ids = sorted(s for b in arti_link_ids for s in model.body_shapes.get(b, []))
        arti_shape_count = len(arti_shape_ids)
        arti_shape_labels = [model.shape_label[s] for s in arti_shape_ids]
        arti_shape_names = [get_name_from_label(label) for label in arti_shape_labels]

        # per-articulation starts and counts; every articulation must look like the template
        starts = {k: [[] for _ in range(world_count)] for k in ("joint", "dof", "coord", "link", "shape")}
        counts = {k: [[] for _ in range(world_count)] for k in ("joint", "dof", "coord", "link", "shape", "root_type")}
        for world_id in range(world_count):
            for arti_id in articulation_ids[world_id]:
                joint_start, joint_end = joint_range(arti_id)
                starts["joint"][world_id].append(joint_start)
                counts["joint"][world_id].append(joint_end - joint_start)
                starts["dof"][world_id].append(int(qd_start[joint_start]))
                counts["dof"][world_id].append(int(qd_start[joint_end]) - int(qd_start[joint_start]))
                starts["coord"][world_id].append(int(q_start[joint_start]))
                counts["coord"][world_id].append(int(q_start[joint_end]) - int(q_start[joint_start]))
                counts["root_type"][world_id].append(int(joint_type[joint_start]))
                link_ids = sorted({int(joint_child[j]) for j in range(joint_start, joint_end)})
                shape_ids = [s for b in link_ids for s in model.body_shapes.get(b, [])]
                starts["link"][world_id].append(min(link_ids))
                counts["link"][world_id].append(len(link_ids))
                starts["shape"][world_id].append(min(shape_ids) if shape_ids else -1)
                counts["shape"][world_id].append(len(shape_ids))
        if not all(_all_equal(counts[k]) for k in counts):
            raise ValueError("Articulations are not identical")

        self.root_joint_type = counts["root_type"][0][0]
        root_joint_dof_count = int(qd_start[arti_joint_begin + 1] - qd_start[arti_joint_begin])
        self.is_fixed_base = root_joint_dof_count == 0  # every root degree of freedom locked
        self.is_floating_base = self.root_joint_type in (JointType.FREE, JointType.DISTANCE)

        arti_counts = {"joint": arti_joint_count, "dof": arti_joint_dof_count, "coord": arti_joint_coord_count, "link": arti_link_count,
                       "shape": arti_shape_count}
        offsets = {k: starts[k][0][0] for k in starts}
        if arti_shape_count == 0:
            offsets["shape"] = 0
        outer, inner = {}, {}
        for k in starts:  # strides between worlds / within a world must be uniform
            stride = _uniform_stride([starts[k][w][0] for w in range(world_count)], "between worlds") if world_count > 1 else None
            outer[k] = arti_counts[k] if stride is None else stride
            if count_per_world > 1:
                per_world = [_uniform_stride(starts[k][w], "within worlds") for w in range(world_count)]
                if not _all_equal(per_world):
                    raise ValueError("Non-uniform strides within worlds are not supported")
                inner[k] = per_world[0]
            else:
                inner[k] = arti_counts[k]

        # joint / link selections (local indices inside the template articulation)
        if include_joints is None and include_joint_types is None:
            joint_include = set(range(arti_joint_count))
        else:
            joint_include = set()
            if include_joints is not None:
                matching = match_labels(arti_joint_names, include_joints)
                for index in matching:
                    if index < 0 or index >= arti_joint_count:
                        raise ValueError(f"include_joints indices must be in range [0, {arti_joint_count}), got {index}")
                joint_include.update(matching)
            if include_joint_types is not None:
                joint_include.update(idx for idx in range(arti_joint_count) if arti_joint_types[idx] in include_joint_types)
        joint_exclude = set()
        if exclude_joints is not None:
            joint_exclude.update(idx for idx in match_labels(arti_joint_names, exclude_joints) if 0 <= idx < arti_joint_count)
        if exclude_joint_types is not None:
            joint_exclude.update(idx for idx in range(arti_joint_count) if arti_joint_types[idx] in exclude_joint_types)
        if include_links is None:
            link_include = set(range(arti_link_count))
        else:
            matching = match_labels(arti_link_names, include_links)
            for index in matching:
                if index < 0 or index >= arti_link_count:
                    raise ValueError(f"include_links indices must be in range [0, {arti_link_count}), got {index}")
            link_include = set(matching)
        link_exclude = set()
        if exclude_links is not None:
            link_exclude.update(idx for idx in match_labels(arti_link_names, exclude_links) if 0 <= idx < arti_link_count)
        selected_joint_indices = sorted(joint_include - joint_exclude)
        selected_link_indices = sorted(link_include - link_exclude)

        # names and value indices of what was selected
        self.joint_names, self.joint_labels = [], []
        self.joint_dof_names, self.joint_dof_counts = [], []
        self.joint_coord_names, self.joint_coord_counts = [], []
        selected_dof_indices, selected_coord_indices = [], []
        for joint_idx in selected_joint_indices:
            joint_id = arti_joint_ids[joint_idx]
            name = arti_joint_names[joint_idx]
            self.joint_names.append(name)
            self.joint_labels.append(arti_joint_labels[joint_idx])
            for starts_arr, offset, names, counts_out, selected in (
                (qd_start, offsets["dof"], self.joint_dof_names, self.joint_dof_counts, selected_dof_indices),
                (q_start, offsets["coord"], self.joint_coord_names, self.joint_coord_counts, selected_coord_indices),
            ):
                begin, end = int(starts_arr[joint_id]), int(starts_arr[joint_id + 1])
                counts_out.append(end - begin)
                if end - begin == 1:
                    names.append(name)
                    selected.append(begin - offset)
                else:
                    for k in range(end - begin):
                        names.append(f"{name}:{k}")
                        selected.append(begin + k - offset)
        self.link_names, self.link_labels, self.link_shapes = [], [], []
        selected_shape_indices, shape_link_idx = [], {}
        for link_idx, arti_link_idx in enumerate(selected_link_indices):
            body_id = arti_link_ids[arti_link_idx]
            self.link_names.append(arti_link_names[arti_link_idx])
            self.link_labels.append(arti_link_labels[arti_link_idx])
            for shape_id in model.body_shapes.get(body_id, []):
                arti_shape_idx = arti_shape_ids.index(shape_id)
                selected_shape_indices.append(arti_shape_idx)
                shape_link_idx[arti_shape_idx] = link_idx
            self.link_shapes.append([])
        selected_shape_indices = sorted(selected_shape_indices)
        self.shape_names, self.shape_labels = [], []
        for shape_idx, arti_shape_idx in enumerate(selected_shape_indices):
            self.shape_names.append(arti_shape_names[arti_shape_idx])
            self.shape_labels.append(arti_shape_labels[arti_shape_idx])
            self.link_shapes[shape_link_idx[arti_shape_idx]].append(shape_idx)

        self.count = articulation_count
        self.world_count = world_count
        self.count_per_world = count_per_world
        self.joint_count = len(selected_joint_indices)
        self.joint_dof_count = len(selected_dof_indices)
        self.joint_coord_count = len(selected_coord_indices)
        self.link_count = len(selected_link_indices)
        self.shape_count = len(selected_shape_indices)

        def layout(k, selected):
            return FrequencyLayout(offsets[k], outer[k], inner[k], arti_counts[k], selected, self.device)

        self.frequency_layouts = {
            AttributeFrequency.JOINT: layout("joint", selected_joint_indices),
            AttributeFrequency.JOINT_DOF: layout("dof", selected_dof_indices),
            AttributeFrequency.JOINT_COORD: layout("coord", selected_coord_indices),
            AttributeFrequency.BODY: layout("link", selected_link_indices),
            AttributeFrequency.SHAPE: layout("shape", selected_shape_indices),
        }
        self.tendon_count = 0  # MuJoCo fixed tendons (selection.py:1017-1163) do not exist in this package
        self.tendon_names = []
        self.joints_contiguous = self.frequency_layouts[AttributeFrequency.JOINT].is_contiguous
        self.joint_dofs_contiguous = self.frequency_layouts[AttributeFrequency.JOINT_DOF].is_contiguous
        self.joint_coords_contiguous = self.frequency_layouts[AttributeFrequency.JOINT_COORD].is_contiguous
        self.links_contiguous = self.frequency_layouts[AttributeFrequency.BODY].is_contiguous
        self.shapes_contiguous = self.frequency_layouts[AttributeFrequency.SHAPE].is_contiguous

        # (world, articulation) -> Model articulation id; default masks
        self.articulation_ids = torch.tensor(articulation_ids, dtype=torch.int32, device=self.device)
        self.full_mask = torch.ones(world_count, dtype=torch.bool).to(self.device)
        selected = np.zeros(model.articulation_count, dtype=np.bool_)
        selected[np.asarray(articulation_ids, dtype=np.int64).reshape(-1)] = True
        self.articulation_mask = torch.from_numpy(selected).to(self.device)

        if verbose:
            print(f"Articulation '{pattern}': {self.count}")
            print(f"  Link count:     {self.link_count} ({'' if self.links_contiguous else 'non-'}contiguous)")
            print(f"  Shape count:    {self.shape_count} ({'' if self.shapes_contiguous else 'non-'}contiguous)")
            print(f"  Joint count:    {self.joint_count} ({'' if self.joints_contiguous else 'non-'}contiguous)")
            print(f"  DOF count:      {self.joint_dof_count} ({'' if self.joint_dofs_contiguous else 'non-'}contiguous)")
            print(f"  Fixed base?     {self.is_fixed_base}")
            print(f"  Floating base?  {self.is_floating_base}")
            print(f"Link names:\n  {self.link_names}\nJoint names:\n  {self.joint_names}\nJoint DOF names:\n  {self.joint_dof_names}")

    @property
    def body_names(self):
        return self.link_names

    @property
    def body_shapes(self):
        return self.link_shapes

    @property
    def body_labels(self):
        return self.link_labels

    # ------------------------------------------------------------------ generic attribute API
    def _plan(self, name: str, source, _slice):
        """How ``source.<name>`` is addressed (reference ``_get_attribute_array``, :1232-1357): the attribute tensor plus a cached
        :class:`_Plan` (view shape, ``nb2_view_layout``, strided-view arguments).  Cached per array like the reference's
        ``lru_cache``, keyed by what the plan depends on - name, slice, base pointer, shape."""
        attrib = source
        for part in name.split("."):
            attrib = getattr(attrib, part)
        if not isinstance(attrib, torch.Tensor):
            raise AttributeError(f"Attribute '{name}' is not an array")
        key = (name, _slice, attrib.data_ptr(), tuple(attrib.shape))
        plan = self._plans.get(key)
        if plan is None:
            plan = self._plans[key] = self._make_plan(name, attrib, _slice)
        return attrib, plan

    def _make_plan(self, name: str, attrib, _slice) -> "_Plan":
        frequency = self.model.get_attribute_frequency(name)
        layout = self.frequency_layouts.get(frequency)
        if layout is None:
            raise AttributeError(f"Unable to determine the layout of frequency '{frequency.name}' for attribute '{name}'")
        if isinstance(_slice, Slice):
            _slice = _slice.get()
        elif not isinstance(_slice, (type(None), int, slice)):
            raise ValueError(f"Invalid slice type: expected slice or int, got {type(_slice)}")
        indices = None
        drop = False  # an int slice drops the value dimension, like NumPy / Warp indexing
        if _slice is None:
            if layout.indices is not None:
                indices, start, count = layout.indices, 0, len(layout.indices)
            else:
                start, count = layout.slice.start, layout.slice.stop - layout.slice.start
        elif isinstance(_slice, int):
            start, count, drop = _slice, 1, True
        else:
            start, count = _slice.start, _slice.stop - _slice.start
        plan = _Plan()
        trailing = tuple(attrib.shape[1:])
        lead = (self.world_count, self.count_per_world) if drop else (self.world_count, self.count_per_world, count)
        plan.shape = (*lead, *trailing)
        plan.numel = int(np.prod(plan.shape, dtype=np.int64))
        plan.indices = indices  # keeps the device index array alive
        plan.abi = None
        if attrib.element_size() == 4:  # the copy kernels move 32-bit words (reference overloads: float / int / transform / spatial_vector)
            row_words = int(np.prod(trailing, dtype=np.int64)) if trailing else 1
            plan.abi = _abi.ViewLayout(self.world_count, self.count_per_world, count, row_words, layout.offset, layout.stride_between_worlds,
                                       layout.stride_within_worlds, start, None if indices is None else indices.data_ptr())
        plan.view = None
        if indices is None:  # contiguous selection: a strided view of the array itself
            vs = attrib.stride(0) if attrib.dim() > 0 else 1
            lead_strides = (layout.stride_between_worlds * vs, layout.stride_within_worlds * vs) + (() if drop else (vs,))
            plan.view = (plan.shape, (*lead_strides, *attrib.stride()[1:]), attrib.storage_offset() + (layout.offset + start) * vs)
            plan.view_ptr = attrib.data_ptr() + (layout.offset + start) * vs * attrib.element_size()
        return plan

    def _get_attribute_array(self, name: str, source, _slice=None):
        """Zero-copy strided view for contiguous selections, else ``None`` (the caller gathers)."""
        attrib, plan = self._plan(name, source, _slice)
        if plan.view is None:
            return None
        if attrib.numel() == 0 or plan.numel == 0:
            return attrib.new_empty(plan.shape)
        return torch.as_strided(attrib, *plan.view)

    def _get_attribute_values(self, name: str, source, _slice=None):
        attrib, plan = self._plan(name, source, _slice)
        if plan.view is not None:
            return self._get_attribute_array(name, source, _slice)
        staging = attrib.new_empty(plan.shape)
        if plan.numel:
            self._launch_copy(attrib, self._copy_layout(attrib, plan), staging, None, gather=True)
        return staging

    @staticmethod
    def _copy_layout(attrib, plan):
        if plan.abi is None:
            raise NotImplementedError(f"ArticulationView copies 32-bit attributes only (got {attrib.dtype})")
        return plan.abi

    def _launch_copy(self, attrib, abi_layout, values, mask, gather: bool):
        if not attrib.is_cuda:
            raise _lib.Nb2Error("ArticulationView copies run on CUDA devices only (no CPU path); use oracle.selection for CPU checks")
        if not attrib.is_contiguous():
            raise ValueError("attribute arrays must be contiguous")
        stream = C.c_void_p(torch.cuda.current_stream(attrib.device).cuda_stream)
        with torch.cuda.device(attrib.device):
            if gather:
                st = _lib.lib().nb2_view_gather(C.c_void_p(attrib.data_ptr()), C.byref(abi_layout), C.c_void_p(values.data_ptr()), stream)
                _lib.check(st, "nb2_view_gather")
            else:
                ndim = 0 if mask is None else mask.dim()
                st = _lib.lib().nb2_view_scatter(C.c_void_p(attrib.data_ptr()), C.byref(abi_layout), C.c_void_p(values.data_ptr()),
                                                 C.c_void_p(None if mask is None else mask.data_ptr()), ndim, stream)
                _lib.check(st, "nb2_view_scatter")

    def _set_attribute_values(self, name: str, target, values, mask=None, _slice=None):
        """Masked write (reference ``_set_attribute_values``, :1380-1439)."""
        attrib, plan = self._plan(name, target, _slice)
        if not isinstance(values, torch.Tensor) or values.dtype != attrib.dtype or values.device != attrib.device:
            values = torch.as_tensor(np.asarray(values) if not isinstance(values, torch.Tensor) else values, dtype=attrib.dtype,
                                     device=attrib.device)
        if values.numel() != plan.numel:
            raise ValueError(f"Expected values with shape {plan.shape}, got {tuple(values.shape)}")
        if plan.view is not None and values.data_ptr() == plan.view_ptr and values.stride() == plan.view[1]:
            return  # in-place modification of the view returned by get_*: nothing to copy
        if not values.is_contiguous():
            values = values.contiguous()
        mask = None if mask is None else self._resolve_mask(mask)
        if plan.numel == 0:
            return
        self._launch_copy(attrib, self._copy_layout(attrib, plan), values, mask, gather=False)

    def get_attribute(self, name: str, source):
        """``[world, articulation, value, ...]`` values of ``source.<name>`` (Model, State or Control)."""
        return self._get_attribute_values(name, source)

    def set_attribute(self, name: str, target, values, mask=None) -> None:
        """Write ``values`` into ``target.<name>`` for the articulations ``mask`` selects (all by default).  After writing
        Model attributes call ``solver.notify_model_changed`` as with the reference."""
        self._set_attribute_values(name, target, values, mask=mask)

    # ------------------------------------------------------------------ convenience wrappers (selection.py:1480-1672)
    def get_root_transforms(self, source):
        if self.is_floating_base:
            return self._get_attribute_values("joint_q", source, _slice=Slice(0, 7))
        return self._get_attribute_values("joint_X_p", self.model, _slice=0)

    def set_root_transforms(self, target, values, mask=None) -> None:
        """Call :meth:`eval_fk` afterwards to move the links."""
        if self.is_floating_base:
            self._set_attribute_values("joint_q", target, values, mask=mask, _slice=Slice(0, 7))
        else:
            self._set_attribute_values("joint_X_p", self.model, values, mask=mask, _slice=0)

    def get_root_velocities(self, source):
        if self.is_floating_base:
            return self._get_attribute_values("joint_qd", source, _slice=Slice(0, 6))
        return None  # non-floating articulations have no root velocity

    def set_root_velocities(self, target, values, mask=None) -> None:
        if self.is_floating_base:
            self._set_attribute_values("joint_qd", target, values, mask=mask, _slice=Slice(0, 6))

    def get_link_transforms(self, source):
        return self._get_attribute_values("body_q", source)

    def get_link_velocities(self, source):
        """``(v_com_world, omega_world)`` per link."""
        return self._get_attribute_values("body_qd", source)

    def get_dof_positions(self, source):
        return self._get_attribute_values("joint_q", source)

    def set_dof_positions(self, target, values, mask=None) -> None:
        self._set_attribute_values("joint_q", target, values, mask=mask)

    def get_dof_velocities(self, source):
        return self._get_attribute_values("joint_qd", source)

    def set_dof_velocities(self, target, values, mask=None) -> None:
        self._set_attribute_values("joint_qd", target, values, mask=mask)

    def get_dof_forces(self, source):
        return self._get_attribute_values("joint_f", source)

    def set_dof_forces(self, target, values, mask=None) -> None:
        self._set_attribute_values("joint_f", target, values, mask=mask)

    # ------------------------------------------------------------------ masks (selection.py:1674-1753)
    def _as_mask(self, mask, shape):
        if isinstance(mask, torch.Tensor):
            return None
        try:
            arr = np.asarray(mask)
            if arr.shape != shape:
                return None
            return torch.as_tensor(arr.astype(np.bool_), device=self.device)
        except Exception:
            return None

    def _resolve_world_mask(self, mask):
        if mask is None:
            return self.full_mask
        if isinstance(mask, torch.Tensor):
            if mask.dtype != torch.bool:
                raise ValueError(f"Expected Boolean mask, got dtype {mask.dtype}")
            if tuple(mask.shape) != (self.world_count,):
                raise ValueError(f"Expected mask shape ({self.world_count},), got {tuple(mask.shape)}")
            if not _same_device(mask.device, self.device):
                raise ValueError(f"Expected mask on device {self.device}, got {mask.device}")
            return mask.contiguous()
        out = self._as_mask(mask, (self.world_count,))
        if out is None:
            raise ValueError(f"Expected Boolean mask with shape ({self.world_count},)")
        return out

    def _resolve_mask(self, mask):
        expected = {(self.world_count,), (self.world_count, self.count_per_world)}
        if isinstance(mask, torch.Tensor):
            if mask.dtype != torch.bool:
                raise ValueError(f"Expected Boolean mask, got dtype {mask.dtype}")
            if tuple(mask.shape) not in expected:
                raise ValueError(f"Expected Boolean mask with shape ({self.world_count}, {self.count_per_world}) or "
                                 f"({self.world_count},), got {tuple(mask.shape)}")
            if not _same_device(mask.device, self.device):
                raise ValueError(f"Expected mask on device {self.device}, got {mask.device}")
            return mask.contiguous()
        for shape in ((self.world_count,), (self.world_count, self.count_per_world)):
            out = self._as_mask(mask, shape)
            if out is not None:
                return out
        raise ValueError(f"Expected Boolean mask with shape ({self.world_count}, {self.count_per_world}) or ({self.world_count},)")

    def get_model_articulation_mask(self, mask=None):
        """Model articulation mask ``[articulation_count]`` from a view mask (all selected articulations by default)."""
        if mask is None:
            return self.articulation_mask
        mask = self._resolve_mask(mask)
        out = torch.empty(self.model.articulation_count, dtype=torch.bool, device=self.device)
        if not out.is_cuda:
            raise _lib.Nb2Error("get_model_articulation_mask(mask) runs on CUDA devices only (no CPU path)")
        with torch.cuda.device(out.device):
            st = _lib.lib().nb2_view_articulation_mask(
                C.c_void_p(mask.data_ptr()), mask.dim(), C.c_void_p(self.articulation_ids.data_ptr()), self.world_count,
                self.count_per_world, C.c_void_p(out.data_ptr()), self.model.articulation_count,
                C.c_void_p(torch.cuda.current_stream(out.device).cuda_stream))
        _lib.check(st, "nb2_view_articulation_mask")
        return out

    def eval_fk(self, target, mask=None) -> None:
        """Forward kinematics of the selected articulations only (reference ``selection.py:1755-1772``)."""
        from .sim.articulation import eval_fk

        eval_fk(self.model, target.joint_q, target.joint_qd, target, mask=self.get_model_articulation_mask(mask=mask))
