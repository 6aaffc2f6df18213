"""Tool layer of the reference (src/vapc/vapc_tools.py:6-446) over the B200 `Vapc`: the six command-line tools
(`subsample`, `mask`, `compute`, `filter`, `filter_and_compute`, `statistics`), `use_tool`, and the small
file-level helpers.  Same function names, arguments, return values and quirks — e.g. `filter_by_attributes`
returns False for an operator outside its list instead of raising (vapc_tools.py:218-225), `use_tool` only
recognises `.las` / `.ply` outputs (vapc_tools.py:413-419) — so pipelines and notebooks written against the
reference run unchanged; the per-point work happens in the CUDA library through `Vapc`.

Mesh (OBJ) tools (vapc_tools.py:451-587) are outside the hot-path scope; the point-cloud
change-detection tools `extract_point_cloud_by_3D_mask_and_label` and `extract_areas_with_change_using_mahalanobis_distance`
(vapc_tools.py:589-660) are here.
"""
from __future__ import annotations

from .bi_vapc import BiTemporalVapc  # noqa: F401  (re-exported like the reference module does)
from .datahandler import DataHandler
from .vapc import Vapc

_FILTER_OPERATORS = ("equal_to", "greater_than", "less_than", "greater_than_or_equal_to", "less_than_or_equal_to", "==", ">", "<", ">=", "<=")


def initiate_vapc(lazfile, voxel_size, origin=None):
    """Load the file(s) and hand the frame to a new Vapc (vapc_tools.py:6-35).  Returns (vapc, data handler)."""
    dh = DataHandler(lazfile)
    dh.load_las_files()
    vapc_pc = Vapc(float(voxel_size), [0, 0, 0] if origin is None else origin)
    vapc_pc.get_data_from_data_handler(dh)
    return vapc_pc, dh


def _reduce(vapc_pc, reduce_to):
    if reduce_to:
        vapc_pc.return_at = reduce_to
        vapc_pc.reduce_to_voxels()
    return vapc_pc.df


def mask(vapc_pc, maskfile, segment_in_or_out, buffer_size, reduce_to=False):
    """Keep ("in") or drop ("out") the points whose voxel lies in the buffered voxel set of `maskfile`
    (vapc_tools.py:38-99): both clouds are shifted by the common integer reduction point, the mask's voxels are
    buffered, the selection runs on the GPU, the shift is undone."""
    vapc_pc.compute_reduction_point()
    dh_mask = DataHandler(maskfile)
    dh_mask.load_las_files()
    vapc_mask = Vapc(vapc_pc.voxel_size, [0, 0, 0])
    vapc_mask.get_data_from_data_handler(dh_mask)
    vapc_mask.compute_reduction_point()
    common = tuple(min(a, b) for a, b in zip(vapc_pc.reduction_point, vapc_mask.reduction_point))
    for v in (vapc_pc, vapc_mask):
        v.reduction_point = common
        v.compute_offset()
    vapc_mask.compute_voxel_buffer(buffer_size=int(buffer_size))
    vapc_mask.voxel_index = False
    vapc_mask.df = vapc_mask.buffer_df
    vapc_pc.select_by_mask(vapc_mask, segment_in_or_out=segment_in_or_out)
    vapc_pc.compute_offset()  # toggles the shift back
    return _reduce(vapc_pc, reduce_to)


def subsample(vapc_pc, reduce_to="closest_to_center_of_gravity"):
    """One point per voxel (vapc_tools.py:102-130)."""
    vapc_pc.return_at = reduce_to
    vapc_pc.reduce_to_voxels()
    return vapc_pc.df


def compute_attributes(vapc_pc, compute, reduce_to="closest_to_center_of_gravity"):
    """The requested per-voxel attributes, optionally reduced to one row per voxel (vapc_tools.py:133-168)."""
    vapc_pc.compute = compute
    vapc_pc.compute_requested_attributes()
    return _reduce(vapc_pc, reduce_to)


def filter_by_attributes(vapc_pc, filters, reduce_to="closest_to_center_of_gravity"):
    """{attribute: {operator: value}} row filters after computing those attributes (vapc_tools.py:171-229).
    An operator outside the reference's list makes the function return False."""
    vapc_pc.compute = list(filters.keys())
    vapc_pc.compute_requested_attributes()
    for attribute, conditions in filters.items():
        for operator, value in conditions.items():
            if operator not in _FILTER_OPERATORS:
                return False
            vapc_pc.filter_attributes(attribute, operator, value)
    return _reduce(vapc_pc, reduce_to)


def filter_by_attributes_and_compute(vapc_pc, filters, compute, reduce_to):
    """Filter first, then compute what was not already needed by the filters (vapc_tools.py:232-283)."""
    vapc_pc.df = filter_by_attributes(vapc_pc, filters, reduce_to=False)
    remaining = [c for c in compute if c not in filters]
    return compute_attributes(vapc_pc, remaining, reduce_to)


def compute_statistics(vapc_pc, statistics, reduce_to="closest_to_center_of_gravity"):
    """{attribute: [statistic, ...]} per voxel (vapc_tools.py:286-326)."""
    vapc_pc.attributes = statistics
    vapc_pc.compute_requested_statistics_per_attributes()
    return _reduce(vapc_pc, reduce_to)


def use_tool(tool_name, infile, outfile, voxel_size, args, reduce_to):
    """Run one tool from file to file (vapc_tools.py:329-419)."""
    vapc_pc, dh = initiate_vapc(infile, voxel_size, origin=[0, 0, 0])
    if tool_name == "subsample":
        dh.df = subsample(vapc_pc=vapc_pc, reduce_to=args["sub_sample_method"])
    elif tool_name == "mask":
        dh.df = mask(vapc_pc=vapc_pc, maskfile=args["maskfile"], segment_in_or_out=args["segment_in_or_out"], buffer_size=args["buffer_size"],
                     reduce_to=reduce_to)
    elif tool_name == "compute":
        dh.df = compute_attributes(vapc_pc=vapc_pc, compute=args["compute"], reduce_to=reduce_to)
    elif tool_name == "filter":
        dh.df = filter_by_attributes(vapc_pc=vapc_pc, filters=args["filters"], reduce_to=reduce_to)
    elif tool_name == "filter_and_compute":
        dh.df = filter_by_attributes_and_compute(vapc_pc=vapc_pc, filters=args["filters"], compute=args["compute"], reduce_to=reduce_to)
    elif tool_name == "statistics":
        dh.df = compute_statistics(vapc_pc=vapc_pc, statistics=args["statistics"], reduce_to=reduce_to)
    else:
        raise ValueError(f"unknown tool '{tool_name}'")
    outfile = str(outfile)
    if outfile.endswith(".las"):
        dh.save_as_las(outfile)
    elif outfile.endswith(".ply"):
        dh.save_as_ply(outfile, voxel_size, shift_to_center=False)
    else:
        return "Unknown output format"


def lasz_to_ply(infile, outfile, voxel_size, shift_to_center=False):
    """A (voxelised) LAS file as cubes in a PLY file (vapc_tools.py:422-446)."""
    dh = DataHandler(infile)
    dh.load_las_files()
    dh.save_as_ply(outfile=outfile, voxel_size=voxel_size, shift_to_center=shift_to_center)
    return True


def point_cloud_to_vapc(point_cloud_file, voxel_size=1):
    """vapc_tools.py:460-466."""
    dh = DataHandler(point_cloud_file)
    dh.load_las_files()
    vp = Vapc(voxel_size)
    vp.df = dh.df
    return vp


def extract_point_cloud_by_3D_mask(scene_file, mask_file, outfile, voxel_size=1, segment_mode="in", mode="p", skiprows=2):
    """Points of `scene_file` inside / outside the voxels occupied by `mask_file` (vapc_tools.py:508-545; point-cloud
    masks only — mesh masks, mode "m", are outside the hot-path scope)."""
    if mode == "m":
        raise NotImplementedError("OBJ mesh masks are outside the scope of vapc_b200 (DESIGN.md section 7)")
    if mode != "p":
        raise ValueError("mode must be either 'm' or 'p', seperated by '_'")
    dh_scene = DataHandler(scene_file)
    dh_scene.load_las_files()
    vp_mask = point_cloud_to_vapc(mask_file, voxel_size=voxel_size)
    vp_scene = Vapc(voxel_size)
    vp_scene.df = dh_scene.df
    vp_scene.select_by_mask(vp_mask, segment_in_or_out=segment_mode)
    dh_scene.df = vp_scene.df
    dh_scene.save_as_las(outfile)


def extract_point_cloud_by_3D_mask_and_label(scene_file, mask_file, outfile, voxel_size=1, segment_mode="in", mode="p", skiprows=2):
    """Points of `scene_file` inside / outside the voxels occupied by `mask_file`, labelled with the mask's change attributes
    `mahalanobi_significance`, `p_value`, `change_type` (vapc_tools.py:589-622; point-cloud masks only — mesh masks, mode "m", are
    outside the hot-path scope).  Selection and label join run on the GPU (Vapc.select_by_mask / label_by_mask)."""
    if mode == "m":
        raise NotImplementedError("OBJ mesh masks are outside the scope of vapc_b200 (DESIGN.md section 7)")
    if mode != "p":
        raise ValueError("mode must be either 'm' or 'p', seperated by '_'")
    dh_scene = DataHandler(scene_file)
    dh_scene.load_las_files()
    vp_mask = point_cloud_to_vapc(mask_file, voxel_size=voxel_size)
    vp_scene = Vapc(voxel_size)
    vp_scene.df = dh_scene.df
    vp_scene.select_by_mask(vp_mask, segment_in_or_out=segment_mode)
    vp_scene.label_by_mask(vp_mask, ["mahalanobi_significance", "p_value", "change_type"])
    dh_scene.df = vp_scene.df
    dh_scene.save_as_las(outfile)


def extract_areas_with_change_using_mahalanobis_distance(point_cloud_1_path, point_cloud_2_path, mask_file, point_cloud_out_1_path,
                                                         point_cloud_out_2_path, voxel_size, alpha_value, delete_mask_file=True):
    """Two-epoch change detection from file to file (vapc_tools.py:625-660): per-voxel centroid + covariance of both epochs, voxel
    join, Mahalanobis test, appearing / disappearing voxels; the voxels with change_type >= 1 are written to `mask_file`, and both
    epochs are clipped to them and labelled (extract_point_cloud_by_3D_mask_and_label).  Returns False (and writes nothing) when no
    voxel changed; the mask file is kept, like the reference keeps it (its removal is commented out there)."""
    vapc_1 = point_cloud_to_vapc(point_cloud_file=point_cloud_1_path, voxel_size=voxel_size)
    vapc_2 = point_cloud_to_vapc(point_cloud_file=point_cloud_2_path, voxel_size=voxel_size)
    bi_vapc = BiTemporalVapc([vapc_1, vapc_2])
    bi_vapc.prepare_data_for_mahalanobis_distance()
    bi_vapc.merge_vapcs_with_same_voxel_index()
    bi_vapc.compute_mahalanobis_distance(alpha=alpha_value)
    bi_vapc.compute_voxels_occupied_in_single_epoch()
    bi_vapc.prepare_data_for_export()
    bi_vapc.df = bi_vapc.df[(bi_vapc.df["change_type"] >= 1)]
    if bi_vapc.df.shape[0] == 0:
        print("No change detected in the area")
        return False
    bi_vapc.save_to_las(mask_file)
    extract_point_cloud_by_3D_mask_and_label(point_cloud_1_path, mask_file, point_cloud_out_1_path, voxel_size)
    extract_point_cloud_by_3D_mask_and_label(point_cloud_2_path, mask_file, point_cloud_out_2_path, voxel_size)
