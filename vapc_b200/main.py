"""JSON-configured command line of the reference (src/vapc/main.py:17-240, run.py) on the B200 package:

    python -m vapc_b200.main config.json

`do_vapc_on_files` keeps the reference's arguments, time-stamped output names (`<tool>_<timestamp>.las|.laz|.ply`
and `<tool>_<timestamp>_cfg.json`) and return value (the path written).  The reference tiles large inputs through
temporary files (`tile`, main.py:97-137) because its pandas path does not fit them in host memory; here a cloud of
1e9 points fits one B200, so `tile` is accepted and recorded but the cloud is processed in one piece — which also
removes the tile-edge artefacts of the reference.
"""
from __future__ import annotations

import argparse
import datetime
import json
import os
import warnings

from .datahandler import DataHandler
from .utilities import timeit, trace
from .vapc_tools import lasz_to_ply, use_tool

__version__ = "0.1.0"


def get_version():
    return __version__


def _merge_inputs(files, out_path):
    """Several input files -> one LAS file (what the reference's las_merge does before voxelising a list)."""
    dh = DataHandler(list(files))
    dh.load_las_files()
    dh.save_as_las(out_path)
    return out_path


@trace
@timeit
def do_vapc_on_files(file, out_dir, voxel_size, vapc_command, tile=False, reduce_to="closest_to_center_of_gravity", save_as=".laz", doc=True):
    stamp = datetime.datetime.now().strftime("%Y_%m_%d_%H-%M-%S")
    config = vapc_command
    config["vapc_version"] = get_version()
    config["file"] = str(file)
    config["out_dir"] = str(out_dir)
    config["voxel_size"] = voxel_size
    config["reduce_to"] = reduce_to
    config["tile"] = tile
    config["save_as"] = save_as
    os.makedirs(out_dir, exist_ok=True)
    outfile = os.path.join(out_dir, f"{vapc_command['tool']}_{stamp}.las")
    outconfig = os.path.join(out_dir, f"{vapc_command['tool']}_{stamp}_cfg.json")
    # the tool's own reduction setting wins when present (main.py:149-150 reads vapc_command["reduce_to"])
    tool_reduce = vapc_command.get("reduce_to", reduce_to)
    infile, merged = file, None
    if isinstance(file, (list, tuple)) and tile:
        merged = infile = _merge_inputs(file, os.path.join(out_dir, "merged_las.las"))
    use_tool(vapc_command["tool"], infile, outfile, voxel_size=voxel_size, args=vapc_command["args"], reduce_to=tool_reduce)
    if merged:
        os.remove(merged)
    if doc:
        with open(outconfig, "w") as cfg:
            json.dump(config, cfg, indent=4)
    if save_as == ".las":
        return outfile
    if save_as == ".laz":
        laz_out = outfile.replace(".las", ".laz")
        try:
            dh = DataHandler(outfile)
            dh.load_las_files()
            dh.save_as_las(laz_out)
        except NotImplementedError:
            warnings.warn("no LAZ backend available (laspy[lazrs] is not installed): the result stays uncompressed at " + outfile)
            return outfile
        os.remove(outfile)
        return laz_out
    if save_as == ".ply":
        ply_out = outfile.replace(".las", ".ply")
        lasz_to_ply(infile=outfile, outfile=ply_out, voxel_size=voxel_size, shift_to_center=False)
        os.remove(outfile)
        return ply_out
    raise ValueError(f"save_as must be '.laz', '.las' or '.ply', not {save_as!r}")


def parse_args(argv=None):
    parser = argparse.ArgumentParser(description="Use Vapc.")
    parser.add_argument("config_file", help="Configuration of the Vapc command")
    return parser.parse_args(argv)


def load_config(config_file):
    with open(config_file, "r") as fh:
        return json.load(fh)


def main(argv=None):
    args = parse_args(argv)
    config = load_config(args.config_file)
    return do_vapc_on_files(file=config["infile"], out_dir=config["outdir"], voxel_size=config["voxel_size"], vapc_command=config["vapc_command"],
                            tile=config["tile"], reduce_to=config["reduce_to"], save_as=config["save_as"])


if __name__ == "__main__":
    main()
