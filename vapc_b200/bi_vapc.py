"""`BiTemporalVapc` — drop-in for the reference's two-epoch change detection
(src/vapc/bi_vapc.py:11-207) on top of the B200 engine.

The voxel join (inner join + both anti-joins) and the centroid-delta / Mahalanobis / chi-square
classification run on the GPU (vapc_join_sorted_keys, vapc_mahalanobis); the merged frame keeps
the reference's column names (`*_x` = epoch 1, `*_y` = epoch 2, `distance`,
`mahalanobi_significance`, `p_value`, `change_type`).  Additionally `d1` / `d2` (the two Mahalanobis
distances, internal temporaries in the reference) are kept in `self.mahalanobis_d`.
"""
from __future__ import annotations

import numpy as np
import pandas as pd
import torch

from . import engine as E
from .utilities import timeit, trace

_COV6 = ["cov_xx", "cov_xy", "cov_xz", "cov_yy", "cov_yz", "cov_zz"]


class BiTemporalVapc:
    def __init__(self, vapcs: list):
        if len(vapcs) != 2:
            raise ValueError("BiTemporalVapc requires exactly two dataframes.")
        self.vapcs = vapcs
        self.merged_df = None
        self.final_df = None
        self.distance_computed = False
        self.mahalanobis_computed = False
        self.prepared_for_bi_temporal_analysis = False
        self.single_occupation_computed = False
        self.mahalanobis_d = None
        self._only = None

    @trace
    @timeit
    def compute_voxels_occupied_in_single_epoch(self):
        """Voxels only in epoch 1 -> change_type 3, only in epoch 2 -> 4 (bi_vapc.py:25-36)."""
        if self.prepared_for_bi_temporal_analysis is False:
            raise ValueError("Data should be prepared for bi_temporal analysis using 'prepare_data_for_mahalanobis_distance()'")
        if self._only is None:
            self._join()
        only1, only2 = self._only
        a = self.vapcs[0].df.iloc[only1].copy()
        b = self.vapcs[1].df.iloc[only2].copy()
        a["change_type"] = 3  # disappearing
        b["change_type"] = 4  # appearing
        self.df_appear_disappear = pd.concat([a, b])
        self.single_occupation_computed = True

    @trace
    @timeit
    def prepare_data_for_mahalanobis_distance(self):
        """Per epoch: centroid + covariance, reduced to one centroid row per voxel (bi_vapc.py:40-47)."""
        for vapc in self.vapcs:
            vapc.return_at = "center_of_gravity"
            vapc.compute_center_of_gravity()
            vapc.compute_covariance_matrix()
            vapc.reduce_to_voxels()
        self.prepared_for_bi_temporal_analysis = True

    def _join(self):
        for vapc in self.vapcs:
            vapc.voxelize()
            vapc.compute_voxel_index()
        dev = torch.device("cuda", torch.cuda.current_device())
        tabs = []
        for vapc in self.vapcs:
            tabs.append([torch.from_numpy(vapc.df[c].to_numpy(np.int64)).to(dev) for c in ("voxel_x", "voxel_y", "voxel_z")])
        i1, i2, only1, only2 = E.join_voxel_tables(tabs[0], tabs[1])
        self._pairs = (i1.cpu().numpy(), i2.cpu().numpy())
        self._only = (only1.cpu().numpy(), only2.cpu().numpy())

    @trace
    @timeit
    def merge_vapcs_with_same_voxel_index(self):
        """Inner join of the two voxel tables on the voxel triple, suffixes _x / _y, epoch-1 row order
        (bi_vapc.py:78-84)."""
        self._join()
        i1, i2 = self._pairs
        df1, df2 = self.vapcs[0].df, self.vapcs[1].df
        left = df1.iloc[i1].add_suffix("_x")
        right = df2.iloc[i2].add_suffix("_y")
        right.index = left.index
        self.merged_df = pd.concat([left, right], axis=1)

    @trace
    @timeit
    def merge_vapcs_with_closest_point(self):
        """Pairs every epoch-1 row with its nearest epoch-2 centroid where the two are MUTUAL nearest neighbours
        (bi_vapc.py:51-74, 173-186), side by side with suffixes _x / _y and the pair `distance`.  The three KD-tree
        queries of the reference are exact grid searches on the GPU (vapc_cell_nn) over the epochs' own voxel grids
        (one centroid per cell)."""
        df1, df2 = self.vapcs[0].df, self.vapcs[1].df
        cols = ["cog_x", "cog_y", "cog_z"]
        sets = [[d[c].to_numpy(np.float64) for c in cols] for d in (df1, df2)]
        idx1, idx2, dist = E.mutual_nearest_neighbors(sets[0], sets[1], cell_sizes=[float(v.voxel_size) for v in self.vapcs],
                                                      origins=[[float(o) for o in v.origin] for v in self.vapcs])
        df1_reduced = df1.iloc[idx1.cpu().numpy()].copy()
        df2_reduced = df2.iloc[idx2.cpu().numpy()].copy()
        df1_reduced["distance"] = dist.cpu().numpy()
        df1_reduced.reset_index(drop=True, inplace=True)
        df2_reduced.reset_index(drop=True, inplace=True)
        self.merged_df = pd.concat([df1_reduced.add_suffix("_x"), df2_reduced.add_suffix("_y")], axis=1)
        self.merged_df.rename(columns={"distance_x": "distance"}, inplace=True)
        self.distance_computed = True

    def _device_tables(self):
        df = self.merged_df
        dev = torch.device("cuda", torch.cuda.current_device())

        def tab(cols):
            return torch.from_numpy(np.ascontiguousarray(df[cols].to_numpy(np.float64).T)).to(dev)

        cog1 = tab(["cog_x_x", "cog_y_x", "cog_z_x"])
        cog2 = tab(["cog_x_y", "cog_y_y", "cog_z_y"])
        cov_cols_x = sorted(c for c in df.columns if c.startswith("cov_") and c.endswith("_x"))
        cov_cols_y = sorted(c for c in df.columns if c.startswith("cov_") and c.endswith("_y"))
        has_cov = len(cov_cols_x) == 9 and len(cov_cols_y) == 9
        cov1 = tab([c + "_x" for c in _COV6]) if has_cov else None
        cov2 = tab([c + "_y" for c in _COV6]) if has_cov else None
        rows = torch.arange(len(df), dtype=torch.int64, device=dev)
        return cog1, cov1, cog2, cov2, rows, has_cov

    def compute_distance(self):
        """Euclidean distance of the joined centroids (bi_vapc.py:86-88, 198-207)."""
        cog1, cov1, cog2, cov2, rows, has_cov = self._device_tables()
        zeros = torch.zeros((6, len(self.merged_df)), dtype=torch.float64, device=cog1.device)
        res = E.mahalanobis(cog1, cov1 if has_cov else zeros, cog2, cov2 if has_cov else zeros, rows, rows)
        self.merged_df["distance"] = res["distance"].cpu().numpy()
        self.distance_computed = True

    @trace
    @timeit
    def compute_mahalanobis_distance(self, alpha: float = 0.005):
        """Mahalanobis test of the centroid shift against either epoch's covariance (+1e-10 I), p = 1 - chi2
        cdf (3 dof), significance, change_type 0 / 1 / 2 (bi_vapc.py:92-138)."""
        if self.merged_df is None:
            raise ValueError("Run compute_nearest_neighbors() before computing Mahalanobis test.")
        cog1, cov1, cog2, cov2, rows, has_cov = self._device_tables()
        if not has_cov:
            raise ValueError("Incomplete covariance matrix columns.")
        res = E.mahalanobis(cog1, cov1, cog2, cov2, rows, rows, alpha=alpha)
        df = self.merged_df
        df["mahalanobi_significance"] = res["significance"].cpu().numpy().astype(np.int64)
        df["p_value"] = res["p_value"].cpu().numpy()
        df["change_type"] = res["change_type"].cpu().numpy().astype(np.int64)
        self.mahalanobis_d = (res["d1"].cpu().numpy(), res["d2"].cpu().numpy())
        self.merged_df = df
        self.mahalanobis_computed = True

    @trace
    @timeit
    def prepare_data_for_export(self):
        """X, Y, Z (= epoch-1 centroid) + computed columns, plus the appear / disappear rows (bi_vapc.py:142-161)."""
        if self.merged_df is None:
            raise ValueError("Run compute_mahalanobis_distance() before saving to LAS.")
        cols = ["cog_x_x", "cog_y_x", "cog_z_x"]
        if self.distance_computed:
            cols.append("distance")
        if self.mahalanobis_computed:
            cols.extend(["mahalanobi_significance", "p_value", "change_type"])
        self.df = self.merged_df[cols].rename(columns={"cog_x_x": "X", "cog_y_x": "Y", "cog_z_x": "Z"})
        if self.single_occupation_computed is True:
            self.df = pd.concat([self.df, self.df_appear_disappear[["X", "Y", "Z", "change_type"]]])

    @trace
    @timeit
    def save_to_las(self, output_file: str):
        """Save the exported frame to LAS through the DataHandler (bi_vapc.py:163-169)."""
        from .datahandler import DataHandler

        dh = DataHandler("")
        dh.df = self.df
        dh.save_as_las(output_file)


# ---- the reference's module-level helpers (bi_vapc.py:171-207), for callers that import them ------------------------------------
@trace
@timeit
def mutual_nearest_neighbors(df1, df2, coord_cols=["cog_x", "cog_y", "cog_z"]):
    """Row numbers (idx1, idx2) of the MUTUAL nearest neighbours of two frames (bi_vapc.py:173-186): the reference's two KD-tree
    queries as exact grid searches on the GPU (engine.mutual_nearest_neighbors -> vapc_cell_nn).  Three coordinate columns."""
    if len(coord_cols) != 3:
        raise ValueError("mutual_nearest_neighbors searches 3-D coordinates: give three columns")
    sets = [[np.ascontiguousarray(d[c].to_numpy(np.float64)) for c in coord_cols] for d in (df1, df2)]
    idx1, idx2, _ = E.mutual_nearest_neighbors(sets[0], sets[1])
    return idx1.cpu().numpy(), idx2.cpu().numpy()


@trace
@timeit
def merge_with_suffix(df1, df2, suffix1="_x", suffix2="_y"):
    """Two frames side by side, columns suffixed (bi_vapc.py:190-194): frame bookkeeping, nothing to compute."""
    return pd.concat([df1.add_suffix(suffix1), df2.add_suffix(suffix2)], axis=1)


@trace
@timeit
def compute_euclidean_distance(df, coord_prefix="cog"):
    """sqrt(dx^2 + dy^2 + dz^2) between the `<prefix>_?_x` and `<prefix>_?_y` columns of a merged frame (bi_vapc.py:198-207), by the
    same kernel as BiTemporalVapc.compute_distance (vapc_mahalanobis' distance output, explicit round-to-nearest operations in the
    reference's order).  Returns a Series on the frame's index like the reference."""
    E._require_cuda()
    dev = torch.device("cuda", torch.cuda.current_device())

    def tab(s):
        return torch.from_numpy(np.ascontiguousarray(df[[f"{coord_prefix}_{a}_{s}" for a in "xyz"]].to_numpy(np.float64).T)).to(dev)

    c1, c2 = tab("x"), tab("y")
    rows = torch.arange(len(df), dtype=torch.int64, device=dev)
    zeros = torch.zeros((6, len(df)), dtype=torch.float64, device=dev)
    return pd.Series(E.mahalanobis(c1, zeros, c2, zeros, rows, rows)["distance"].cpu().numpy(), index=df.index)
