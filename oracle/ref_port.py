"""CPU port of the reference's pandas data flow for the benchmarked path — used ONLY as the timed
CPU baseline (`bench.py` `cpu_baseline` / `--impl reference`) and checked against the golden fixtures
in tests/test_oracle_golden.py.  TEST / MEASUREMENT INFRASTRUCTURE, never imported by vapc_b200.

Unlike oracle/vapc_oracle.py (which is organised for checking), this module keeps the reference's
*computational structure* so that its run time stands for the reference's: one pandas DataFrame with a
row per point, every per-voxel value broadcast back to the points with groupby(...).transform, six
temporary product columns and ten transform passes for the covariance (vapc.py:947-1006), and ONE PYTHON
CALL PER VOXEL (DataFrame.cov -> np.linalg.eig) for the eigenvalues followed by a merge (vapc.py:1055-1089)
— the step that takes 99 % of the reference's full-attribute-set time.  The reference is
single-threaded; so is this.
"""
import numpy as np
import pandas as pd

KEYS = ["voxel_x", "voxel_y", "voxel_z"]


def voxelize(df, voxel_size, origin=(0, 0, 0)):
    """vapc.py:198-201"""
    for i, dim in enumerate(["X", "Y", "Z"]):
        df[f"voxel_{dim.lower()}"] = np.floor((df[dim] - origin[i]) / voxel_size).astype(int)
    return df


def point_count(df):
    """vapc.py:724-725"""
    df["point_count"] = df.groupby(KEYS)["X"].transform("size")
    return df


def point_density(df, voxel_size):
    """vapc.py:740"""
    df["point_density"] = df["point_count"] / (voxel_size ** 3)
    return df


def center_of_gravity(df):
    """vapc.py:843-850"""
    cog = df.groupby(KEYS, sort=False)[["X", "Y", "Z"]].transform("mean")
    cog.columns = ["cog_x", "cog_y", "cog_z"]
    df[["cog_x", "cog_y", "cog_z"]] = cog
    return df


def std_of_cog(df):
    """vapc.py:877-884"""
    std = df.groupby(KEYS, sort=False)[["X", "Y", "Z"]].transform("std")
    std.columns = ["std_x", "std_y", "std_z"]
    df[["std_x", "std_y", "std_z"]] = std
    return df


def covariance_matrix(df):
    """vapc.py:947-1039: product columns, ten groupby.transform passes, raw-moment formula."""
    for a, b in (("X", "X"), ("Y", "Y"), ("Z", "Z"), ("X", "Y"), ("X", "Z"), ("Y", "Z")):
        df[a + b if a != b else a + "2"] = df[a] * df[b]
    grouped = df.groupby(KEYS)
    df["n"] = grouped["X"].transform("size")
    for col, src in (("sum_x", "X"), ("sum_y", "Y"), ("sum_z", "Z"), ("sum_x2", "X2"), ("sum_y2", "Y2"), ("sum_z2", "Z2"),
                     ("sum_xy", "XY"), ("sum_xz", "XZ"), ("sum_yz", "YZ")):
        df[col] = grouped[src].transform("sum")
    n = df["n"]
    for name, sab, sa, sb in (("cov_xx", "sum_x2", "sum_x", "sum_x"), ("cov_yy", "sum_y2", "sum_y", "sum_y"), ("cov_zz", "sum_z2", "sum_z", "sum_z"),
                              ("cov_xy", "sum_xy", "sum_x", "sum_y"), ("cov_xz", "sum_xz", "sum_x", "sum_z"), ("cov_yz", "sum_yz", "sum_y", "sum_z")):
        df[name] = np.where(n > 1, (df[sab] - (df[sa] * df[sb]) / n) / (n - 1), np.nan)
    df["cov_yx"], df["cov_zx"], df["cov_zy"] = df["cov_xy"], df["cov_xz"], df["cov_yz"]
    df.drop(columns=["X2", "Y2", "Z2", "XY", "XZ", "YZ", "n", "sum_x", "sum_y", "sum_z", "sum_x2", "sum_y2", "sum_z2", "sum_xy", "sum_xz",
                     "sum_yz"], inplace=True)
    cov = ["cov_xx", "cov_xy", "cov_xz", "cov_yx", "cov_yy", "cov_yz", "cov_zx", "cov_zy", "cov_zz"]
    return df[cov + [c for c in df.columns if c not in cov]]


def eigenvalues(df):
    """vapc.py:1055-1089: one Python call per voxel, then a left merge."""

    def per_voxel(g):
        c = g[["X", "Y", "Z"]].cov()
        if c.isna().any().any():
            return np.array([np.nan, np.nan, np.nan])
        w, _ = np.linalg.eig(c)
        return w[w.argsort()[::-1]].flatten()

    eig = df.groupby(KEYS).apply(per_voxel)
    eig_df = pd.DataFrame(eig.values.tolist(), index=eig.index, columns=["Eigenvalue_1", "Eigenvalue_2", "Eigenvalue_3"]).reset_index()
    return df.merge(eig_df, how="left", on=KEYS)


def geometric_features(df):
    """vapc.py:1112-1145"""
    l1, l2, l3 = df["Eigenvalue_1"], df["Eigenvalue_2"], df["Eigenvalue_3"]
    total = l1 + l2 + l3
    for i, l in enumerate((l1, l2, l3), 1):
        df[f"Eigenvalue_{i}_n"] = l / total

    def safe_log(x):
        return np.where(x > 0, np.log(x), 0)

    df["Sum_of_Eigenvalues"] = l1 + l2 + l3
    df["Omnivariance"] = (l1 * l2 * l3) ** (1 / 3)
    df["Eigenentropy"] = -(df["Eigenvalue_1_n"] * safe_log(df["Eigenvalue_1_n"]) + df["Eigenvalue_2_n"] * safe_log(df["Eigenvalue_2_n"])
                           + df["Eigenvalue_3_n"] * safe_log(df["Eigenvalue_3_n"]))
    df["Anisotropy"] = (l1 - l3) / l1
    df["Planarity"] = (l2 - l3) / l1
    df["Linearity"] = (l1 - l2) / l1
    df["Surface_Variation"] = l3 / df["Sum_of_Eigenvalues"]
    df["Sphericity"] = l3 / l1
    return df


def full_attribute_set(x, y, z, voxel_size, origin=(0, 0, 0), with_eigen=True):
    """BASELINE config 2 on the CPU: voxelize + point_count + point_density + center_of_gravity + std_of_cog +
    covariance_matrix (+ eigenvalues + geometric_features).  Returns the per-point frame like the reference."""
    import warnings

    warnings.filterwarnings("ignore")
    df = pd.DataFrame({"X": x, "Y": y, "Z": z})
    df = voxelize(df, voxel_size, origin)
    df = point_count(df)
    df = point_density(df, voxel_size)
    df = center_of_gravity(df)
    df = std_of_cog(df)
    df = covariance_matrix(df)
    if with_eigen:
        df = eigenvalues(df)
        df = geometric_features(df)
    return df
