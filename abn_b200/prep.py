"""Data preparation in front of the score-cache hot path.

Mirrors what the R drivers do to ``data.df`` before their per-candidate loops:

* ``R/build_score_cache_mle.R:398-415``  centre/scale Gaussian columns, binomial factor -> 0/1,
  multinomial -> factor; ``:563-599`` one-hot expansion of multinomial columns (all levels kept)
* ``R/build_score_cache_bayes.R:40-60``  binary factor codes - 1, centre Gaussians, all double

The result is the column-major FP64 matrix + per-node column ranges that ``abn_dataset``
(``include/abn_b200.h``) describes.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import _capi as capi


@dataclass
class Column:
    """One data-frame column: numeric values, or a factor (1-based codes + levels)."""
    name: str
    values: np.ndarray | None = None
    codes: np.ndarray | None = None
    levels: list | None = None

    @property
    def is_factor(self) -> bool:
        return self.codes is not None


@dataclass
class Prepared:
    names: list
    dists: list              # family name per node
    X: np.ndarray            # (n, n_cols) Fortran order float64
    node_family: np.ndarray  # int32 [d]
    node_col0: np.ndarray    # int32 [d+1]
    col_names: list = field(default_factory=list)
    centre: bool = True

    @property
    def n(self) -> int:
        return self.X.shape[0]

    @property
    def d(self) -> int:
        return len(self.names)


def as_columns(data_df) -> list:
    """Accepts a pandas.DataFrame, a dict name -> array / (codes, levels), or a list of Column."""
    if isinstance(data_df, list) and all(isinstance(c, Column) for c in data_df):
        return data_df
    cols = []
    try:
        import pandas as pd
    except Exception:  # pragma: no cover
        pd = None
    if pd is not None and isinstance(data_df, pd.DataFrame):
        for name in data_df.columns:
            s = data_df[name]
            if isinstance(s.dtype, pd.CategoricalDtype):
                cols.append(Column(str(name), codes=np.asarray(s.cat.codes, dtype=np.int32) + 1,
                                   levels=[str(x) for x in s.cat.categories]))
            elif s.dtype == object or s.dtype == bool or str(s.dtype).startswith("string"):
                cols.append(_factor(str(name), np.asarray(s)))
            else:
                cols.append(Column(str(name), values=np.asarray(s)))
        return cols
    for name, v in dict(data_df).items():
        if isinstance(v, tuple) and len(v) == 2:
            cols.append(Column(str(name), codes=np.asarray(v[0], dtype=np.int32), levels=list(v[1])))
        else:
            a = np.asarray(v)
            if a.dtype.kind in "OUSb":
                cols.append(_factor(str(name), a))
            else:
                cols.append(Column(str(name), values=a))
    return cols


def _factor(name, a) -> Column:
    """R's factor(x): levels = sorted unique values."""
    lv, inv = np.unique(a.astype(str), return_inverse=True)
    return Column(name, codes=inv.astype(np.int32) + 1, levels=list(lv))


def _refactor(col: Column):
    """factor(f): drop unused levels, keep level order. Returns 0-based codes, levels."""
    if col.is_factor:
        used = np.unique(col.codes)
        remap = np.zeros(max(len(col.levels), int(used.max())) + 1, dtype=np.int64)
        remap[used] = np.arange(len(used))
        return remap[col.codes], [col.levels[u - 1] for u in used]
    lv, inv = np.unique(col.values, return_inverse=True)
    return inv, [str(x) for x in lv]


def r_mean(x: np.ndarray) -> float:
    """R's mean(): long-double sum with one refinement pass (src/main/summary.c)."""
    xl = x.astype(np.longdouble)
    n = len(xl)
    s = xl.sum() / n
    s = s + (xl - s).sum() / n
    return float(s)


def r_sd(x: np.ndarray) -> float:
    """R's sd(): sqrt of the two-pass long-double variance with n-1 (src/library/stats/src/cov.c)."""
    xl = x.astype(np.longdouble)
    n = len(xl)
    m = xl.sum() / n
    m = m + (xl - m).sum() / n
    m = np.longdouble(float(m))
    v = ((xl - m) ** 2).sum() / (n - 1)
    return float(np.sqrt(np.float64(v)))


def prepare(data_df, data_dists: dict, method: str, centre: bool = True) -> Prepared:
    cols = as_columns(data_df)
    names = [c.name for c in cols]
    dists = [str(data_dists[nm]) for nm in names]
    if not any(d == "gaussian" for d in dists):
        centre = False  # R/build_score_cache.R:668-670
    if method == "bayes" and any(d == "multinomial" for d in dists):
        raise ValueError("Multinomial nodes are not yet implemented for method 'bayes'. Try with method='mle'.")
    blocks, fam, col0, col_names = [], [], [0], []
    for c, dist in zip(cols, dists):
        if dist not in capi.FAMILY_CODE:
            raise ValueError(f"unknown distribution {dist!r} for node {c.name!r}")
        if dist == "gaussian":
            x = np.asarray(c.values, dtype=np.float64)
            if centre:
                x = (x - r_mean(x)) / r_sd(x)
            blk, nm = x[:, None], [c.name]
        elif dist == "binomial":
            if c.is_factor:
                if method == "bayes":
                    x = c.codes.astype(np.float64) - 1.0  # as.numeric(f) - 1, _bayes.R:43
                else:
                    x = _refactor(c)[0].astype(np.float64)  # as.numeric(factor(x)) - 1, _mle.R:410
            else:
                v = np.asarray(c.values)
                if method == "mle" and v.dtype.kind in "iu":
                    # not class "numeric" in R: as.numeric(factor(x)) - 1, _mle.R:408-410 (codes 1/2 or -1/1 become 0/1)
                    x = _refactor(c)[0].astype(np.float64)
                else:
                    x = v.astype(np.float64)
            # check.valid.data (R/abn-internal.R:277-290): binary with both cases observed; a numeric column must
            # already be coded 0/1 - anything else would be scored silently with y outside {0, 1}
            uniq = np.unique(x)
            if len(uniq) < 2:
                raise ValueError(f"{c.name} is invalid as it must be binary with both cases being observed.")
            if len(uniq) > 2:
                raise ValueError(f"{c.name} is invalid as it must be binary. Multi-category variables should be split "
                                 "into separate binary variables or defined as multinomial distributed.")
            if not np.array_equal(uniq, [0.0, 1.0]):
                raise ValueError(f"{c.name} is invalid - a numeric binomial column must be coded 0/1 (or be a factor)")
            blk, nm = x[:, None], [c.name]
        elif dist == "poisson":
            blk, nm = np.asarray(c.values, dtype=np.float64)[:, None], [c.name]
        else:  # multinomial: model.matrix(~ -1 + factor(x)), _mle.R:594-597
            codes, levels = _refactor(c)
            blk = np.zeros((len(codes), len(levels)), dtype=np.float64)
            blk[np.arange(len(codes)), codes] = 1.0
            nm = [f"{c.name}{lv}" for lv in levels]
        blocks.append(blk)
        fam.append(capi.FAMILY_CODE[dist])
        col0.append(col0[-1] + blk.shape[1])
        col_names += nm
    X = np.asfortranarray(np.concatenate(blocks, axis=1))
    return Prepared(names, dists, X, np.asarray(fam, dtype=np.int32), np.asarray(col0, dtype=np.int32),
                    col_names, centre)
