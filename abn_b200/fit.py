"""fitAbn(method = "mle") for ONE given DAG on the score-cache kernels - SURVEY.md 8f rank 3, first cut.

The reference's ``fitAbn.mle`` (R/fitabn_mle.R:12-230) regresses every node on its parents with the same
IRLS functions the score cache uses (``regressionLoop`` :240-840) and returns, per node, the log-likelihood,
AIC / BIC / MDL, the coefficients and their standard errors. This mirror covers the part of it that runs on
the B200 kernels: Gaussian, binomial and Poisson children without grouping. Not built: multinomial children
(``nnet::multinom`` with its own standard errors) and ``group.var`` (mixed models). ``method = "bayes"`` returns
the modes and log marginal likelihoods only (``compute.fixed = FALSE``), no marginal densities.

Differences to the cache path, all as in the reference: binomial children are fitted WITHOUT the bias reduction
first (``irls_binomial_cpp_fast``, :585) and fall back to the bias-reduced fit when that fails (:586-594), and
the fits follow the reference's IRLS pass for pass (``ABN_FLAG_EXACT_PATH``) because coefficients and standard
errors - unlike scores - are only as converged as the last step.
"""
from __future__ import annotations

import numpy as np

from . import _capi as capi
from . import cache as _cache
from . import prep as _prep


def fitAbn(dag, data_df, data_dists, method="mle", control=None, centre=True, device=0):
    """``fitAbn(dag, data.df, data.dists, method = "mle")``. ``dag`` is a d x d 0/1 matrix, row i = parents of node i
    (the reference's orientation, R/fitabn.R), or a dict ``{child: [parents]}``. Returns a dict with the
    reference's fields: ``mliknode, mlik, aicnode, aic, bicnode, bic, mdlnode, mdl, df, sse, mse, coef, Stderror, method``
    plus ``error.code`` (0 ok / 1 not converged / 2 non-finite, per node)."""
    if method not in ("mle", "bayes"):
        raise ValueError("`method` unknown.")
    cols = _prep.as_columns(data_df)
    names = [c.name for c in cols]
    if set(data_dists) != set(names):
        raise ValueError("data.dists and data.df have different variable names")
    d = len(names)
    if isinstance(dag, dict):
        m = np.zeros((d, d), dtype=np.int32)
        for ch, ps in dag.items():
            for p in ps:
                m[names.index(ch), names.index(p)] = 1
        dag = m
    dag = np.asarray(dag)
    if dag.shape != (d, d) or (np.diag(dag) != 0).any():
        raise ValueError("`dag` must be a d x d 0/1 matrix with an empty diagonal")
    if method == "bayes":
        return _fit_bayes(dag, names, data_df, data_dists, control, centre, device)
    P = _prep.prepare(data_df, data_dists, "mle", centre=centre)
    multi_children = [names[i] for i in range(d) if P.node_family[i] == capi.MULTINOMIAL]
    width = max(1, int(dag.sum(axis=1).max()))
    children = np.arange(d, dtype=np.int32)
    parents = np.full((d, width), -1, dtype=np.int32)
    for i in range(d):
        ps = np.flatnonzero(dag[i])
        parents[i, :len(ps)] = ps
    ncols = P.node_col0[1:] - P.node_col0[:-1]
    stride = 2 + int(max((ncols[np.flatnonzero(dag[i])].sum() for i in range(d)), default=0))

    ctrl = dict(_cache._check_control(control, "mle"))
    ctrl.update({"exact.path": True, "device": device})
    with _cache.DeviceData(P, device) as dd:
        plain = dd.score(children, parents, _cache.to_c_control({**ctrl, "bias.reduction": False}, "mle", d),
                         coef_stride=stride, se=True)
        # R/fitabn_mle.R:585-594: the bias-reduced IRLS replaces a binomial fit that failed or came back NA
        redo = np.flatnonzero((P.node_family == capi.BINOMIAL) & ((plain["status"] != 0) | ~np.isfinite(plain["mlik"])))
        if len(redo):
            br = dd.score(children[redo], parents[redo], _cache.to_c_control(ctrl, "mle", d), coef_stride=stride, se=True)
            for k in ("mlik", "aic", "bic", "mdl", "status", "coef", "se", "sse"):
                plain[k][redo] = br[k]

    out = {"mliknode": {}, "aicnode": {}, "bicnode": {}, "mdlnode": {}, "df": {}, "sse": {}, "mse": {}, "coef": {},
           "Stderror": {}, "error.code": {}}
    for i, nm in enumerate(names):
        if nm in multi_children:
            for k in ("mliknode", "aicnode", "bicnode", "mdlnode"):
                out[k][nm] = plain[{"mliknode": "mlik", "aicnode": "aic", "bicnode": "bic", "mdlnode": "mdl"}[k]][i]
            out["df"][nm], out["coef"][nm], out["Stderror"][nm] = np.nan, None, None
            out["sse"][nm] = out["mse"][nm] = np.nan
            out["error.code"][nm] = int(plain["status"][i])
            continue
        ps = np.flatnonzero(dag[i])
        pcols = [c for p in ps for c in range(P.node_col0[p], P.node_col0[p + 1])]
        any_multi = any(P.node_family[p] == capi.MULTINOMIAL for p in ps)
        if any_multi:      # no intercept, one coefficient per level column (R/fitabn_mle.R:546-548, :806-808)
            labels = [P.col_names[c] for c in pcols]
            sl = slice(1, 1 + len(pcols))
        else:
            labels = [f"{nm}|intercept"] + [names[p] for p in ps]
            sl = slice(0, 1 + len(pcols))
        out["mliknode"][nm] = float(plain["mlik"][i])
        out["aicnode"][nm] = float(plain["aic"][i])
        out["bicnode"][nm] = float(plain["bic"][i])
        out["mdlnode"][nm] = float(plain["mdl"][i])
        out["df"][nm] = P.n - (len(pcols) + 1)                         # :711
        out["sse"][nm] = float(plain["sse"][i])                        # :712
        out["mse"][nm] = out["sse"][nm] / out["df"][nm]                # :713
        out["coef"][nm] = dict(zip(labels, plain["coef"][i][sl]))
        out["Stderror"][nm] = dict(zip(labels, plain["se"][i][sl]))
        out["error.code"][nm] = int(plain["status"][i])
    out["mlik"] = float(np.nansum(list(out["mliknode"].values())))    # :213-215
    out["aic"] = float(np.nansum(list(out["aicnode"].values())))
    out["bic"] = float(np.nansum(list(out["bicnode"].values())))
    out["mdl"] = float(np.nansum(list(out["mdlnode"].values())))
    out["method"] = "mle"
    out["dag"] = dag
    return out


_ERROR_DESC = {0: "success", 1: "warning: mode results may be unreliable (optimiser terminated unusually)",
               2: "error - logscore is NA - model could not be fitted"}  # R/fitabn_bayes.R:172-176


def _fit_bayes(dag, names, data_df, data_dists, control, centre, device):
    """``fitAbn(method = "bayes", compute.fixed = FALSE)`` (R/fitabn_bayes.R:60-180, 536-565): per node the Laplace
    log marginal likelihood and the posterior modes ``fit_single_node`` leaves behind (src/fit_single_node.c:182-194),
    named ``child|(Intercept)``, ``child|parent`` and, for a Gaussian node, ``child|precision``. Marginal densities
    (``compute.fixed = TRUE``) and ``mse`` are not built."""
    d = len(names)
    P = _prep.prepare(data_df, data_dists, "bayes", centre=centre)
    width = max(1, int(dag.sum(axis=1).max()))
    children = np.arange(d, dtype=np.int32)
    parents = np.full((d, width), -1, dtype=np.int32)
    for i in range(d):
        ps = np.flatnonzero(dag[i])
        parents[i, :len(ps)] = ps
    ctrl = dict(_cache._check_control(control, "bayes"))
    ctrl["device"] = device
    with _cache.DeviceData(P, device) as dd:
        res = dd.score(children, parents, _cache.to_c_control(ctrl, "bayes", d), coef_stride=width + 2)
    out = {"modes": {}, "error.code": {}, "error.code.desc": {}, "hessian.accuracy": {}, "mliknode": {}, "used.INLA": {}}
    for i, nm in enumerate(names):
        ps = np.flatnonzero(dag[i])
        labels = [f"{nm}|(Intercept)"] + [f"{nm}|{names[p]}" for p in ps]
        vals = list(res["coef"][i][:1 + len(ps)])
        if P.node_family[i] == capi.GAUSSIAN:
            labels.append(f"{nm}|precision")
            vals.append(res["coef"][i][1 + len(ps)])
        out["modes"][nm] = dict(zip(labels, (float(v) for v in vals)))
        out["error.code"][nm] = int(res["status"][i])
        out["error.code.desc"][nm] = _ERROR_DESC.get(int(res["status"][i]), str(int(res["status"][i])))
        out["hessian.accuracy"][nm] = 0.0
        out["mliknode"][nm] = float(res["mlik"][i])
        out["used.INLA"][nm] = False
    out["mlik"] = float(sum(out["mliknode"].values()))
    out["coef"] = {nm: list(m.values()) for nm, m in out["modes"].items()}   # modes2coefs
    out["method"] = "bayes"
    out["dag"] = dag
    return out
