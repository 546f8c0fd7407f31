"""Synthetic workloads of BASELINE.json (generator fixed in SURVEY.md section 8d).

numpy ``Generator(PCG64(20261019 + config))``; latent z in R^5 ~ N(0, I) per row; per column j:
a_j ~ U(-0.5, 0.5), w_j ~ N(0, 0.4^2 I_5), s_j = a_j + w_j'z.  binomial: Bernoulli(sigmoid(s_j));
Gaussian: s_j + N(0, 1); Poisson: Poisson(exp(0.5 + 0.5 s_j)); multinomial: argmax of L independent
(s_{j,l} + Gumbel).  Correlated columns, no separation, |eta| << 37.
"""
from __future__ import annotations

import numpy as np

CONFIGS = {
    # id: (n, families, max_parents, method)
    3: dict(n=100_000, fams=["binomial"] * 50, max_parents=4, method="mle"),
    4: dict(n=1_000_000, fams=[("gaussian", "binomial", "poisson")[i % 3] for i in range(100)], max_parents=3,
            method="mle"),
    5: dict(n=500_000, fams=[("gaussian", "binomial", "poisson")[i % 3] for i in range(39)] + ["multinomial"],
            max_parents=4, method="mle"),
}


def make_columns(n: int, fams: list, seed: int, levels: int = 4):
    """Returns (names, dict name -> array / (codes, levels), dists)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    z = rng.standard_normal((n, 5))
    data, dists, names = {}, {}, []
    counters = {"gaussian": 0, "binomial": 0, "poisson": 0, "multinomial": 0}
    for fam in fams:
        counters[fam] += 1
        name = f"{fam[0]}{counters[fam]}"
        names.append(name)
        dists[name] = fam
        if fam == "multinomial":
            best, arg = None, None
            for l in range(levels):
                a = rng.uniform(-0.5, 0.5)
                w = rng.normal(0.0, 0.4, 5)
                s = a + z @ w + rng.gumbel(size=n)
                if best is None:
                    best, arg = s, np.zeros(n, dtype=np.int32)
                else:
                    upd = s > best
                    arg[upd] = l
                    best = np.where(upd, s, best)
            data[name] = (arg + 1, [f"L{l}" for l in range(levels)])
            continue
        a = rng.uniform(-0.5, 0.5)
        w = rng.normal(0.0, 0.4, 5)
        s = a + z @ w
        if fam == "binomial":
            data[name] = (rng.random(n) < 1.0 / (1.0 + np.exp(-s))).astype(np.float64)
        elif fam == "gaussian":
            data[name] = s + rng.standard_normal(n)
        else:
            data[name] = rng.poisson(np.exp(0.5 + 0.5 * s)).astype(np.float64)
    return names, data, dists


def make_config(cfg: int, n: int | None = None, n_nodes: int | None = None):
    """Data of BASELINE.json config `cfg` (3, 4 or 5), optionally shrunk for tests."""
    spec = CONFIGS[cfg]
    fams = spec["fams"] if n_nodes is None else spec["fams"][:n_nodes]
    names, data, dists = make_columns(n or spec["n"], fams, 20261019 + cfg)
    return names, data, dists, spec["max_parents"], spec["method"]
