"""abn_b200 - B200-native score cache for Additive Bayesian Networks (the hot path of furrer-lab/abn).

Public surface mirrors the reference's operator interface for this path:

* :func:`buildScoreCache`, :func:`build_control`  (R/build_score_cache.R)
* :func:`fitAbn` (``method = "mle"``, first cut: R/fitabn_mle.R on the same kernels)
* :class:`DeviceData`, :func:`enumerate_candidates`, :func:`node_defn`  (thin views of the C ABI,
  include/abn_b200.h)
* :mod:`abn_b200.rdata` reads the bundled ``.RData`` / ``.rds`` data sets and caches without R.
"""
from .cache import (AbnCache, DeviceData, build_control, buildScoreCache, enumerate_candidates, node_defn,
                    to_c_control)
from .fit import fitAbn
from .prep import prepare

__all__ = ["AbnCache", "DeviceData", "build_control", "buildScoreCache", "enumerate_candidates", "fitAbn", "node_defn",
           "prepare", "to_c_control"]
__version__ = "0.1.0"
