"""TEST INFRASTRUCTURE — thin adaptor over the ICP oracle (oracle/icp_oracle.py) for the downstream check of
SURVEY.md §8c/§8d ("ICP fitness/RMSE on the produced pyramids within 1e-3"): the SAME ICP runs on the oracle
pyramid and on the GPU pyramid."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle"))
import icp_oracle  # noqa: E402


def icp_point_to_point(src, tgt, init, max_corr, max_iter, rel_fitness=1e-6, rel_rmse=1e-6):
    r = icp_oracle.registration_icp(src, tgt, max_corr, init, rel_fitness, rel_rmse, max_iter)
    return r["transformation"], r["fitness"], r["inlier_rmse"]


def multiscale_icp(pyr_src, pyr_tgt, max_corrs, max_iters):
    """pyr_* = [level 0 (full), level 1, ...]; coarsest first, chained transformations."""
    rs = icp_oracle.multiscale(pyr_src, pyr_tgt, max_corrs, max_iters)
    return rs[-1]["transformation"], [(r["fitness"], r["inlier_rmse"]) for r in rs]
