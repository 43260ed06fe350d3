"""Acceptance metric: per-basin Nash-Sutcliffe efficiency in float64 on the host, following
Hy2DL/aux_functions/functions_evaluation.py:7-50 (offline; SURVEY.md row a13 keeps it on the CPU)."""
from __future__ import annotations

from typing import Dict

import numpy as np


def nse(df_results: Dict[str, "object"], average: bool = True):
    """df_results: {basin: frame-like with 'y_sim' and 'y_obs' columns}. Median over basins if average."""
    scores = []
    for frame in df_results.values():
        sim = np.asarray(frame["y_sim"], dtype=np.float64)
        obs = np.asarray(frame["y_obs"], dtype=np.float64)
        keep = ~np.isnan(sim)
        sim, obs = sim[keep], obs[keep]
        keep = ~np.isnan(obs)
        sim, obs = sim[keep], obs[keep]
        if sim.size > 1:
            scores.append(1.0 - np.sum((sim - obs) ** 2) / np.sum((obs - obs.mean()) ** 2))
        else:
            scores.append(np.nan)
    return np.nanmedian(scores) if average else np.asarray(scores)
