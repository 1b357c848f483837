"""No-op pyplot: the hot path never plots; only calc_log_e_rate's post-processing does (CL:546-565)."""
import numpy as _np


def hist(data, bins=10, **_kw):
    vals, edges = _np.histogram(_np.asarray(data, dtype=float), bins=bins)
    return vals.astype(float), edges, None


def _noop(*_a, **_k):
    return None


ylabel = xlabel = plot = title = savefig = show = close = legend = bar = xticks = figure = _noop
