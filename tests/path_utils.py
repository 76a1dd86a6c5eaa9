"""Geometry helpers of the rounded-path comparisons (tests only)."""
import numpy as np


def polyline(x_v_rounded, path):
    """waypoints of a rounded path: the segments x_v[:2] -> x_v[2:] of its vertices in path order"""
    pts = []
    for v in path:
        x = np.asarray(x_v_rounded[v], float)
        pts += [x[:2], x[2:]]
    return np.array(pts)


def _sample(P, per=64):
    t = np.linspace(0.0, 1.0, per)[:, None]
    return np.concatenate([a + t * (b - a) for a, b in zip(P[:-1], P[1:])] + [P[-1:]])


def _dist_topolyline(X, Q):
    """exact distance of every point of X to the polyline Q"""
    a, b = Q[:-1][None, :, :], Q[1:][None, :, :]
    ab = b - a
    den = np.maximum(np.sum(ab * ab, axis=2), 1e-300)
    t = np.clip(np.sum((X[:, None, :] - a) * ab, axis=2) / den, 0.0, 1.0)
    return np.linalg.norm(X[:, None, :] - (a + t[..., None] * ab), axis=2).min(axis=1)


def hausdorff(P, Q):
    """symmetric Hausdorff distance of two polylines (sampled points of one against the exact segments of the other)"""
    return float(max(_dist_topolyline(_sample(P), Q).max(), _dist_topolyline(_sample(Q), P).max()))


def gold_path(As, d, keys, tag):
    """the stored rounded result of the reference as (ordered path, x_v_rounded dict, final cost)"""
    from gcs_admm_b200.graph import build_graph
    on = {k for k, y in zip(keys, d[f"{tag}_y_v_rounded"]) if y > 0.5}
    xr = {k: np.asarray(x, float) for k, x in zip(keys, d[f"{tag}_x_v_rounded"])}
    path, cur = ['s'], 's'
    while cur != 't':            # order the stored vertex set by continuity x_v[2:] == x_w[:2]
        nxt = [w for w in on if w not in path and np.max(np.abs(xr[cur][2:] - xr[w][:2])) < 1e-5]
        assert nxt, (cur, on)
        cur = min(nxt, key=lambda w: np.linalg.norm(xr[w][:2] - xr[w][2:]) == 0.0)
        path.append(cur)
    return path, xr, float(sum(np.linalg.norm(xr[v][:2] - xr[v][2:]) for v in on))
