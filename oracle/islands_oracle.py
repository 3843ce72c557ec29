"""CPU oracle for the island analytics (csrc/islands.cu)  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

The reference has no implementation of the analysis its README describes (README.md:34-36: inspect the returned
states "for the theorized islands"); this numpy restatement of the definition in include/glom_b200.h is therefore the
only checker ("parity unpinned": there is nothing in the reference to pin it on).  Only tests/ may import it."""
import numpy as np


def islands(states, side_h, side_w, threshold):
    """states (..., n, L, d) -> dict of cos_right, cos_down, agreement (..., L, n), labels int32, num_islands (..., L)."""
    x = np.asarray(states, dtype=np.float64)
    *lead, n, L, d = x.shape
    assert n == side_h * side_w
    g = np.moveaxis(x, -2, -3).reshape(*lead, L, side_h, side_w, d)          # (..., L, h, w, d)
    nrm = np.sqrt((g * g).sum(-1))
    cr = np.zeros(g.shape[:-1])
    cd = np.zeros(g.shape[:-1])
    cr[..., :, :-1] = (g[..., :, :-1, :] * g[..., :, 1:, :]).sum(-1) / np.maximum(nrm[..., :, :-1] * nrm[..., :, 1:], 1e-12)
    cd[..., :-1, :] = (g[..., :-1, :, :] * g[..., 1:, :, :]).sum(-1) / np.maximum(nrm[..., :-1, :] * nrm[..., 1:, :], 1e-12)
    s = np.zeros_like(cr)
    c = np.zeros_like(cr)
    s[..., :, :-1] += cr[..., :, :-1]; c[..., :, :-1] += 1          # right neighbour
    s[..., :, 1:] += cr[..., :, :-1]; c[..., :, 1:] += 1            # left neighbour
    s[..., :-1, :] += cd[..., :-1, :]; c[..., :-1, :] += 1          # lower neighbour
    s[..., 1:, :] += cd[..., :-1, :]; c[..., 1:, :] += 1            # upper neighbour
    agreement = np.where(c > 0, s / np.maximum(c, 1), 1.0)
    flat_r = cr.reshape(-1, side_h, side_w)
    flat_d = cd.reshape(-1, side_h, side_w)
    labels = np.empty((flat_r.shape[0], n), dtype=np.int32)
    counts = np.empty(flat_r.shape[0], dtype=np.int32)
    for k in range(flat_r.shape[0]):                                  # union-find per (slab, level)
        parent = list(range(n))

        def find(a):
            while parent[a] != a:
                parent[a] = parent[parent[a]]
                a = parent[a]
            return a
        for h in range(side_h):
            for w in range(side_w):
                i = h * side_w + w
                if w + 1 < side_w and flat_r[k, h, w] >= threshold:
                    a, b = find(i), find(i + 1)
                    parent[max(a, b)] = min(a, b)
                if h + 1 < side_h and flat_d[k, h, w] >= threshold:
                    a, b = find(i), find(i + side_w)
                    parent[max(a, b)] = min(a, b)
        roots = [find(i) for i in range(n)]
        labels[k] = roots
        counts[k] = len(set(roots))
    shp = (*lead, L, n)
    return dict(cos_right=cr.reshape(shp), cos_down=cd.reshape(shp), agreement=agreement.reshape(shp),
                labels=labels.reshape(shp), num_islands=counts.reshape(*lead, L))
