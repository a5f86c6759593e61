"""Offline estimate (CPU, numpy), not part of the product: emulate k_degree's grouping (windows of W sorted points split
by coarse row, candidate stream = 9 stencil rows widened to [cx'min-1, cx'max+1]) and count the exact pair tests that remain
when candidate cells (edge h/sub) are first classified against the group's tight bounding box."""
import sys, numpy as np
sys.path.insert(0, '/root/repo')
from pbnet_b200 import scenes

def analyse(xyz, r, W=128, sub=1, thr=0):
    h = np.float32(r / 2 * (1 + 2.0 ** -7))
    hs = h / sub
    mn = xyz.min(0)
    fs = np.floor((xyz - mn) / hs).astype(np.int64)       # sub-cell coords
    cc = fs // (2 * sub)                                    # coarse coords
    # sort: coarse z,y,x then sub-cell (morton-ish: z,y,x local)
    loc = fs - cc * (2 * sub)
    lkey = (loc[:, 2] * (2 * sub) + loc[:, 1]) * (2 * sub) + loc[:, 0]
    ckey = (cc[:, 2] << 40) | (cc[:, 1] << 20) | cc[:, 0]
    order = np.lexsort((lkey, ckey))
    p = xyz[order].astype(np.float64); cc = cc[order]; ckey = ckey[order]; lkey = lkey[order]
    n = len(p)
    # candidate cells (sub-cells)
    full = ckey * 4096 + lkey
    uk, start, cnt = np.unique(full, return_index=True, return_counts=True)
    lo = np.minimum.reduceat(p, start, axis=0); hi = np.maximum.reduceat(p, start, axis=0)
    cell_ck = ckey[start]; cell_cc = cc[start]
    # coarse cell table
    uck, cstart = np.unique(cell_ck, return_index=True)     # first sub-cell of each coarse cell
    cend = np.append(cstart[1:], len(uk))
    row = ckey >> 20
    r2 = r * r
    T_now = T_new = IN = ncls = 0
    for b in range(0, n, W):
        e = min(n, b + W)
        rows = row[b:e]
        heads = np.nonzero(np.r_[True, rows[1:] != rows[:-1]])[0]
        for gi, g0 in enumerate(heads):
            g1 = heads[gi + 1] if gi + 1 < len(heads) else e - b
            q = p[b + g0:b + g1]
            qlo = q.min(0); qhi = q.max(0)
            cy, cz = cc[b + g0, 1], cc[b + g0, 2]
            cx0, cx1 = cc[b + g0, 0] - 1, cc[b + g1 - 1, 0] + 1
            nq = g1 - g0
            for dz in (-1, 0, 1):
                for dy in (-1, 0, 1):
                    base = ((cz + dz) << 40) | ((cy + dy) << 20)
                    i0 = np.searchsorted(uck, base | max(cx0, 0)); i1 = np.searchsorted(uck, base | cx1, side='right')
                    if cy + dy < 0 or cz + dz < 0 or i1 <= i0: continue
                    s0, s1 = cstart[i0], cend[i1 - 1]
                    c_lo = lo[s0:s1]; c_hi = hi[s0:s1]; c_n = cnt[s0:s1]
                    T_now += nq * int(c_n.sum())
                    gap = np.maximum(0, np.maximum(c_lo - qhi, qlo - c_hi)); far = np.maximum(c_hi - qlo, qhi - c_lo)
                    dmin = (gap ** 2).sum(1); dmax = (far ** 2).sum(1)
                    skip_out = (dmin > r2) & (c_n >= thr); skip_in = (dmax <= r2) & (c_n >= thr)
                    IN += nq * int(c_n[skip_in].sum())
                    T_new += nq * int(c_n[~skip_out & ~skip_in].sum())
                    ncls += len(c_n)
    return dict(n=n, T_now=T_now, T_new=T_new, IN=IN, ncls=ncls)

if __name__ == '__main__':
    sc = scenes.make_scene(22, 150000)
    calls = scenes.class_calls(sc)
    for W, sub, thr in [(128, 1, 0), (128, 2, 0), (64, 1, 0), (64, 2, 0), (64, 2, 4), (32, 2, 0), (64, 4, 0)]:
        tot = {}
        for call in calls:
            res = analyse(call['xyz_shift'], scenes.RADIUS, W, sub, thr)
            for k, v in res.items(): tot[k] = tot.get(k, 0) + v
        n = tot['n']
        print(f"W={W} sub={sub} thr={thr}: tests now {tot['T_now']/n:.0f}/pt -> {tot['T_new']/n:.0f}/pt (+{tot['IN']/n:.0f} counted all-in), cell classifications {tot['ncls']/n:.1f}/pt")
