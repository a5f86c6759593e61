"""Offline estimate (CPU, numpy), not part of the product.  Points sorted by coarse cell (x fastest) and then by a Morton
code of `levels` octree levels inside the coarse cell (level 1 = the fine cells of the product).  Query groups = W
consecutive sorted points (split at coarse-row changes), candidate buckets = B consecutive sorted points with tight
bounding boxes.  Counts the exact pair tests left when a bucket is skipped if the boxes prove all-in / all-out."""
import sys, numpy as np
sys.path.insert(0, '/root/repo')
from pbnet_b200 import scenes

def morton_local(loc, levels):
    k = np.zeros(len(loc), dtype=np.int64)
    for l in range(levels - 1, -1, -1):      # most significant level first
        b = (loc >> l) & 1
        k = (k << 3) | (b[:, 2] << 2) | (b[:, 1] << 1) | b[:, 0]
    return k

def analyse(xyz, r, W, B, levels):
    h = np.float32(r / 2 * (1 + 2.0 ** -7))
    mn = xyz.min(0)
    sub = 1 << (levels - 1)
    fs = np.floor((xyz - mn) / (h / sub)).astype(np.int64)
    cc = fs // (2 * sub)
    lkey = morton_local(fs - cc * (2 * sub), levels)
    ckey = (cc[:, 2] << 40) | (cc[:, 1] << 20) | cc[:, 0]
    order = np.lexsort((lkey, ckey))
    p = xyz[order].astype(np.float64); cc = cc[order]; ckey = ckey[order]
    n = len(p)
    nb = (n + B - 1) // B
    bstart = np.arange(nb) * B
    blo = np.minimum.reduceat(p, bstart, axis=0); bhi = np.maximum.reduceat(p, bstart, axis=0)
    uck, cstart = np.unique(ckey, return_index=True)
    cend = np.append(cstart[1:], n)
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
                    j0, j1 = cstart[i0], cend[i1 - 1]
                    T_now += nq * (j1 - j0)
                    b0, b1 = j0 // B, (j1 + B - 1) // B
                    c_lo = blo[b0:b1]; c_hi = bhi[b0:b1]
                    c_n = np.minimum(bstart[b0:b1] + B, j1) - np.maximum(bstart[b0:b1], j0)
                    gap = np.maximum(0, np.maximum(c_lo - qhi, qlo - c_hi)); far = np.maximum(c_hi - qlo, qhi - c_lo)
                    dmin = (gap ** 2).sum(1); dmax = (far ** 2).sum(1)
                    out = dmin > r2; inn = dmax <= r2
                    IN += nq * int(c_n[inn].sum())
                    T_new += nq * int(c_n[~out & ~inn].sum())
                    ncls += len(c_n)
    return dict(n=n, T_now=T_now, T_new=T_new, IN=IN, ncls=ncls)

if __name__ == '__main__':
    sc = scenes.make_scene(int(sys.argv[1]) if len(sys.argv) > 1 else 22, 150000)
    calls = scenes.class_calls(sc)
    for W, B, L in [(8, 8, 1), (8, 8, 2), (8, 8, 3), (16, 8, 3), (16, 16, 3), (32, 8, 3), (32, 16, 3), (64, 8, 3), (64, 16, 3), (128, 16, 3), (8, 8, 4), (16,8,4), (32, 8, 4)]:
        tot = {}
        for call in calls:
            res = analyse(call['xyz_shift'], scenes.RADIUS, W, B, L)
            for k, v in res.items(): tot[k] = tot.get(k, 0) + v
        n = tot['n']
        print(f"W={W} B={B} levels={L}: tests now {tot['T_now']/n:.0f}/pt -> {tot['T_new']/n:.0f}/pt (+{tot['IN']/n:.0f} all-in), bucket classifications per QUERY GROUP*1/W {tot['ncls']/n:.1f}/pt", flush=True)
