"""Offline estimate (CPU, numpy): how many exact pair tests of k_degree survive if candidate cells are first
classified against tight per-cell bounding boxes (all-in / all-out / mixed)?  Not part of the product."""
import sys, numpy as np
sys.path.insert(0, '/root/repo')
from pbnet_b200 import scenes

def analyse(xyz, r, sub=1):
    h = np.float32(r / 2 * (1 + 2.0 ** -7)) / sub
    mn = xyz.min(0)
    c = np.floor((xyz - mn) / h).astype(np.int64)
    key = (c[:, 2] << 40) | (c[:, 1] << 20) | c[:, 0]
    order = np.argsort(key, kind='stable')
    key = key[order]; p = xyz[order].astype(np.float64); c = c[order]
    uk, start, cnt = np.unique(key, return_index=True, return_counts=True)
    F = len(uk)
    lo = np.minimum.reduceat(p, start, axis=0); hi = np.maximum.reduceat(p, start, axis=0)
    cc = c[start]
    cell_of = np.repeat(np.arange(F), cnt)
    R = 2 * sub + 1  # offsets -R..R cover min distance <= r
    res = dict(F=F, n=len(p), T_sten=0, T_B=0, T_A=0, cellpairs=0, ptcell=0, in_B=0, in_A=0, hits=0, T_coarse=0)
    r2 = r * r
    rng = range(-R, R + 1)
    for dz in rng:
        for dy in rng:
            for dx in rng:
                nk = ((cc[:, 2] + dz) << 40) | ((cc[:, 1] + dy) << 20) | (cc[:, 0] + dx)
                ok = (cc[:, 0] + dx >= 0) & (cc[:, 1] + dy >= 0) & (cc[:, 2] + dz >= 0)
                j = np.searchsorted(uk, nk)
                j[j >= F] = F - 1
                m = ok & (uk[j] == nk)
                q = np.nonzero(m)[0]; cnd = j[m]
                if len(q) == 0: continue
                # cell-level bounds
                gap = np.maximum(0, np.maximum(lo[cnd] - hi[q], lo[q] - hi[cnd]))
                far = np.maximum(hi[cnd] - lo[q], hi[q] - lo[cnd])
                dmin = (gap ** 2).sum(1); dmax = (far ** 2).sum(1)
                w = cnt[q] * cnt[cnd]
                res['T_sten'] += int(w.sum())
                live = dmin <= r2
                res['cellpairs'] += int(live.sum())
                allin = dmax <= r2
                res['in_B'] += int(w[allin].sum())
                mixedB = live & ~allin
                res['T_B'] += int(w[mixedB].sum())
                # per-point vs candidate cell box, for cell pairs that are mixed at cell level
                qm = q[mixedB]; cm = cnd[mixedB]
                if len(qm) == 0: continue
                # expand points of query cells
                reps = cnt[qm]
                pi = np.concatenate([np.arange(s, s + k) for s, k in zip(start[qm], reps)]) if len(qm) < 200000 else None
                ci = np.repeat(cm, reps)
                pp = p[pi]
                gap = np.maximum(0, np.maximum(lo[ci] - pp, pp - hi[ci]))
                far = np.maximum(hi[ci] - pp, pp - lo[ci])
                dmin = (gap ** 2).sum(1); dmax = (far ** 2).sum(1)
                res['ptcell'] += len(pi)
                pin = dmax <= r2; pout = dmin > r2
                res['in_A'] += int(cnt[ci][pin].sum())
                res['T_A'] += int(cnt[ci][~pin & ~pout].sum())
    return res

if __name__ == '__main__':
    seed = int(sys.argv[1]) if len(sys.argv) > 1 else 22
    npts = int(sys.argv[2]) if len(sys.argv) > 2 else 150000
    sc = scenes.make_scene(seed, npts)
    tot = {}
    for sub in (1, 2):
        tot = {}
        for call in scenes.class_calls(sc):
            res = analyse(call['xyz_shift'], scenes.RADIUS, sub)
            for k, v in res.items(): tot[k] = tot.get(k, 0) + v
        n = tot['n']
        print(f"sub={sub} pts={n} cells={tot['F']} per point: stencil tests {tot['T_sten']/n:.0f}  cell-pair tests/pt {tot['cellpairs']/n:.1f}  "
              f"B: in {tot['in_B']/n:.0f} mixed tests {tot['T_B']/n:.0f}   A: pt-cell tests {tot['ptcell']/n:.1f} in {tot['in_A']/n:.0f} mixed tests {tot['T_A']/n:.0f}")
