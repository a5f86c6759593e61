"""Deterministic synthetic ScanNet-v2-shaped scenes — the workload of bench.py and the
parity tests.  No ScanNet data exists in this environment, so scenes are generated:

* room 8 x 6 x 3 m, N points; 38 % on the floor / four walls (classes 0/1, dropped before grouping as in
  network/PBNet.py:151-152), the rest on 20-30 box / ellipsoid surfaces sampled at ~2 cm spacing;
* object classes drawn in proportion to ``count_mean`` (network/PBNet.py:33-34), sizes
  ``count_mean[c] * U(0.3, 2.5)`` rescaled to the foreground budget;
* predicted offset = (centroid - xyz) * U(0.85, 1.0) + N(0, 0.015^2) for 85 % of the points (they
  collapse into high-density blobs) and N(0, 0.03^2) for the other 15 % (low-density points);
* predicted class = true class with 3 % uniform label noise;
* ``xyz_shift = fp32(xyz_orig) + fp32(offset)`` exactly as ``ins_orig.cpu() + ins_offset.cpu()``
  (network/PBNet.py:165).

Generator: ``numpy.random.Generator(PCG64(seed))``; scene ``s`` of a set uses seed ``22 + s``
(``--manual_seed 22``, config/config.py:15).
"""
from __future__ import annotations

import numpy as np

# network/PBNet.py:33-34 (classes 0,1 = wall/floor have no entry)
COUNT_MEAN = np.array([-1., -1., 3917., 12056., 2303., 8331., 3948., 3166., 5629., 11719., 1003.,
                       3317., 4912., 10221., 3889., 4136., 2120., 945., 3967., 2589.], dtype=np.float32)
BASE_SEED = 22
ROOM = np.array([8.0, 6.0, 3.0])
RADIUS = 0.04      # config/config.py:45
MIN_PTS = 31       # config/config.py:44
VOXEL_SIZE = 0.02  # config/config.py:26
N_VAL_SCENES = 312  # datasets/scannetv2/scannetv2_val.txt


def _sample_box_surface(rng, n, half):
    """n points uniform on the surface of an axis-aligned box with half extents `half`."""
    a = np.array([half[1] * half[2], half[0] * half[2], half[0] * half[1]])  # face areas per axis
    axis = rng.choice(3, size=n, p=a / a.sum())
    p = rng.uniform(-1.0, 1.0, size=(n, 3)) * half
    sign = rng.choice([-1.0, 1.0], size=n)
    p[np.arange(n), axis] = sign * half[axis]
    return p


def _sample_ellipsoid_surface(rng, n, half):
    v = rng.normal(size=(n, 3))
    v /= np.linalg.norm(v, axis=1, keepdims=True) + 1e-12
    return v * half


def make_scene(seed: int, n_points: int = 150_000, hp_frac: float = 0.85):
    """Returns dict(xyz_orig f32[N,3], offset f32[N,3], sem i64[N], n_objects).  ``hp_frac`` = fraction of
    object points whose offset collapses them onto the object centroid (0.85 = SURVEY.md §8d C0/C1)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    n_bg = int(round(0.38 * n_points))
    n_fg = n_points - n_bg
    # ---- background: floor + 4 walls --------------------------------------------------------------
    area = np.array([ROOM[0] * ROOM[1], ROOM[0] * ROOM[2], ROOM[0] * ROOM[2], ROOM[1] * ROOM[2],
                     ROOM[1] * ROOM[2]])
    which = rng.choice(5, size=n_bg, p=area / area.sum())
    bg = rng.uniform(0.0, 1.0, size=(n_bg, 3)) * ROOM
    bg[which == 0, 2] = 0.0
    bg[which == 1, 1] = 0.0
    bg[which == 2, 1] = ROOM[1]
    bg[which == 3, 0] = 0.0
    bg[which == 4, 0] = ROOM[0]
    bg_sem = np.where(which == 0, 1, 0).astype(np.int64)  # 0 wall, 1 floor
    # ---- objects ------------------------------------------------------------------------------------
    n_obj = int(rng.integers(20, 31))
    p_cls = COUNT_MEAN[2:] / COUNT_MEAN[2:].sum()
    cls = rng.choice(np.arange(2, 20), size=n_obj, p=p_cls)
    raw = COUNT_MEAN[cls] * rng.uniform(0.3, 2.5, size=n_obj)
    sizes = np.maximum((raw * (n_fg / raw.sum())).astype(np.int64), 16)
    sizes[-1] = max(16, n_fg - int(sizes[:-1].sum()))
    if sizes.sum() != n_fg:  # the clamp above can overshoot on tiny scenes
        sizes = np.maximum((sizes * (n_fg / sizes.sum())).astype(np.int64), 1)
        sizes[0] += n_fg - int(sizes.sum())
    xyz_list, off_list, sem_list = [], [], []
    for k in range(n_obj):
        n = int(sizes[k])
        area_k = n * (0.02 ** 2)  # ~2 cm sample spacing
        aspect = rng.uniform(0.5, 1.5, size=3)
        is_box = rng.random() < 0.6
        # surface area of a box with half extents s*aspect: 8 s^2 (a0a1+a0a2+a1a2)
        cross = aspect[0] * aspect[1] + aspect[0] * aspect[2] + aspect[1] * aspect[2]
        s = np.sqrt(area_k / (8.0 * cross)) if is_box else np.sqrt(area_k / (4.19 * cross))
        half = s * aspect
        pts = _sample_box_surface(rng, n, half) if is_box else _sample_ellipsoid_surface(rng, n, half)
        centre = np.array([rng.uniform(half[0], ROOM[0] - half[0]) if 2 * half[0] < ROOM[0] else ROOM[0] / 2,
                           rng.uniform(half[1], ROOM[1] - half[1]) if 2 * half[1] < ROOM[1] else ROOM[1] / 2,
                           half[2] + rng.uniform(0.0, 0.5)])
        pts = pts + centre
        centroid = pts.mean(axis=0)
        hp = rng.random(n) < hp_frac
        off = np.where(hp[:, None],
                       (centroid - pts) * rng.uniform(0.85, 1.0, size=(n, 1)) + rng.normal(0, 0.015, size=(n, 3)),
                       rng.normal(0, 0.03, size=(n, 3)))
        xyz_list.append(pts)
        off_list.append(off)
        sem_list.append(np.full(n, cls[k], dtype=np.int64))
    xyz = np.concatenate([bg] + xyz_list, axis=0)
    off = np.concatenate([rng.normal(0, 0.03, size=(n_bg, 3))] + off_list, axis=0)
    sem = np.concatenate([bg_sem] + sem_list, axis=0)
    # 3 % uniform label noise on the predicted class
    noisy = rng.random(n_points) < 0.03
    sem = np.where(noisy, rng.integers(0, 20, size=n_points), sem).astype(np.int64)
    # scan order is arbitrary in ScanNet: shuffle so that class subsets are not object-sorted
    perm = rng.permutation(n_points)
    return dict(xyz_orig=xyz[perm].astype(np.float32), offset=off[perm].astype(np.float32),
                sem=sem[perm], n_objects=n_obj)


def scene_sizes(n_scenes: int = N_VAL_SCENES, seed: int = BASE_SEED):
    """Point counts of the synthetic val set: clip(lognormal(ln 140k, 0.45), 50k, 250k)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    n = np.exp(rng.normal(np.log(140_000.0), 0.45, size=n_scenes))
    return np.clip(n, 50_000, 250_000).astype(np.int64)


def rotate_copies(xyz: np.ndarray, copies: int = 3):
    """The eval loader's rotated copies: theta = 0.35*pi + i*2*pi/3 about z
    (datasets/scannetv2/dataset_preprocess.py:91,324)."""
    out = []
    for i in range(copies):
        t = 0.35 * np.pi + i * 2.0 * np.pi / 3.0
        m = np.array([[np.cos(t), np.sin(t), 0.0], [-np.sin(t), np.cos(t), 0.0], [0.0, 0.0, 1.0]])
        out.append((xyz.astype(np.float64) @ m).astype(np.float32))
    return out


def class_calls(scene: dict, copies: int = 1):
    """The per-class call list of PBNet.forward (network/PBNet.py:151-179): one entry per foreground
    class that passes the ``count < count_mean*0.05`` skip.  Each entry is
    dict(sem_id, xyz_shift f32[I,3], xyz_orig f32[I,3], sem i64[I], seg_counts i32[copies], index i64[I]).
    With copies > 1 the scene is replicated as rotated copies (batch index = copy), as in evaluation."""
    if copies == 1:
        xyz_all = [scene["xyz_orig"]]
        off_all = [scene["offset"]]
    else:
        xyz_all = rotate_copies(scene["xyz_orig"], copies)
        off_all = rotate_copies(scene["offset"], copies)
    sem = scene["sem"]
    calls = []
    for sem_id in range(2, 20):
        ind = np.nonzero(sem == sem_id)[0]
        if ind.shape[0] * copies < COUNT_MEAN[sem_id] * np.float32(0.05):
            continue
        orig = np.concatenate([x[ind] for x in xyz_all], axis=0)
        offs = np.concatenate([o[ind] for o in off_all], axis=0)
        calls.append(dict(sem_id=sem_id, xyz_shift=(orig + offs).astype(np.float32), xyz_orig=orig,
                          sem=np.full(orig.shape[0], sem_id, dtype=np.int64),
                          seg_counts=np.full(copies, ind.shape[0], dtype=np.int32), index=ind))
    return calls


def make_dense_case(seed: int, n_points: int, hp_fraction: float, radius: float = RADIUS, blob_points: int = 2001,
                    sem_id: int = 9):
    """SURVEY.md §8d config C4 (dense-neighbour pathological case): one single-class segment;
    ``hp_fraction`` of the points sit in tight Gaussian blobs (sigma = r/4, ``blob_points`` each, degree
    ~ blob_points - 1), the rest is a uniform background whose expected degree stays below min_pts.
    Returns (xyz_shift, xyz_orig, sem)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    n_hp = int(round(hp_fraction * n_points)) // blob_points * blob_points
    n_bg = n_points - n_hp
    # background density: expected neighbours in an r-ball = 8 -> volume per point = (4/3)pi r^3 / 8
    side = (max(n_bg, 1) * (4.0 / 3.0) * np.pi * radius ** 3 / 8.0) ** (1.0 / 3.0)
    side = max(side, 1.0)
    bg = rng.uniform(0.0, side, size=(n_bg, 3))
    k = n_hp // blob_points
    centres = rng.uniform(0.0, side, size=(k, 3))
    blobs = (centres[:, None, :] + rng.normal(0.0, radius / 4.0, size=(k, blob_points, 3))).reshape(-1, 3)
    xs = np.concatenate([bg, blobs]).astype(np.float32)
    perm = rng.permutation(n_points)
    xs = xs[perm]
    xo = (xs + rng.normal(0.0, 0.05, size=xs.shape)).astype(np.float32)  # original coords: smeared copies
    return xs, xo, np.full(n_points, sem_id, dtype=np.int32)
