"""numpy restatement of the voxelize / devoxelize contract — TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED: the arithmetic lives in MinkowskiEngine (module ``MinkowskiEngine``, installed from
``git+https://github.com/NVIDIA/MinkowskiEngine`` master with no pinned version, README.md:21), which is
neither vendored under /root/reference nor installed here, and the reference has no tests at this
boundary.  This file restates ME's documented contract (ME >= 0.5, ``MinkowskiEngine/utils/quantization.py``):
``discrete = floor(coordinates / quantization_size)`` as int32, unique rows, ``coords[index][inverse] ==
coords``, features picked at ``index`` (default) or averaged (UNWEIGHTED_AVERAGE).  Voxel ORDER is
implementation-defined in ME; tests compare modulo a permutation (ours is numpy.unique's lexicographic order).
Call sites anchored: datasets/scannetv2/dataset_preprocess.py:269-274,348-353; network/PBNet.py:130-134,236-250.
"""
import numpy as np


def quantize(coords, quantization_size=None, batch=None):
    c = np.asarray(coords)
    if c.dtype not in (np.float32, np.float64):
        c = c.astype(np.float32)
    if c.shape[1] == 4 and batch is None:
        batch, c = c[:, 0].astype(np.int32), c[:, 1:]
    if quantization_size:
        c = c / c.dtype.type(quantization_size)
    q = np.floor(c).astype(np.int32)
    b = np.zeros(len(q), np.int32) if batch is None else np.asarray(batch, np.int32)
    return np.concatenate([b[:, None], q], axis=1)


def sparse_quantize(coords, quantization_size=None, batch=None):
    """Returns (vcoords[V,4] lexicographic, index[V] first occurrence, inverse[N])."""
    q = quantize(coords, quantization_size, batch)
    vc, index, inverse = np.unique(q, axis=0, return_index=True, return_inverse=True)
    return vc.astype(np.int32), index.astype(np.int64), inverse.reshape(-1).astype(np.int64)


def voxel_rows(rows, inverse, n_voxels, mode):
    rows = np.asarray(rows, np.float64)
    out = np.zeros((n_voxels, rows.shape[1]), np.float64)
    if mode == "pick":
        first = np.full(n_voxels, -1, np.int64)
        for p in range(len(inverse) - 1, -1, -1):
            first[inverse[p]] = p
        return rows[first]
    np.add.at(out, inverse, rows)
    if mode == "mean":
        out /= np.bincount(inverse, minlength=n_voxels)[:, None]
    return out


def devoxelize(vfeat, inverse):
    return np.asarray(vfeat)[inverse]
