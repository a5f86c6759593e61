"""ctypes front-end of the CPU oracle (oracle/pb_oracle.c) — TEST INFRASTRUCTURE ONLY.

Only tests/, ``__graft_entry__.smoke()`` and bench.py's ``cpu_baseline`` / ``--impl reference`` legs may
import this module.  The product package (pbnet_b200/) never does.

``oracle_cluster`` mirrors ``pbnet_ops.cluster`` of the reference
(lib/PB_lib/torch_io/pbnet_ops.py:14-75) on numpy arrays.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libpb_oracle.so")
_lib = None


class OracleStats(ctypes.Structure):
    _fields_ = [
        ("pair_tests", ctypes.c_int64),
        ("sum_deg", ctypes.c_int64),
        ("n_hp", ctypes.c_int64),
        ("n_noise", ctypes.c_int64),
        ("nn_tests", ctypes.c_int64),
    ]

    def as_dict(self):
        return {k: int(getattr(self, k)) for k, _ in self._fields_}


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (seconds)."""
    src = os.path.join(_HERE, "pb_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "--silent"])
    return _SO


def _load():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        fp = ctypes.POINTER(ctypes.c_float)
        ip = ctypes.POINTER(ctypes.c_int32)
        sig = [fp, fp, fp, fp, fp, fp, ip, ip, ctypes.c_int, fp, ip, ctypes.c_float, ctypes.c_int,
               ip, ip, ip, fp, ip, ip, ctypes.POINTER(OracleStats)]
        for name in ("pb_oracle_literal", "pb_oracle_grid"):
            fn = getattr(_lib, name)
            fn.argtypes = sig
            fn.restype = ctypes.c_int
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _fp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def _ip(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))


def oracle_binary_cluster(xyz_shift, xyz_orig, sem, seg_counts, radius18, min_pts18, para_f=0.05,
                          nv_flag=True, mode="grid", return_stats=False):
    """Raw oracle call.  Returns dict(cluster_id, cluster_num, den_queue (raw degree), center, clt_sem)."""
    lib = _load()
    xyz_shift = np.asarray(xyz_shift, dtype=np.float32).reshape(-1, 3)
    xyz_orig = np.asarray(xyz_orig, dtype=np.float32).reshape(-1, 3)
    n = xyz_shift.shape[0]
    x, y, z = (_f32(xyz_shift[:, i]) for i in range(3))
    xo, yo, zo = (_f32(xyz_orig[:, i]) for i in range(3))
    sem = _i32(sem)
    seg = _i32(seg_counts)
    assert int(seg.sum()) == n, "sum(seg_counts) != number of points"
    radius18 = _f32(radius18)
    min_pts18 = _i32(min_pts18)
    assert radius18.shape == (18,) and min_pts18.shape == (18,)
    cluster_id = np.full(n, -1, dtype=np.int32)
    cluster_num = np.zeros(len(seg), dtype=np.int32)
    den = np.zeros(n, dtype=np.int32)
    center = np.zeros(3 * max(n, 1), dtype=np.float32)
    clt_sem = np.zeros(max(n, 1), dtype=np.int32)
    nclt = np.zeros(1, dtype=np.int32)
    st = OracleStats()
    fn = lib.pb_oracle_grid if mode == "grid" else lib.pb_oracle_literal
    rc = fn(_fp(x), _fp(y), _fp(z), _fp(xo), _fp(yo), _fp(zo), _ip(sem), _ip(seg), len(seg),
            _fp(radius18), _ip(min_pts18), ctypes.c_float(para_f), int(bool(nv_flag)),
            _ip(cluster_id), _ip(cluster_num), _ip(den), _fp(center), _ip(clt_sem), _ip(nclt),
            ctypes.byref(st))
    if rc != 0:
        raise ValueError(f"pb_oracle: invalid arguments (rc={rc})")
    k = int(nclt[0])
    out = dict(cluster_id=cluster_id, cluster_num=cluster_num, den_queue=den,
               center=center[:3 * k].copy(), clt_sem=clt_sem[:k].copy())
    if return_stats:
        out["stats"] = st.as_dict()
    return out


def oracle_cluster(ins_offseted, ins_orig, sem, ins_bp, radius, min_pts, batch_size=None, mode="grid"):
    """numpy mirror of pbnet_ops.cluster: returns (cluster_id, cluster_num, den_queue + 1, center)."""
    r18 = (np.ones(18, dtype=np.float32) * np.float32(radius)).astype(np.float32)
    m18 = (np.ones(18) * min_pts).astype(np.int32)
    o = oracle_binary_cluster(ins_offseted, ins_orig, sem, ins_bp, r18, m18, 0.05, True, mode)
    return o["cluster_id"], o["cluster_num"], o["den_queue"] + 1, o["center"]


def oracle_normals(xyz, face, num_face=None):
    """Vertex normals as lib/PB_lib/src/normal/cal_normal.cu computes them (pb_oracle_normals in pb_oracle.c).
    xyz f32[V,3], face i32[F,3]; ``num_face`` = how many leading faces take part (the reference wrapper passes V,
    lib/PB_lib/torch_io/pbnet_ops.py:163)."""
    L = _load()
    L.pb_oracle_normals.argtypes = [ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int32), ctypes.c_int, ctypes.c_int,
                                    ctypes.POINTER(ctypes.c_float)]
    L.pb_oracle_normals.restype = ctypes.c_int
    xyz = _f32(xyz).reshape(-1, 3)
    face = _i32(face).reshape(-1, 3)
    nf = face.shape[0] if num_face is None else int(num_face)
    assert 0 <= nf <= face.shape[0]
    out = np.empty_like(xyz)
    rc = L.pb_oracle_normals(_fp(xyz), _ip(face), xyz.shape[0], nf, _fp(out))
    if rc != 0:
        raise ValueError(f"pb_oracle_normals failed ({rc}): vertex index out of range")
    return out
