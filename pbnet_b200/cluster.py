"""Host side of the grouping path: thin Python over the C ABI (include/pbnet_b200.h).

``Context.binary_cluster`` is the array-level call (torch CPU / CUDA tensors or numpy arrays);
``pbnet_b200.pbnet_ops.cluster`` and ``pbnet_b200.shim.PB_lib.binary_cluster`` mirror the reference's
operator surface (lib/PB_lib/torch_io/pbnet_ops.py:12-82, lib/PB_lib/src/PB_lib_api.cpp:7) on top of it.
"""
from __future__ import annotations

import ctypes
import os
import threading

import numpy as np
import torch

from ._lib import PBError, lib

PB_MEM_HOST, PB_MEM_DEVICE = 0, 1


CUDA_STREAM_LEGACY = 1  # cudaStreamLegacy: the explicit handle of the default stream (a NULL stream argument
#                         means "the context's own stream" in the C ABI)


def stream_handle(stream) -> int:
    """cudaStream_t value for the C ABI from a torch stream / raw handle; torch's default stream (handle 0)
    maps to cudaStreamLegacy so the call is ordered with the tensors' producers and consumers."""
    h = stream.cuda_stream if hasattr(stream, "cuda_stream") else int(stream)
    return h if h else CUDA_STREAM_LEGACY


def _addr(a):
    if a is None:
        return None
    if isinstance(a, torch.Tensor):
        return a.data_ptr()
    return a.ctypes.data


def _as_host_i32(a):
    if isinstance(a, torch.Tensor):
        a = a.detach().cpu().numpy()
    return np.ascontiguousarray(a, dtype=np.int32)


def _as_host_f32(a):
    if isinstance(a, torch.Tensor):
        a = a.detach().cpu().numpy()
    return np.ascontiguousarray(a, dtype=np.float32)


_T_DTYPES = (torch.float32, torch.int32)
_N_DTYPES = (np.dtype("float32"), np.dtype("int32"))


def _check_arrays(items, is_torch, on_device, device, n, exact):
    """items: (name, array, 0 = float32 | 1 = int32).  Contiguous 1-D arrays of the right dtype, all in one place."""
    for name, a, k in items:
        if isinstance(a, torch.Tensor) != is_torch:
            raise TypeError(f"{name}: mixing torch and numpy arrays")
        if is_torch:
            ok = a.dtype == _T_DTYPES[k] and a.dim() == 1 and a.is_contiguous() and a.is_cuda == on_device and (
                not on_device or a.device.index == device)
        else:
            ok = a.dtype == _N_DTYPES[k] and a.ndim == 1 and a.flags["C_CONTIGUOUS"] and (exact or a.flags["WRITEABLE"])
        if not ok:
            raise TypeError(f"{name}: need a contiguous 1-D {'int32' if k else 'float32'} array next to x (same device)")
        if (a.shape[0] != n) if exact else (a.shape[0] < n):
            raise ValueError(f"{name}: length {a.shape[0]} does not fit {n}")


class Context:
    """Owns a pb_ctx (stream + device workspace).  Not thread-safe: one Context per thread."""

    def __init__(self, device: int = 0, profiling: bool = False):
        self._lib = lib()
        h = ctypes.c_void_p()
        rc = self._lib.pb_create(int(device), ctypes.byref(h))
        if rc != 0:
            raise PBError(rc, f"pb_create(device={device}) failed — a CUDA device is required (no CPU fallback)")
        self._h = h
        self.device = int(device)
        if profiling:
            self.set_profiling(True)
        if os.environ.get("PB_CHUNK_POINTS"):  # experiments: chunk size of the two-stream pipelining
            self.set_chunk_points(int(os.environ["PB_CHUNK_POINTS"]))

    def close(self):
        if getattr(self, "_h", None):
            self._lib.pb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_profiling(self, on: bool):
        self._lib.pb_set_profiling(self._h, int(bool(on)))

    def set_chunk_points(self, points: int):
        """Chunk size (points) of the two-stream pipelining of large batched calls; 0 = automatic."""
        self._lib.pb_set_chunk_points(self._h, int(points))

    def set_small_calls(self, mode: int):
        """Small-call kernel (one cooperative launch for a per-class call): 0 = never, 1 / -1 = whenever eligible (default)."""
        self._lib.pb_set_small_calls(self._h, int(mode))

    def selftest_division(self, n_samples: int = 1 << 28, seed: int = 0) -> int:
        """Mismatches between k_centres' reciprocal-based quotient and div.rn.f32 (expected 0)."""
        bad = ctypes.c_int64(-1)
        rc = self._lib.pb_selftest_division(self._h, int(n_samples), int(seed), ctypes.byref(bad))
        if rc != 0:
            raise PBError(rc, self._lib.pb_last_error(self._h).decode())
        return int(bad.value)

    @property
    def last_launch_count(self) -> int:
        return int(self._lib.pb_last_launch_count(self._h))

    def stage_ms(self) -> dict:
        n = self._lib.pb_stage_count()
        return {self._lib.pb_stage_name(i).decode(): float(self._lib.pb_stage_ms(self._h, i)) for i in range(n)}

    def counters(self) -> dict:
        names = ["pair_tests", "sum_deg", "n_hp", "lp_queries", "cells", "raw_clusters", "chunks", "mixed_mode", "coarse_cells",
                 "small_path", "intra_tests", "centre_halves", "centre_halves_replayed",
                 "centre_replay_cycles", "centre_gather_cycles"]
        return {k: int(self._lib.pb_counter(self._h, i)) for i, k in enumerate(names)}

    # ------------------------------------------------------------------------------------------------
    def binary_cluster(self, x, y, z, xo, yo, zo, sem, seg_counts, radius18, min_pts18, para_f=0.05,
                       assign_lp=True, call_seg_counts=None, cluster_id=None, cluster_num=None, degree=None,
                       center=None, clt_sem=None, stream=None):
        """Runs the grouping path on SoA fp32 coordinates / int32 classes.

        All of x..sem must live in the same place: host (numpy or torch CPU — pinned torch tensors make
        the copies asynchronous) or CUDA device ``self.device``.  ``seg_counts`` / ``call_seg_counts`` /
        ``radius18`` / ``min_pts18`` are host metadata.  Outputs are allocated next to the inputs unless
        passed in.  Returns dict(cluster_id, cluster_num, degree, center, clt_sem, n_clusters,
        call_clusters)."""
        is_torch = isinstance(x, torch.Tensor)
        on_device = is_torch and x.is_cuda
        n = int(x.shape[0])
        seg = _as_host_i32(seg_counts)
        S = int(seg.shape[0])
        r18 = _as_host_f32(radius18)
        m18 = _as_host_i32(min_pts18)
        if r18.shape != (18,) or m18.shape != (18,):
            raise ValueError("radius / min_pts tables must have 18 entries")
        _check_arrays((("x", x, 0), ("y", y, 0), ("z", z, 0), ("xo", xo, 0), ("yo", yo, 0), ("zo", zo, 0), ("sem", sem, 1)),
                      is_torch, on_device, self.device, n, exact=True)

        def new(shape, dtype_t, dtype_n, like_pinned=False):
            if is_torch:
                if on_device:
                    return torch.empty(shape, dtype=dtype_t, device=x.device)
                return torch.empty(shape, dtype=dtype_t, pin_memory=like_pinned)
            return np.empty(shape, dtype=dtype_n)

        pinned = is_torch and not on_device and x.is_pinned()
        given = []
        if cluster_id is None:
            cluster_id = new(n, torch.int32, np.int32, pinned)
        else:
            given.append(("cluster_id", cluster_id, 1, n))
        if cluster_num is None:
            cluster_num = new(S, torch.int32, np.int32, pinned)
        else:
            given.append(("cluster_num", cluster_num, 1, S))
        if degree is None:
            degree = new(n, torch.int32, np.int32, pinned)
        else:
            given.append(("degree", degree, 1, n))
        cap = max(n, 1)
        if center is None:
            center = new(3 * cap, torch.float32, np.float32, pinned)
        else:
            given.append(("center", center, 0, 0))
        if clt_sem is None:
            clt_sem = new(cap, torch.int32, np.int32, pinned)
        else:
            given.append(("clt_sem", clt_sem, 1, 0))
        # caller-supplied outputs go to C as raw pointers: check them like the inputs
        for name, a, kind_, need in given:
            _check_arrays(((name, a, kind_),), is_torch, on_device, self.device, need, exact=False)
        nclt = ctypes.c_int64(0)
        kind = PB_MEM_DEVICE if on_device else PB_MEM_HOST
        sptr = stream_handle(stream) if stream is not None else (stream_handle(torch.cuda.current_stream(x.device))
                                                                 if on_device else None)
        if call_seg_counts is None:
            call_clusters = None
            rc = self._lib.pb_binary_cluster(
                self._h, _addr(x), _addr(y), _addr(z), _addr(xo), _addr(yo), _addr(zo), _addr(sem), _addr(seg), S, n,
                _addr(r18), _addr(m18), float(para_f), int(bool(assign_lp)), _addr(cluster_id), _addr(cluster_num),
                _addr(degree), _addr(center), int(center.shape[0]), _addr(clt_sem), int(clt_sem.shape[0]),
                ctypes.byref(nclt), kind, sptr)
        else:
            calls = _as_host_i32(call_seg_counts)
            call_clusters = np.zeros(calls.shape[0], dtype=np.int64)
            rc = self._lib.pb_binary_cluster_batched(
                self._h, _addr(x), _addr(y), _addr(z), _addr(xo), _addr(yo), _addr(zo), _addr(sem), _addr(seg), S,
                _addr(calls), int(calls.shape[0]), n, _addr(r18), _addr(m18), float(para_f), int(bool(assign_lp)),
                _addr(cluster_id), _addr(cluster_num), _addr(degree), _addr(center), int(center.shape[0]),
                _addr(clt_sem), int(clt_sem.shape[0]), ctypes.byref(nclt), _addr(call_clusters), kind, sptr)
        if rc != 0:
            raise PBError(rc, self._lib.pb_last_error(self._h).decode())
        k = int(nclt.value)
        return dict(cluster_id=cluster_id, cluster_num=cluster_num, degree=degree, center=center[:3 * k],
                    clt_sem=clt_sem[:k], n_clusters=k, call_clusters=call_clusters)


_tls = threading.local()


def default_context(device: int = 0) -> Context:
    """Per-thread, per-device cached context (workspace is reused across calls)."""
    d = getattr(_tls, "ctx", None)
    if d is None:
        d = _tls.ctx = {}
    if device not in d:
        d[device] = Context(device)
    return d[device]
