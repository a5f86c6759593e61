"""Voxelize / devoxelize — host-side mirror of the MinkowskiEngine calls PBNet makes around its backbone
(SURVEY.md §8 a12-a14), on top of pb_voxelize / pb_voxel_rows / pb_devoxelize (include/pbnet_b200.h).

    sparse_quantize(coordinates, features, quantization_size=..., return_index=True, return_inverse=True)
        <- ME.utils.sparse_quantize            datasets/scannetv2/dataset_preprocess.py:269-274,348-353
    batched_coordinates / sparse_collate       dataset_preprocess.py:296,375 ; network/PBNet.py:237,264
    voxelize(features, coordinates, mode)      <- ME.SparseTensor(features, coordinates).{C,F,inverse_mapping}
                                               network/PBNet.py:236-247,261-271
    devoxelize(vfeat, inverse)                 <- X_v[v2p] (+ autograd scatter-add)   network/PBNet.py:130-134,250

MinkowskiEngine is not installed here and its voxel order is implementation-defined; this module returns
voxels in lexicographic (batch, x, y, z) order with the smallest point index as representative.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from ._lib import PBError, lib
from .cluster import PB_MEM_DEVICE, PB_MEM_HOST, default_context, stream_handle


def _bind():
    L = lib()
    if getattr(L, "_voxel_bound", False):
        return L
    vp = ctypes.c_void_p
    L.pb_voxelize.argtypes = [vp, vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, ctypes.c_int64, ctypes.c_double,
                              vp, vp, vp, vp, vp, ctypes.c_int64, ctypes.POINTER(ctypes.c_int64), ctypes.c_int, vp]
    L.pb_voxelize.restype = ctypes.c_int
    L.pb_voxel_rows.argtypes = [vp, vp, ctypes.c_int64, ctypes.c_int, vp, vp, ctypes.c_int64, ctypes.c_int, vp,
                                ctypes.c_int, vp]
    L.pb_voxel_rows.restype = ctypes.c_int
    L.pb_devoxelize.argtypes = [vp, vp, ctypes.c_int64, ctypes.c_int, vp, ctypes.c_int64, vp, ctypes.c_int, vp]
    L.pb_devoxelize.restype = ctypes.c_int
    L._voxel_bound = True
    return L


def _ctx_for(t: torch.Tensor):
    dev = t.device.index if t.is_cuda else (torch.cuda.current_device() if torch.cuda.is_available() else 0)
    return default_context(dev)


def _check(ctx, rc):
    if rc != 0:
        raise PBError(rc, ctx._lib.pb_last_error(ctx._h).decode())


class VoxelMap:
    """Result of a voxelization: vcoords[V,4] (batch,x,y,z), index[V], inverse[N], and the CSR
    (order[N], vox_start[V+1]) used by the segmented reductions."""

    def __init__(self, vcoords, index, inverse, order, vox_start):
        self.vcoords, self.index, self.inverse, self.order, self.vox_start = vcoords, index, inverse, order, vox_start

    @property
    def n_voxels(self):
        return int(self.vcoords.shape[0])


def voxel_map(coordinates, quantization_size=None, batch=None) -> VoxelMap:
    """coordinates: [N,3] (x,y,z) or [N,4] (batch,x,y,z) fp32/fp64 tensor (CPU or CUDA) or numpy array."""
    L = _bind()
    was_numpy = not isinstance(coordinates, torch.Tensor)
    c = torch.as_tensor(coordinates)
    if c.dtype not in (torch.float32, torch.float64):
        c = c.to(torch.float32)
    c = c.contiguous()
    n, stride = int(c.shape[0]), int(c.shape[1])
    ctx = _ctx_for(c)
    dev = c.device
    kind = PB_MEM_DEVICE if c.is_cuda else PB_MEM_HOST
    cap = max(n, 1)
    vcoords = torch.empty((cap, 4), dtype=torch.int32, device=dev)
    index = torch.empty(cap, dtype=torch.int64, device=dev)
    inverse = torch.empty(max(n, 1), dtype=torch.int64, device=dev)
    order = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
    vstart = torch.empty(cap + 1, dtype=torch.int32, device=dev)
    b = None
    if batch is not None:
        b = torch.as_tensor(batch).to(device=dev, dtype=torch.int32).contiguous()
    nv = ctypes.c_int64(0)
    sptr = stream_handle(torch.cuda.current_stream(dev)) if c.is_cuda else None
    rc = L.pb_voxelize(ctx._h, c.data_ptr(), int(c.dtype == torch.float64), stride, int(stride == 4 and b is None),
                       b.data_ptr() if b is not None else None, n,
                       float(quantization_size) if quantization_size else 0.0, vcoords.data_ptr(), index.data_ptr(),
                       inverse.data_ptr(), order.data_ptr(), vstart.data_ptr(), cap, ctypes.byref(nv), kind, sptr)
    _check(ctx, rc)
    V = int(nv.value)
    vm = VoxelMap(vcoords[:V], index[:V], inverse[:n], order[:n], vstart[:V + 1])
    vm._numpy = was_numpy
    return vm


def _out_like(out, shape, device):
    if out is None:
        return torch.empty(shape, dtype=torch.float32, device=device)
    if tuple(out.shape) != tuple(shape) or out.dtype != torch.float32 or out.device != device or not out.is_contiguous():
        raise TypeError(f"out: need a contiguous float32 tensor of shape {tuple(shape)} on {device}")
    return out


def voxel_rows(rows: torch.Tensor, vm: VoxelMap, mode: str = "pick", out: torch.Tensor | None = None) -> torch.Tensor:
    """Per-voxel reduction of point rows: 'pick' (representative), 'mean', 'sum' (ascending point order)."""
    L = _bind()
    rows = rows.to(torch.float32).contiguous()
    n, C = int(rows.shape[0]), int(rows.shape[1])
    ctx = _ctx_for(rows)
    V = vm.n_voxels
    out = _out_like(out, (V, C), rows.device)
    kind = PB_MEM_DEVICE if rows.is_cuda else PB_MEM_HOST
    sptr = stream_handle(torch.cuda.current_stream(rows.device)) if rows.is_cuda else None
    order, vstart = vm.order.to(rows.device), vm.vox_start.to(rows.device)
    rc = L.pb_voxel_rows(ctx._h, rows.data_ptr(), n, C, order.data_ptr(), vstart.data_ptr(), V,
                         {"pick": 0, "mean": 1, "sum": 2}[mode], out.data_ptr(), kind, sptr)
    _check(ctx, rc)
    return out


def devoxelize_raw(vfeat: torch.Tensor, inverse: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    L = _bind()
    vfeat = vfeat.to(torch.float32).contiguous()
    inverse = inverse.to(device=vfeat.device, dtype=torch.int64).contiguous()
    V, C = int(vfeat.shape[0]), int(vfeat.shape[1])
    n = int(inverse.shape[0])
    ctx = _ctx_for(vfeat)
    out = _out_like(out, (n, C), vfeat.device)
    kind = PB_MEM_DEVICE if vfeat.is_cuda else PB_MEM_HOST
    sptr = stream_handle(torch.cuda.current_stream(vfeat.device)) if vfeat.is_cuda else None
    rc = L.pb_devoxelize(ctx._h, vfeat.data_ptr(), V, C, inverse.data_ptr(), n, out.data_ptr(), kind, sptr)
    _check(ctx, rc)
    return out


class _Devoxelize(torch.autograd.Function):
    """out = vfeat[inverse]; backward = deterministic segmented sum over the points of each voxel."""

    @staticmethod
    def forward(ctx, vfeat, vm):
        ctx.vm = vm
        return devoxelize_raw(vfeat, vm.inverse)

    @staticmethod
    def backward(ctx, grad):
        return voxel_rows(grad.contiguous(), ctx.vm, "sum"), None


def devoxelize(vfeat: torch.Tensor, vm_or_inverse):
    """X_p = X_v[v2p] (network/PBNet.py:130-134,250).  With a VoxelMap the op is differentiable."""
    if isinstance(vm_or_inverse, VoxelMap):
        return _Devoxelize.apply(vfeat, vm_or_inverse)
    return devoxelize_raw(vfeat, vm_or_inverse)


# ---- MinkowskiEngine-shaped surface ------------------------------------------------------------------
def sparse_quantize(coordinates, features=None, labels=None, ignore_label=-100, return_index=False,
                    return_inverse=False, return_maps_only=False, quantization_size=None, device="cpu"):
    """Mirror of ME.utils.sparse_quantize for the keyword surface PBNet uses.  Returns
    (quantized_coords[V,3] int32, features[index], index, inverse) like ME does for
    return_index=True, return_inverse=True (dataset_preprocess.py:269-274)."""
    if labels is not None:
        raise NotImplementedError("labels / ignore_label voting is not used by PBNet")
    is_np = not isinstance(coordinates, torch.Tensor)
    vm = voxel_map(coordinates, quantization_size)
    index, inverse = vm.index, vm.inverse
    q = vm.vcoords[:, 1:]
    conv = (lambda t: t.cpu().numpy()) if is_np else (lambda t: t)
    if return_maps_only:
        out = [conv(index)]
        if return_inverse:
            out.append(conv(inverse))
        return tuple(out) if len(out) > 1 else out[0]
    out = [conv(q)]
    if features is not None:
        f = features[conv(index)] if is_np else features[index]
        out.append(f)
    if return_index:
        out.append(conv(index))
    if return_inverse:
        out.append(conv(inverse))
    return tuple(out) if len(out) > 1 else out[0]


def batched_coordinates(coords_list, dtype=torch.float32):
    """ME.utils.batched_coordinates: prepend the list position as batch column."""
    parts = []
    for b, c in enumerate(coords_list):
        c = torch.as_tensor(c)
        parts.append(torch.cat([torch.full((c.shape[0], 1), b, dtype=c.dtype, device=c.device), c], dim=1))
    return torch.cat(parts, 0).to(dtype)


def sparse_collate(coords_list, feats_list):
    """ME.utils.sparse_collate for (coords, feats) lists (dataset_preprocess.py:296,375)."""
    bc = batched_coordinates(coords_list, dtype=torch.int32)
    feats = torch.cat([torch.as_tensor(f) for f in feats_list], 0)
    return bc, feats


def voxelize(features: torch.Tensor, coordinates: torch.Tensor, mode: str = "pick"):
    """ME.SparseTensor(features, coordinates=batched float coords) as PBNet builds it
    (network/PBNet.py:236-247): returns (voxel_features[V,C], voxel_coords[V,4] int32, VoxelMap).
    ``VoxelMap.inverse`` is ME's ``inverse_mapping``; mode='mean' = UNWEIGHTED_AVERAGE."""
    vm = voxel_map(coordinates, None)
    return voxel_rows(features, vm, mode), vm.vcoords, vm
