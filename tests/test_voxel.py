"""Voxelize / devoxelize: numpy oracle self-checks on CPU; CUDA path vs oracle under -m gpu.
Integer outputs (voxel coords, index, inverse) bit-exact; means within 1e-5 relative (north star)."""
import numpy as np
import pytest

from oracle import voxel_oracle as vo


def cloud(seed, n, dtype):
    rng = np.random.Generator(np.random.PCG64(seed))
    xyz = rng.uniform(-1.5, 6.5, size=(n, 3))
    xyz[: n // 5] = np.round(xyz[: n // 5] / 0.02) * 0.02  # points exactly on voxel boundaries
    return xyz.astype(dtype)


def test_oracle_contract_cpu():
    for dtype in (np.float32, np.float64):
        xyz = cloud(1, 5000, dtype)
        vc, index, inverse = vo.sparse_quantize(xyz, 0.02)
        q = vo.quantize(xyz, 0.02)
        assert np.array_equal(vc[inverse], q)                  # coords[index][inverse] == coords
        assert np.array_equal(q[index], vc)
        assert len(np.unique(vc, axis=0)) == len(vc)
        for v in range(0, len(vc), 97):                        # representative = first occurrence
            assert index[v] == np.nonzero(inverse == v)[0].min()
        feats = np.random.default_rng(0).normal(size=(5000, 6)).astype(np.float32)
        mean = vo.voxel_rows(feats, inverse, len(vc), "mean")
        assert np.allclose(mean[inverse[0]], feats[inverse == inverse[0]].astype(np.float64).mean(0))
        assert np.array_equal(vo.voxel_rows(feats, inverse, len(vc), "pick"), feats[index].astype(np.float64))


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.float32, np.float64], ids=["f32", "f64"])
@pytest.mark.parametrize("where", ["host", "device"])
def test_sparse_quantize_matches_contract(dtype, where):
    import torch
    from pbnet_b200 import voxel
    xyz = cloud(2, 200_000, dtype)
    vc, index, inverse = vo.sparse_quantize(xyz, 0.02)
    t = torch.from_numpy(xyz)
    if where == "device":
        t = t.cuda()
    vm = voxel.voxel_map(t, 0.02)
    assert np.array_equal(vm.vcoords.cpu().numpy(), vc)
    assert np.array_equal(vm.index.cpu().numpy(), index)
    assert np.array_equal(vm.inverse.cpu().numpy(), inverse)
    # ME-shaped surface on numpy input (dataset_preprocess.py:269-274)
    feats = np.random.default_rng(1).normal(size=(len(xyz), 6)).astype(np.float32)
    qc, f, idx, inv = voxel.sparse_quantize(xyz, feats, quantization_size=0.02, return_index=True, return_inverse=True)
    assert np.array_equal(qc, vc[:, 1:]) and np.array_equal(idx, index) and np.array_equal(inv, inverse)
    assert np.array_equal(f, feats[index])
    assert np.array_equal(qc[inv], vo.quantize(xyz, 0.02)[:, 1:])


@pytest.mark.gpu
def test_voxelize_batched_mean_and_devoxelize_grad():
    import torch
    from pbnet_b200 import voxel
    rng = np.random.default_rng(3)
    clouds = [cloud(10 + b, 30_000 + 1000 * b, np.float32) / np.float32(0.02) for b in range(4)]  # voxel units, PBNet.py:236
    coords = voxel.batched_coordinates([torch.from_numpy(c) for c in clouds]).cuda()
    n = coords.shape[0]
    feats = torch.from_numpy(rng.normal(size=(n, 34)).astype(np.float32)).cuda()
    vc, index, inverse = vo.sparse_quantize(coords.cpu().numpy(), None)
    for mode in ("pick", "mean"):
        vf, vcoords, vm = voxel.voxelize(feats, coords, mode)
        assert np.array_equal(vcoords.cpu().numpy(), vc) and np.array_equal(vm.inverse.cpu().numpy(), inverse)
        want = vo.voxel_rows(feats.cpu().numpy(), inverse, len(vc), mode)
        np.testing.assert_allclose(vf.cpu().numpy(), want, rtol=1e-5, atol=1e-6)
    # devoxelize gather + autograd scatter-add (network/PBNet.py:130-134,250)
    for C in (32, 20, 3, 1):
        vfeat = torch.from_numpy(rng.normal(size=(len(vc), C)).astype(np.float32)).cuda().requires_grad_(True)
        out = voxel.devoxelize(vfeat, vm)
        assert torch.equal(out, vfeat[vm.inverse])
        g = torch.from_numpy(rng.normal(size=(n, C)).astype(np.float32)).cuda()
        out.backward(g)
        want = vo.voxel_rows(g.cpu().numpy(), inverse, len(vc), "sum")
        np.testing.assert_allclose(vfeat.grad.cpu().numpy(), want, rtol=1e-5, atol=1e-5)


@pytest.mark.gpu
def test_voxel_edge_cases():
    import torch
    from pbnet_b200 import voxel
    vm = voxel.voxel_map(torch.zeros((0, 3)), 0.02)
    assert vm.n_voxels == 0 and vm.inverse.shape[0] == 0
    one = torch.tensor([[0.019999, -0.02, 0.02]])
    vm = voxel.voxel_map(one.cuda(), 0.02)
    assert vm.vcoords.cpu().tolist() == vo.sparse_quantize(one.numpy(), 0.02)[0].tolist()
    same = torch.zeros((1000, 3)).cuda()
    vm = voxel.voxel_map(same, 0.02)
    assert vm.n_voxels == 1 and int(vm.index[0]) == 0
