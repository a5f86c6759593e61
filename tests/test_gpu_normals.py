"""GPU: mesh vertex normals (C ABI pb_cal_normal_line, the fourth op of the reference's PB_lib module,
lib/PB_lib/src/normal/cal_normal.cu) against the CPU restatement (oracle/pb_oracle.c::pb_oracle_normals) and, when the
compiled reference travels with the snapshot (oracle/_ref), against the reference itself — bit for bit."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _mesh(seed, V, F, degenerate=True):
    rng = np.random.Generator(np.random.PCG64(seed))
    xyz = rng.uniform(-3, 3, size=(V, 3)).astype(np.float32)
    face = rng.integers(0, V - 5, size=(F, 3)).astype(np.int32)       # the last 5 vertices stay without faces -> (0,0,1)
    if degenerate:
        face[3] = [7, 7, 9]                                             # repeated vertex: zero normal -> NaN, as in the reference
        face[5] = [11, 12, 11]
    return xyz, face


def _same(a, b):
    return np.array_equal(np.isnan(a), np.isnan(b)) and np.array_equal(a[~np.isnan(a)].view(np.uint32), b[~np.isnan(b)].view(np.uint32))


@pytest.mark.parametrize("seed,V,F,deg", [(1, 500, 1200, True), (2, 3000, 3000, False), (3, 2000, 9000, False), (4, 64, 0, False)])
def test_normals_equal_oracle(seed, V, F, deg):
    import torch
    from oracle import pb_oracle as po
    from pbnet_b200.shim import PB_lib
    xyz, face = _mesh(seed, V, max(F, 8), deg)
    face = face[:F] if F else face[:0]
    for nf in sorted({F, F // 2}):
        want = po.oracle_normals(xyz, face, nf)
        for dev in ("cpu", "cuda"):
            out = torch.zeros(V, 3, dtype=torch.float32, device=dev)
            PB_lib.cal_normal_line(torch.from_numpy(xyz).to(dev), torch.from_numpy(face).to(dev).reshape(-1, 3).contiguous(), out, V, nf)
            assert _same(out.cpu().numpy(), want), (seed, nf, dev)
    assert np.array_equal(want[-1], [0, 0, 1])


def test_wrapper_mirrors_reference_argument_quirk():
    from oracle import pb_oracle as po
    from pbnet_b200 import pbnet_ops
    xyz, face = _mesh(9, 800, 2000, False)
    got = pbnet_ops.get_normal_line(xyz, face).numpy()
    assert _same(got, po.oracle_normals(xyz, face, 800))      # num_face = V, lib/PB_lib/torch_io/pbnet_ops.py:163
    with pytest.raises(ValueError):
        pbnet_ops.get_normal_line(xyz, face[:100])             # fewer faces than vertices: out-of-bounds read in the reference


def test_normals_equal_compiled_reference():
    ref_dir = os.path.join(ROOT, "oracle", "_ref")
    sys.path.insert(0, ref_dir)
    try:
        import torch
        import PB_lib as REF  # noqa: N811  the UNMODIFIED compiled reference
    except Exception as e:  # pragma: no cover
        pytest.skip(f"compiled reference not available: {e}")
    finally:
        sys.path.remove(ref_dir)
    import torch
    from pbnet_b200.shim import PB_lib
    for seed, V, F in ((21, 700, 1500), (22, 2500, 2500), (23, 1500, 6000)):
        xyz, face = _mesh(seed, V, F, True)
        a = torch.zeros(V, 3, dtype=torch.float32)
        b = torch.zeros(V, 3, dtype=torch.float32)
        REF.cal_normal_line(torch.from_numpy(xyz), torch.from_numpy(face), a, V, F)
        PB_lib.cal_normal_line(torch.from_numpy(xyz), torch.from_numpy(face), b, V, F)
        assert _same(b.numpy(), a.numpy()), seed
