"""-m gpu: the zero-edit drop-in boundary, executed.

1. The UNMODIFIED reference wrapper (lib/PB_lib/torch_io/pbnet_ops.py, byte-compiled into oracle/_ref/ by
   oracle/build_ref.py because /root/reference does not exist on the GPU box) is imported against the shim module
   ``PB_lib`` (pbnet_b200/shim/PB_lib.py) and its ``cluster(...)`` is called on CPU tensors, as network/PBNet.py:176 does.
2. ``PB_lib.binary_cluster`` is called directly with the 20 positional arguments of lib/PB_lib/src/pbnet/cluster.h:13-18,
   checking the in-place outputs and the ``resize_`` of ``center`` / ``clt_sem`` (cluster.cu:112-118).
3. The same 20-argument call goes to the live compiled reference (oracle/_ref/PB_lib*.so) and must agree bit for bit.
"""
import importlib.machinery
import importlib.util
import os
import sys

import numpy as np
import pytest

from tests import helpers as H

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "oracle", "_ref")


def _shim():
    import pbnet_b200
    pbnet_b200.install_shim()
    sys.modules.pop("PB_lib", None)
    import PB_lib
    assert PB_lib.__file__.startswith(pbnet_b200.shim_dir()), PB_lib.__file__
    return PB_lib


def _reference_wrapper():
    """The reference's own pbnet_ops module object, loaded from its byte code with the shim as ``PB_lib``."""
    pyc = os.path.join(REF_DIR, "ref_pbnet_ops.pyc.bin")
    if not os.path.exists(pyc):
        pytest.skip("oracle/_ref/ref_pbnet_ops.pyc.bin missing (run __graft_entry__.build() where /root/reference exists)")
    shim = _shim()
    loader = importlib.machinery.SourcelessFileLoader("ref_pbnet_ops", pyc)
    spec = importlib.util.spec_from_loader("ref_pbnet_ops", loader)
    m = importlib.util.module_from_spec(spec)
    loader.exec_module(m)
    assert m.PB_lib is shim
    return m


def _marshal(c):
    """Arguments exactly as lib/PB_lib/torch_io/pbnet_ops.py:16-50 builds them."""
    import torch
    t = torch.from_numpy
    xs, xo = t(c["xyz_shift"]), t(c["xyz_orig"])
    x, y, z = (xs[:, i].type(torch.float32).contiguous() for i in range(3))
    l1 = torch.abs(x) + torch.abs(y) + torch.abs(z)
    bp = t(np.ascontiguousarray(c["seg_counts"], np.int32))
    imap = torch.cat([torch.arange(0, int(b), 1) for b in bp], dim=0).type(torch.int32).contiguous()
    ox, oy, oz = (xo[:, i].type(torch.float32).contiguous() for i in range(3))
    sem = t(np.ascontiguousarray(c["sem"])).type(torch.int32).contiguous()
    n = xs.shape[0]
    r = (torch.ones(18) * 0.04).type(torch.float32).contiguous()
    m = (torch.ones(18) * 31).type(torch.int32).contiguous()
    cid = (torch.ones(n) * -1).type(torch.int32).contiguous()
    cnum = torch.zeros([bp.shape[0]]).type(torch.int32).contiguous()
    den = torch.zeros(n, dtype=torch.int32)
    cen = torch.zeros(n, dtype=torch.float32)
    cs = torch.zeros(n, dtype=torch.int32)
    return [x, y, z, l1, imap, ox, oy, oz, sem, bp, r, m, cid, cnum, den, cen, cs, int(bp.shape[0]), 0.05, True]


def _three_copy_calls(seed=91, npts=60000):
    from pbnet_b200 import scenes
    return scenes.class_calls(scenes.make_scene(seed, npts), 3)


def test_unmodified_reference_wrapper_runs_on_the_shim():
    import torch
    from oracle import pb_oracle as po
    m = _reference_wrapper()
    n_calls = 0
    for c in _three_copy_calls():
        want = po.oracle_cluster(c["xyz_shift"], c["xyz_orig"], c["sem"], c["seg_counts"], 0.04, 31)
        got = m.cluster(torch.from_numpy(c["xyz_shift"]), torch.from_numpy(c["xyz_orig"]), torch.from_numpy(c["sem"]),
                        torch.from_numpy(np.ascontiguousarray(c["seg_counts"], np.int32)), 0.04, 31, 3)
        assert len(got) == 4
        for g, w, name in zip(got, want, ("cluster_id", "cluster_num", "den_queue+1", "center")):
            assert not g.is_cuda
            assert np.array_equal(g.numpy().view(np.uint32), np.asarray(w).view(np.uint32)), f"class {c['sem_id']} {name}"
        assert got[3].shape[0] == 3 * int(got[1].sum())      # center resized to 3K (cluster.cu:112-118)
        n_calls += 1
    assert n_calls > 5


def test_binary_cluster_20_positional_arguments_in_place():
    from oracle import pb_oracle as po
    shim = _shim()
    for c in _three_copy_calls(92, 40000)[:4]:
        a = _marshal(c)
        n = a[0].shape[0]
        ptrs = [t.data_ptr() for t in (a[12], a[13], a[14])]
        assert shim.binary_cluster(*a) is None
        cid, cnum, den, cen, cs = a[12:17]
        assert [t.data_ptr() for t in (cid, cnum, den)] == ptrs           # written in place
        want = po.oracle_binary_cluster(c["xyz_shift"], c["xyz_orig"], c["sem"], c["seg_counts"], H.R18, H.M18)
        k = int(want["cluster_num"].sum())
        assert cen.shape[0] == 3 * k and cs.shape[0] == k                  # resize_ of cluster.cu:112-118
        got = dict(cluster_id=cid.numpy(), cluster_num=cnum.numpy(), den_queue=den.numpy(), center=cen.numpy(), clt_sem=cs.numpy())
        assert H.diff_report(got, want) == [], f"class {c['sem_id']}"
        assert n == len(c["sem"])
    # batch_size smaller than ins_bp: only the first batch_size segments are processed (cluster.cu:57 loops batch_size)
    c = _three_copy_calls(92, 40000)[0]
    a = _marshal(c)
    n1 = int(c["seg_counts"][0])
    a2 = [t[:n1].contiguous() if hasattr(t, "shape") and t.shape[0] == a[0].shape[0] else t for t in a]
    a2[17] = 1
    shim.binary_cluster(*a2)
    want = po.oracle_binary_cluster(c["xyz_shift"][:n1], c["xyz_orig"][:n1], c["sem"][:n1], [n1], H.R18, H.M18)
    assert np.array_equal(a2[12].numpy(), want["cluster_id"]) and int(a2[13][0]) == int(want["cluster_num"][0])
    assert a2[13][1:].tolist() == [0, 0]


def test_shim_agrees_with_live_compiled_reference():
    from tests.test_gpu_iou import load_reference_module
    shim = _shim()
    ref = load_reference_module()            # oracle/_ref/PB_lib*.so, the UNMODIFIED compiled reference
    if ref is None:
        pytest.skip("compiled reference (oracle/_ref) not built")
    assert ref is not shim and hasattr(ref, "binary_cluster")
    for c in _three_copy_calls(93, 50000)[:5]:
        a, b = _marshal(c), _marshal(c)
        shim.binary_cluster(*a)
        ref.binary_cluster(*b)
        for i, name in zip(range(12, 17), ("cluster_id", "cluster_num", "den_queue", "center", "clt_sem")):
            assert a[i].shape == b[i].shape, name
            assert np.array_equal(a[i].numpy().view(np.uint32), b[i].numpy().view(np.uint32)), f"class {c['sem_id']} {name}"
