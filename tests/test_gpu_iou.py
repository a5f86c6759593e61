"""-m gpu: get_iou / cal_iou_and_masklabel against the numpy restatement and, when the
compiled reference travelled to the box (oracle/_ref), against the UNMODIFIED reference kernels live."""
import importlib.util
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def make_case(seed, n_points=60000, n_inst=37, n_prop=90):
    rng = np.random.Generator(np.random.PCG64(seed))
    labels = rng.integers(0, n_inst, size=n_points).astype(np.int64)
    labels[rng.random(n_points) < 0.1] = -100
    pnum = np.bincount(labels[labels >= 0], minlength=n_inst).astype(np.int32)
    idx, off = [], [0]
    for p in range(n_prop):
        if p % 11 == 0:
            members = np.zeros(0, np.int64)                                   # empty proposal
        elif p % 3 == 0:
            k = int(rng.integers(0, n_inst))                                  # mostly one instance
            pool = np.nonzero(labels == k)[0]
            members = rng.choice(pool, size=min(len(pool), int(rng.integers(50, 1500))), replace=False)
            members = np.concatenate([members, rng.integers(0, n_points, size=len(members) // 5)])
        else:
            members = rng.integers(0, n_points, size=int(rng.integers(1, 3000)))
        idx.append(members.astype(np.int32))
        off.append(off[-1] + len(members))
    idx = np.concatenate(idx).astype(np.int32)
    scores = rng.random(len(idx)).astype(np.float32)
    return idx, np.asarray(off, np.int32), labels, pnum, scores


def load_reference_module():
    import glob
    so = glob.glob(os.path.join(ROOT, "oracle", "_ref", "PB_lib*.so"))
    if not so:
        return None
    import torch  # noqa: F401  (libtorch must be loaded first)
    spec = importlib.util.spec_from_file_location("PB_lib", so[0])
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


@pytest.mark.parametrize("seed", [1, 2])
def test_get_iou_and_mask_label(seed):
    import torch
    from oracle import iou_oracle as io
    from pbnet_b200 import pbnet_ops
    idx, off, labels, pnum, scores = make_case(seed)
    t = lambda a: torch.from_numpy(a).cuda()
    ref = load_reference_module()
    # get_iou (mode 0 over whole proposals)
    got = pbnet_ops.get_iou(t(idx), t(off), t(labels), t(pnum)).cpu().numpy()
    want = io.get_iou(idx, off, labels, pnum)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    if ref is not None:
        r = torch.zeros((len(off) - 1, len(pnum)), dtype=torch.float32, device="cuda")
        ref.get_iou(t(idx), t(off), t(labels), t(pnum), r, len(pnum), len(off) - 1)
        torch.cuda.synchronize()
        assert np.array_equal(r.cpu().numpy().view(np.uint32), got.view(np.uint32))
    # cal_iou_and_masklabel, both modes
    for mode in (0, 1):
        iou, ml = pbnet_ops.cal_iou_and_masklabel(t(idx), t(off), t(labels), t(pnum), t(scores).view(-1, 1), mode)
        w_iou = io.get_iou(idx, off, labels, pnum, scores, mode)
        w_ml = io.mask_label(idx, off, labels, w_iou, np.full(len(idx), -1.0, np.float32))
        assert np.array_equal(iou.cpu().numpy().view(np.uint32), w_iou.view(np.uint32))
        assert np.array_equal(ml.cpu().numpy().reshape(-1), w_ml)
        if ref is not None:
            r = torch.zeros((len(off) - 1, len(pnum)), dtype=torch.float32, device="cuda")
            rml = torch.full((len(idx), 1), -1.0, dtype=torch.float32, device="cuda")
            ref.cal_iou_and_masklabel(t(idx), t(off), t(labels), t(pnum), r, len(pnum), len(off) - 1,
                                      t(scores).view(-1, 1), rml, mode)
            torch.cuda.synchronize()
            assert np.array_equal(r.cpu().numpy().view(np.uint32), iou.cpu().numpy().view(np.uint32))
            assert np.array_equal(rml.cpu().numpy(), ml.cpu().numpy())


def test_shim_signatures_run():
    """The reference-facing module surface (PB_lib.get_iou / cal_iou_and_masklabel with in-place outputs)."""
    import sys
    import torch
    import pbnet_b200
    from oracle import iou_oracle as io
    pbnet_b200.install_shim()
    sys.modules.pop("PB_lib", None)
    import PB_lib
    idx, off, labels, pnum, scores = make_case(3, 20000, 12, 30)
    t = lambda a: torch.from_numpy(a).cuda()
    out = torch.zeros((len(off) - 1, len(pnum)), dtype=torch.float32, device="cuda")
    PB_lib.get_iou(t(idx), t(off), t(labels), t(pnum), out, len(pnum), len(off) - 1)
    assert np.array_equal(out.cpu().numpy().view(np.uint32), io.get_iou(idx, off, labels, pnum).view(np.uint32))
    sys.modules.pop("PB_lib", None)
