"""CPU: the C-ABI library loads and exports every symbol include/pbnet_b200.h declares; host-side logic
(workload assembly, sharding, generator determinism, drop-in module surface).  No compute calls."""
import ctypes
import hashlib
import inspect
import os
import re
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "pbnet_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pb_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_header_symbol():
    from pbnet_b200 import _lib, build
    build.build()
    L = ctypes.CDLL(build.SO)
    names = header_functions()
    assert len(names) >= 11
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/pbnet_b200.h but not exported"
    assert sorted(_lib.SYMBOLS) == names


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from pbnet_b200._lib import PBError
    from pbnet_b200.cluster import Context
    with pytest.raises(PBError) as e:
        Context(0)
    assert e.value.code == 2  # PB_ERR_CUDA: the product path fails loudly, it never routes to the oracle


def test_product_package_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "pbnet_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "pb_oracle" not in txt and "from oracle" not in txt and "import oracle" not in txt, f


def test_scene_generator_is_deterministic():
    from pbnet_b200 import scenes
    a, b = scenes.make_scene(22, 20000), scenes.make_scene(22, 20000)
    for k in ("xyz_orig", "offset", "sem"):
        assert np.array_equal(a[k], b[k])
    h = hashlib.sha1(a["xyz_orig"].tobytes() + a["offset"].tobytes() + a["sem"].tobytes()).hexdigest()
    assert h == hashlib.sha1(b["xyz_orig"].tobytes() + b["offset"].tobytes() + b["sem"].tobytes()).hexdigest()
    sizes = scenes.scene_sizes()
    assert len(sizes) == 312 and sizes.min() >= 50_000 and sizes.max() <= 250_000
    calls = scenes.class_calls(a, 3)
    for c in calls:
        assert c["seg_counts"].tolist() == [len(c["index"])] * 3 and 2 <= c["sem_id"] <= 19
        assert np.array_equal(c["xyz_shift"][:len(c["index"])],
                              (scenes.rotate_copies(a["xyz_orig"])[0][c["index"]] + scenes.rotate_copies(a["offset"])[0][c["index"]]).astype(np.float32))


def test_workload_and_sharding():
    from pbnet_b200 import scenes, workload
    sizes = np.minimum(scenes.scene_sizes(12), 20000)
    w = workload.build(range(12), sizes, 1, workers=1, cache_dir=None)
    assert int(w["seg_counts"].sum()) == int(w["n_points"]) == len(w["x"]) == len(w["sem"])
    assert int(w["call_seg_counts"].sum()) == len(w["seg_counts"]) and int(w["call_points"].sum()) == int(w["n_points"])
    # single class per segment (what PBNet feeds, network/PBNet.py:151-179)
    o = 0
    for n in w["seg_counts"]:
        assert w["sem"][o:o + n].min() == w["sem"][o:o + n].max()
        o += n
    shards = workload.shard_scenes(sizes, 4)
    assert sorted(sum(shards, [])) == list(range(12))
    loads = [int(sizes[s].sum()) for s in shards]
    assert max(loads) - min(loads) <= int(sizes.max())
    # sharded workloads concatenate to the full one (scene order inside a shard is ascending)
    parts = [workload.build(s, sizes, 1, workers=1, cache_dir=None) for s in shards]
    assert sum(int(p["n_points"]) for p in parts) == int(w["n_points"])


def test_dropin_surfaces_match_reference():
    """Names / arity of the reference operator surface (lib/PB_lib/src/PB_lib_api.cpp:7-10,
    lib/PB_lib/torch_io/pbnet_ops.py:14,82)."""
    import pbnet_b200
    from pbnet_b200 import pbnet_ops
    pbnet_b200.install_shim()
    sys.modules.pop("PB_lib", None)
    import PB_lib
    for n in ("binary_cluster", "get_iou", "cal_iou_and_masklabel", "cal_normal_line"):
        assert callable(getattr(PB_lib, n))
    assert len(inspect.signature(PB_lib.binary_cluster).parameters) == 20
    assert list(inspect.signature(pbnet_ops.Cluster.forward).parameters)[1:] == [
        "ins_offseted", "ins_orig", "sem", "ins_bp", "radius", "min_pts", "batch_size"]
    assert len(inspect.signature(PB_lib.cal_normal_line).parameters) == 5  # lib/PB_lib/src/normal/cal_normal.h:10
    assert callable(pbnet_ops.get_normal_line)
    assert len(inspect.signature(PB_lib.get_iou).parameters) == 7
    assert len(inspect.signature(PB_lib.cal_iou_and_masklabel).parameters) == 10
    ref_ops = "/root/reference/lib/PB_lib/torch_io/pbnet_ops.py"
    if os.path.exists(ref_ops):  # the UNMODIFIED reference wrapper imports against the shim
        import importlib.util
        spec = importlib.util.spec_from_file_location("ref_pbnet_ops", ref_ops)
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        assert m.PB_lib is PB_lib and callable(m.cluster)
    sys.modules.pop("PB_lib", None)


def test_weak_scaling_set_extends_the_single_gpu_set():
    """bench.py --scaling weak: the 312*N-scene set starts with the N=1 set (same sizes, same seeds), and the LPT
    partition covers every scene exactly once with balanced point counts."""
    from pbnet_b200 import scenes, workload
    s1, s8 = scenes.scene_sizes(312), scenes.scene_sizes(312 * 8)
    assert np.array_equal(s8[:312], s1)
    shards = workload.shard_scenes(s8, 8)
    flat = sorted(i for sh in shards for i in sh)
    assert flat == list(range(312 * 8))
    loads = np.array([int(s8[sh].sum()) for sh in shards], np.float64)
    assert loads.max() / loads.min() < 1.01


def test_local_scene_threshold_and_mask_helpers():
    import torch
    from pbnet_b200 import evalpost, grouping
    from pbnet_b200.scenes import COUNT_MEAN
    thr = grouping._big_thresholds(COUNT_MEAN)
    want = (torch.tensor(COUNT_MEAN) * 0.2).numpy()      # network/PBNet.py:210: an fp32 tensor product
    assert thr.dtype == np.float32 and np.array_equal(thr, want)
    assert thr[2] == np.float32(783.4000244140625)        # 3917 * 0.2 in fp32
    lab = torch.tensor([0, -100, 1, 1, 0], dtype=torch.int32)
    m = evalpost.dense_masks(lab, 2)
    assert m.tolist() == [[1, 0, 0, 0, 1], [0, 0, 1, 1, 0]]


def test_numa_helpers_on_a_host_without_cuda():
    """sharding.bind_to_gpu_numa_node degrades to None where the topology is not exposed; the cpulist parser handles ranges."""
    from pbnet_b200 import sharding
    assert sharding._parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert sharding._parse_cpulist("") == set()
    import torch
    if not torch.cuda.is_available():
        assert sharding.bind_to_gpu_numa_node(0) is None
