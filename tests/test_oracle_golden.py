"""CPU: the oracle (oracle/pb_oracle.c, both variants) against the committed golden vectors produced by
the UNMODIFIED compiled reference on a B200 (tests/golden/make_golden.py).  This is what pins the oracle."""
import numpy as np
import pytest

from oracle import pb_oracle as po
from tests import helpers as H


@pytest.mark.parametrize("path", H.golden_files(), ids=lambda p: p.split("/")[-1][:-4])
def test_oracle_grid_matches_reference(path):
    d, ref = H.load_golden(path)
    got = po.oracle_binary_cluster(d["xyz_shift"], d["xyz_orig"], d["sem"], d["seg_counts"], d["radius"], d["min_pts"],
                                   0.05, bool(d["nv_flag"]), mode="grid")
    assert H.diff_report(got, ref) == []


@pytest.mark.parametrize("path", [p for p in H.golden_files() if np.load(p)["sem"].shape[0] <= 4000],
                         ids=lambda p: p.split("/")[-1][:-4])
def test_oracle_literal_matches_reference(path):
    d, ref = H.load_golden(path)
    got = po.oracle_binary_cluster(d["xyz_shift"], d["xyz_orig"], d["sem"], d["seg_counts"], d["radius"], d["min_pts"],
                                   0.05, bool(d["nv_flag"]), mode="literal")
    assert H.diff_report(got, ref) == []


def test_golden_set_is_complete():
    files = H.golden_files()
    assert len(files) >= 40
    names = {f.split("/")[-1] for f in files}
    for must in ("mixed_1seg.npz", "mixed_2seg_minpts.npz", "mixed_novote.npz", "single_point.npz", "dups200_c17.npz",
                 "blob700_c10_empty_segs.npz"):
        assert must in names
