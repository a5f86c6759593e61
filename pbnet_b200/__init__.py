"""pbnet_b200 — B200-native (sm_100a) implementation of PBNet's instance-grouping hot path behind the
reference's operator surface.  See DESIGN.md / INTEGRATION.md."""
import os
import sys

__all__ = ["install_shim", "shim_dir"]


def shim_dir() -> str:
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "shim")


def install_shim() -> None:
    """Makes ``import PB_lib`` resolve to pbnet_b200/shim/PB_lib.py so the reference's unmodified
    lib/PB_lib/torch_io/pbnet_ops.py runs on this library."""
    d = shim_dir()
    if d not in sys.path:
        sys.path.insert(0, d)
